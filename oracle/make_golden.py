"""Generate tests/golden/*.pt by running the UNMODIFIED reference (/root/reference) on CPU in the build container.

    python -m oracle.make_golden            # all fixtures (a few minutes on 8 cores)

Each fixture holds the inputs' recipe (seed + option overrides; the synthetic world is regenerated from it) and the
reference's outputs: per-epoch loss terms, classifier-weight trajectory, BatchNorm buffers after each session,
eval-mode probe features, query predictions and the accuracy lists.  The word-embedding fixture is a re-packing of
the reference's word_embeds/miniImageNet_dim500.pickle (input data, 139 words x 500 floats).
"""
import os
import pickle
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "subspace-reg_b200"))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (seed, n_sessions, n_base_batch, option overrides)
    "subspace_s2e3": (1, 2, 64, dict(max_novel_epochs=3)),
    "semantic_s2e2": (2, 2, 32, dict(max_novel_epochs=2, attraction_override=None, glove=True, label_pull=0.2,
                                     temperature=3.0)),
    "mapping_s2e2": (3, 2, 32, dict(max_novel_epochs=2, attraction_override="mapping_linear_label2image", glove=True,
                                    label_pull=0.1)),
    # DropBlock with the reference's default block_size = 5 (no --no_dropblock), and the Adam branch of get_optim
    "dropblock_s2e3": (4, 2, 32, dict(max_novel_epochs=3, no_dropblock=False)),
    "adam_s2e3": (5, 2, 32, dict(max_novel_epochs=3, adam=True, weight_decay=5e-3)),
}


def word_embed_fixture():
    src = "/root/reference/word_embeds/miniImageNet_dim500.pickle"
    with open(src, "rb") as f:
        d = pickle.load(f)
    words = sorted(d.keys())
    np.savez_compressed(os.path.join(GOLD, "word_embeds_dim500.npz"), words=np.array(words),
                        vectors=np.stack([d[w] for w in words]).astype(np.float32))


def slim(rec):
    out = dict(weighted=rec.get('weighted'), novel=rec.get('novel'), base=rec.get('base'),
               acc_novel_avg=rec['acc_novel_avg'], acc_base_avg=rec['acc_base_avg'], counters=rec['counters'], sessions=[])
    for i, s in enumerate(rec['sessions']):
        t = dict(epochs=s['epochs'], terms=s['terms'], W=s['W'], novel_session_acc=s['novel_session_acc'],
                 query_pred=s['query_pred'], acc_base=s['acc_base'], vocab_novel=s['vocab_novel'], bn=s['bn'],
                 probe_feat=s.get('probe_feat'))
        if i == 0:
            t['W_traj'] = s['W_traj']
        out['sessions'].append(t)
    return out


def main(names):
    from oracle import reference_harness
    from srb200 import synthetic
    os.makedirs(GOLD, exist_ok=True)
    word_embed_fixture()
    for name in names:
        seed, n_sessions, n_base_batch, over = CASES[name]
        t0 = time.time()
        world = synthetic.make_world(seed, n_sessions=n_sessions, n_base_batch=n_base_batch,
                                     word_embed_path="/root/reference/word_embeds", **over)
        extra = {}
        if over.get("attraction_override") == "mapping_linear_label2image":
            g = torch.Generator().manual_seed(1000 + seed)
            extra["mapping_linear_label2image"] = {"map.weight": torch.randn(640, 300, generator=g) * 0.05,
                                                   "map.bias": torch.randn(640, generator=g) * 0.01}
        orig_make_ckpt = synthetic.make_ckpt
        synthetic.make_ckpt = lambda model, w: dict(orig_make_ckpt(model, w), **extra)
        try:
            rec, sd0 = reference_harness.run_reference(world, n_sessions, seed, probe_rows=8)
        finally:
            synthetic.make_ckpt = orig_make_ckpt
        fixture = dict(case=name, seed=seed, n_sessions=n_sessions, n_base_batch=n_base_batch, overrides=over,
                       ckpt_extra=extra, reference=slim(rec), torch_version=torch.__version__,
                       init_checksum={k: float(v.double().sum()) for k, v in sd0.items() if v.dtype.is_floating_point})
        torch.save(fixture, os.path.join(GOLD, name + ".pt"))
        print("%s: %.0f s, epochs %s, weighted %s" % (name, time.time() - t0, [s['epochs'] for s in rec['sessions']],
                                                     rec.get('weighted')), flush=True)


if __name__ == "__main__":
    main(sys.argv[1:] or list(CASES))
