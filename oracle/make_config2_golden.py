"""Generate tests/golden/config2_seed{N}.pt: BASELINE config 2 (60 base + 8 sessions x 5-way 5-shot, memory_replay 1,
n_base_support_samples 1, base batch 1000, stopping rule live) run END TO END by the CPU oracle.

    python -m oracle.make_config2_golden 1 2        # ~15 min per seed on 8 cores

The oracle's 'cached' schedule is used: it is asserted equal to the 'literal' schedule in tests/test_oracle_pins.py and
the oracle itself is pinned bit for bit to the unmodified reference on the short goldens (oracle/make_golden.py); the
literal schedule of a full sweep is ~25 h of CPU per seed.

Per session the fixture holds what north_star asks to be compared: epoch count, the loss-term trace, final classifier
weight, BatchNorm buffers, query / base predictions, and the oracle's top-1 / top-2 logit margin for every scored image
(so that any prediction mismatch of a lower-precision path can be judged against the margin), plus 8 eval-mode probe
features.
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "subspace-reg_b200"))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def margins(logits):
    t = logits.topk(2, 1).values
    return (t[:, 0] - t[:, 1]).float()


def main(seeds, n_sessions=8, n_base_batch=1000, tag="config2", **over):
    from oracle import init as oinit, session
    from srb200 import synthetic
    torch.set_num_threads(os.cpu_count())
    for seed in seeds:
        t0 = time.time()
        world = synthetic.make_world(seed, n_sessions=n_sessions, n_base_batch=n_base_batch,
                                     word_embed_path="/root/reference/word_embeds", **over)
        sd = oinit.init_state_dict(seed)
        rec = session.run_sessions(sd, world, n_sessions=n_sessions, schedule='cached', verbose=True, probe_rows=8)
        out = dict(case="%s_seed%d" % (tag, seed), seed=seed, n_sessions=n_sessions, n_base_batch=n_base_batch,
                   overrides=over, torch_version=torch.__version__, weighted=rec['weighted'], novel=rec['novel'],
                   base=rec['base'], base0=rec['base0'], acc_novel_avg=rec['acc_novel_avg'],
                   acc_base_avg=rec['acc_base_avg'], counters=rec['counters'], sessions=[])
        for s in rec['sessions']:
            out['sessions'].append(dict(
                epochs=s['epochs'], terms=s['terms'].astype(np.float32), W=s['W'], novel_session_acc=s['novel_session_acc'],
                query_pred=[p.to(torch.int16) for p in s['query_pred']], base_pred=s['base_pred'].to(torch.int16),
                query_margin=[margins(q) for q in s['query_logits']], base_margin=margins(s['base_logits']),
                acc_base=s['acc_base'], memory_inds=s['memory_inds'], vocab_novel=s['vocab_novel'],
                bn={k: v for k, v in s['bn'].items()}, probe_feat=s['probe_feat']))
        torch.save(out, os.path.join(GOLD, "%s_seed%d.pt" % (tag, seed)))
        print("seed %d: %.0f s, epochs %s, weighted %s" % (seed, time.time() - t0, [s['epochs'] for s in rec['sessions']],
                                                         rec['weighted']), flush=True)


if __name__ == "__main__":
    main([int(a) for a in sys.argv[1:]] or [1])
