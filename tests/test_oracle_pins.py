"""The oracle is pinned against outputs of the UNMODIFIED reference (tests/golden/*.pt, made by oracle/make_golden.py
in the build container): same torch ops in the same order on the same inputs -> bit-identical on the same torch
build, and within 1e-5 elsewhere."""
import os

import numpy as np
import pytest
import torch


def _world(golden, word_embed_dir):
    from srb200 import synthetic
    return synthetic.make_world(golden['seed'], n_sessions=golden['n_sessions'], n_base_batch=golden['n_base_batch'],
                                word_embed_path=word_embed_dir, **golden['overrides'])


def _close(a, b, tol=1e-5):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return ((a - b).abs().max() <= tol * (b.abs().max() + 1e-12)).item()


@pytest.mark.parametrize("case,schedule", [("subspace_s2e3", "cached"), ("semantic_s2e2", "cached"),
                                           ("mapping_s2e2", "literal"), ("dropblock_s2e3", "cached"),
                                           ("adam_s2e3", "cached")])
def test_oracle_matches_reference_golden(case, schedule, golden_dir, word_embed_dir):
    from oracle import init as oinit, session
    g = torch.load(os.path.join(golden_dir, case + ".pt"), weights_only=False)
    sd = oinit.init_state_dict(g['seed'])
    for k, v in g['init_checksum'].items():            # same random init as the reference's create_model
        assert abs(float(sd[k].double().sum()) - v) <= 1e-9 * max(1.0, abs(v)), k
    world = _world(g, word_embed_dir)
    rec = session.run_sessions(sd, world, n_sessions=g['n_sessions'], schedule=schedule, ckpt=g.get('ckpt_extra'),
                               keep_w_trajectory=True, probe_rows=8)
    ref = g['reference']
    same_build = torch.__version__ == g['torch_version']
    tol = 0.0 if same_build else 1e-5
    for a, b in zip(rec['sessions'], ref['sessions']):
        assert a['epochs'] == b['epochs']
        assert _close(a['terms'][:, :6], b['terms'], max(tol, 1e-12))
        assert _close(a['W'], b['W'], max(tol, 1e-12))
        if 'W_traj' in b:
            assert _close(a['W_traj'], b['W_traj'], max(tol, 1e-12))
        for k, v in b['bn'].items():
            assert _close(a['bn'][k], v, max(tol, 1e-12)), k
        assert _close(a['probe_feat'], b['probe_feat'], max(tol, 1e-12))
        for p, q in zip(a['query_pred'], b['query_pred']):
            assert (p == q).all()
        assert a['novel_session_acc'] == b['novel_session_acc']
        assert a['vocab_novel'] == b['vocab_novel']
    assert rec['counters'] == ref['counters']
    assert rec['weighted'] == ref['weighted'] and rec['novel'] == ref['novel'] and rec['base'] == ref['base']
    assert abs(rec['acc_novel_avg'] - ref['acc_novel_avg']) < 1e-9 and abs(rec['acc_base_avg'] - ref['acc_base_avg']) < 1e-9


def test_get_embeds_oov_and_multiword(word_embed_dir):
    """models/util.py:50-67 semantics: mean over words, an OOV word zeroes the running sum, float64 promotion."""
    from oracle import regularizer as rg
    path = os.path.join(word_embed_dir, "miniImageNet_dim500.pickle")
    e = rg.get_embeds(path, ["komondor", "house finch", "robin"])
    assert e.dtype == torch.float64 and e.shape == (3, 500)
    assert float(e[0].abs().sum()) == 0.0
    import pickle
    t = pickle.load(open(path, "rb"))
    np.testing.assert_array_equal(e[1].float().numpy(), (t["house"] + t["finch"]) / 2)
    e2 = rg.get_embeds(path, ["robin"])
    assert e2.dtype == torch.float32
