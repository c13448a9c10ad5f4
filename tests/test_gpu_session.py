"""GPU parity of the whole incremental-session path against the CPU oracle and the committed reference goldens.

Tolerances (BASELINE.json north_star): fp32 head / projection 1e-5 (kernel-level tests); bf16 tensor-core convolutions:
features and everything downstream are compared at the tolerance bf16 operands allow (stated per assert).
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _world(golden, word_embed_dir, **over):
    from srb200 import synthetic
    o = dict(golden['overrides'])
    o.update(over)
    return synthetic.make_world(golden['seed'], n_sessions=golden['n_sessions'], n_base_batch=golden['n_base_batch'],
                                word_embed_path=word_embed_dir, **o)


def _run_b200(world, n_sessions, ckpt_extra=None):
    import contextlib
    import io
    from eval.language_eval import few_shot_finetune_incremental_test
    from models.util import create_model
    from srb200 import synthetic
    opt = world.opt
    opt.n_sessions_override = n_sessions
    net = synthetic.init_model(create_model, opt, world.seed)
    ckpt = synthetic.make_ckpt(net, world)
    ckpt.update(ckpt_extra or {})
    net = net.cuda()
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        novel, base = few_shot_finetune_incremental_test(net, ckpt, torch.nn.CrossEntropyLoss(), world.meta_valloader,
                                                         world.base_val_loader, opt,
                                                         base_support_loader=world.base_support_loader)
    rec = few_shot_finetune_incremental_test.last_record
    rec['stdout'] = buf.getvalue()
    return rec, net


@pytest.mark.parametrize("case", ["subspace_s2e3", "semantic_s2e2", "mapping_s2e2"])
def test_session_vs_reference_golden(case, golden_dir, word_embed_dir):
    """Short runs (2 sessions x 2-3 epochs) against the UNMODIFIED reference's recorded outputs."""
    g = torch.load(os.path.join(golden_dir, case + ".pt"), weights_only=False)
    world = _world(g, word_embed_dir)
    rec, net = _run_b200(world, g['n_sessions'], g.get('ckpt_extra'))
    ref = g['reference']
    assert rec['counters'] == ref['counters']                      # BasicBlock.num_batches_tracked bookkeeping
    for s, (a, b) in enumerate(zip(rec['sessions'], ref['sessions'])):
        assert a['epochs'] == b['epochs']
        # loss terms: CE terms inherit the bf16 feature error, the regulariser terms are pure fp32
        np.testing.assert_allclose(a['terms'][:, 0], b['terms'][:, 0], rtol=2e-2, err_msg="total loss s%d" % s)
        np.testing.assert_allclose(a['terms'][:, 3:6], b['terms'][:, 3:6], rtol=2e-3, atol=1e-6, err_msg="reg terms s%d" % s)
        wa, wb = a['W'].cpu(), b['W']
        assert ((wa - wb).norm() / wb.norm()).item() < 2e-3, "classifier weights s%d" % s
        # BN running statistics after the session's train-mode pass
        for k, v in b['bn'].items():
            u = a['bn'][k].cpu()
            if 'num_batches_tracked' in k:
                assert int(u) == int(v), k
            else:
                assert ((u - v).norm() / (v.norm() + 1e-12)).item() < 2e-2, k
        pf = a['probe_feat'].cpu()
        assert ((pf - b['probe_feat']).norm() / b['probe_feat'].norm()).item() < 3e-2, "eval features s%d" % s
        assert a['vocab_novel'] == b['vocab_novel']
    np.testing.assert_allclose(rec['weighted'], ref['weighted'], atol=2.0)


def test_session_vs_oracle_converged(golden_dir, word_embed_dir):
    """Config 1 (one session run to the reference's stopping rule) against the oracle on the same inputs: epoch
    count, loss trace, predictions and accuracies."""
    from oracle import init as oinit, session
    from srb200 import synthetic
    seed = 1
    world = synthetic.make_world(seed, n_sessions=1, n_base_batch=64, word_embed_path=word_embed_dir)
    sd = oinit.init_state_dict(seed)
    orec = session.run_sessions(sd, world, n_sessions=1, schedule='cached')
    world2 = synthetic.make_world(seed, n_sessions=1, n_base_batch=64, word_embed_path=word_embed_dir)
    rec, net = _run_b200(world2, 1)
    a, b = rec['sessions'][0], orec['sessions'][0]
    print("epochs b200 %d oracle %d" % (a['epochs'], b['epochs']))
    n = min(a['epochs'], b['epochs'])
    rel = np.abs(a['terms'][:n, 0] - b['terms'][:n, 0]) / np.abs(b['terms'][:n, 0])
    print("loss trace max rel err %.3e" % rel.max())
    assert rel.max() < 2e-2
    # the stopping rule is |dloss| < 1e-4 ten times in a row: bf16 feature noise may move it by a few epochs
    assert abs(a['epochs'] - b['epochs']) <= max(10, int(0.05 * b['epochs']))
    agree = np.mean([(p.numpy() == q.numpy()).mean() for p, q in zip(a['query_pred'], b['query_pred'])])
    lo = b['query_logits'][0]
    top2 = lo.topk(2, 1).values
    print("query prediction agreement %.4f; oracle min top1-top2 margin %.3e" % (agree, (top2[:, 0] - top2[:, 1]).min().item()))
    assert agree >= 0.97
    assert abs(rec['novel'][0] - orec['novel'][0]) <= 3.0
    assert abs(rec['base'][0] - orec['base'][0]) <= 3.2
