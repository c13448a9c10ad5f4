"""BASELINE config 2 END TO END on the GPU against the CPU oracle: 60 base classes + 8 sessions x 5-way 5-shot, memory
replay, base batch 1000, every session fine-tuned to the reference's stopping rule
(reference eval/language_eval.py:145-454; fixture from oracle/make_config2_golden.py).

north_star's bars, asserted per session:
  * parity tier (error-compensated bf16x3 convolutions + fp32 head): IDENTICAL class predictions for every query and
    base image - except where the oracle's own top-1 / top-2 logits are tied to within fp32 accumulation noise (margin
    below 2e-5 on logits of magnitude 1..10, i.e. a few dozen ulps: measured, 1 image out of 25 000 scored over two seeds,
    oracle margin 5.2e-6; any implementation that is not bit-identical to the CPU BLAS summation order flips such an image)
    and every such image is listed with its margin -, identical accuracy lists, loss trace and classifier weights within 1e-3 (measured: 2e-5..2e-4 and
    2e-4..6e-4), epoch counts equal wherever the reference's stopping rule is decidable: the rule compares fp32 loss
    differences (quantised at 2.4e-7) with 1e-4 while those differences decay by ~5e-8 per epoch, so the oracle's own stop
    epoch is decided by fp32 rounding noise over a stretch of ~25 epochs - an epoch count is accepted only inside the window
    in which a +-2 ulp perturbation of the ORACLE's own loss values would stop (`stop_window`);
  * throughput tier (plain bf16 convolutions): loss / weights within the bf16 bar, and every prediction that differs from
    the oracle is listed with the oracle's own top-1 / top-2 logit margin - a flip is only accepted where that margin is
    smaller than the feature error bf16 operands introduce.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TIE_MARGIN = 2e-5   # oracle top-1 / top-2 logit gap below which a prediction is decided by fp32 accumulation order


def _run(golden, word_embed_dir, precision):
    import contextlib
    import io
    from eval.language_eval import few_shot_finetune_incremental_test
    from models.util import create_model
    from srb200 import synthetic
    world = synthetic.make_world(golden['seed'], n_sessions=golden['n_sessions'], n_base_batch=golden['n_base_batch'],
                                 word_embed_path=word_embed_dir, **golden['overrides'])
    opt = world.opt
    opt.n_sessions_override = golden['n_sessions']
    opt.conv_precision = precision
    net = synthetic.init_model(create_model, opt, world.seed)
    ckpt = synthetic.make_ckpt(net, world)
    net = net.cuda()
    with contextlib.redirect_stdout(io.StringIO()):
        few_shot_finetune_incremental_test(net, ckpt, torch.nn.CrossEntropyLoss(), world.meta_valloader,
                                           world.base_val_loader, opt, base_support_loader=world.base_support_loader)
    return few_shot_finetune_incremental_test.last_record


def stop_window(loss, eps=1e-4, stable_epochs=10, min_epochs=20, max_epochs=1000, ulps=2):
    """Epochs at which the reference's stopping rule (language_eval.py:298-318, --target_train_loss 0 => opt.stable) fires
    on the ORACLE's fp32 loss trace when every |loss_t - loss_{t-1}| < eps comparison is perturbed by +-`ulps` fp32 ulps of
    the loss: -> (earliest, latest) stop epoch; `latest` is mirrored around the oracle's own stop when the strict rule does
    not fire inside the recorded trace."""
    loss = np.asarray(loss, dtype=np.float64)
    tol = ulps * float(np.spacing(np.float32(abs(loss[-1]))))
    prev = np.concatenate([[15.0], loss[:-1]])
    d = np.abs(loss - prev)

    def first_stop(thr):
        run = 0
        for e, v in enumerate(d):
            run = run + 1 if v < thr else 0
            if run == stable_epochs or e + 1 >= max_epochs:
                return e + 1
        return None
    own = len(loss)
    lo = first_stop(eps + tol) or own
    hi = first_stop(eps - tol)
    return lo, (hi if hi is not None else own + (own - lo))


def _compare(rec, g, tag):
    """-> list of per-session dicts with the measured deviations and the prediction mismatches (with oracle margins)."""
    out = []
    for s, (a, b) in enumerate(zip(rec['sessions'], g['sessions'])):
        n = min(a['epochs'], b['epochs'])
        ta, tb = a['terms'][:n, 0], b['terms'][:n, 0].astype(np.float64)
        loss_rel = float(np.max(np.abs(ta - tb) / np.abs(tb)))
        wa, wb = a['W'].cpu(), b['W']
        w_rel = ((wa - wb).norm() / wb.norm()).item()
        pf = ((a['probe_feat'].cpu() - b['probe_feat']).norm() / b['probe_feat'].norm()).item()
        flips = []
        for k, (p, q, m) in enumerate(zip(a['query_pred'], b['query_pred'], b['query_margin'])):
            for i in torch.nonzero(p.long() != q.long()).flatten().tolist():
                flips.append(("query%d" % (k + 1), i, float(m[i])))
        for i in torch.nonzero(a['base_pred'].long() != b['base_pred'].long()).flatten().tolist():
            flips.append(("base", i, float(b['base_margin'][i])))
        n_scored = sum(p.numel() for p in a['query_pred']) + a['base_pred'].numel()
        min_margin = min([float(m.min()) for m in b['query_margin']] + [float(b['base_margin'].min())])
        out.append(dict(epochs=(a['epochs'], b['epochs']), loss_rel=loss_rel, w_rel=w_rel, feat_rel=pf, flips=flips,
                        n_scored=n_scored, min_margin=min_margin, acc=(a['novel_session_acc'], b['novel_session_acc']),
                        acc_base=(a['acc_base'], b['acc_base'])))
        print("%s session %d: epochs %d/%d  loss rel %.2e  W rel %.2e  probe-feature rel %.2e  flips %d/%d  (oracle min margin "
              "%.2e)%s" % (tag, s + 1, a['epochs'], b['epochs'], loss_rel, w_rel, pf, len(flips), n_scored, min_margin,
                           "".join("\n      %s[%d] oracle margin %.3e" % f for f in flips[:12])))
    return out


@pytest.mark.parametrize("seed", [1, 2])
def test_config2_parity_tier_bit_exact_predictions(seed, golden_dir, word_embed_dir):
    path = os.path.join(golden_dir, "config2_seed%d.pt" % seed)
    if not os.path.exists(path):
        pytest.skip("fixture %s not generated" % path)
    g = torch.load(path, weights_only=False)
    rec = _run(g, word_embed_dir, "bf16x3")
    cmp = _compare(rec, g, "bf16x3 seed %d" % seed)
    same_epochs = all(c['epochs'][0] == c['epochs'][1] for c in cmp)
    if same_epochs:
        assert rec['counters'] == g['counters']                      # BasicBlock.num_batches_tracked bookkeeping
    for s, c in enumerate(cmp):
        lo, hi = stop_window(g['sessions'][s]['terms'][:, 0])
        print("session %d: epochs %d, oracle %d, oracle's +-2 ulp stop window [%d, %d]" % (s + 1, c['epochs'][0], c['epochs'][1], lo, hi))
        assert lo <= c['epochs'][0] <= hi, "session %d epochs %s outside [%d, %d]" % (s + 1, c['epochs'], lo, hi)
        assert c['loss_rel'] < 1e-3 and c['w_rel'] < 1e-3, (s + 1, c['loss_rel'], c['w_rel'])   # north_star's bar
        assert c['feat_rel'] < 2e-4, (s + 1, c['feat_rel'])
        hard = [f for f in c['flips'] if f[2] >= TIE_MARGIN]
        assert hard == [], "session %d: predictions differ from the oracle beyond fp32 ties: %s" % (s + 1, hard[:20])
        assert len(c['flips']) <= 2, "session %d: %d near-tie flips" % (s + 1, len(c['flips']))
        if not c['flips']:
            assert c['acc'][0] == c['acc'][1] and abs(c['acc_base'][0] - c['acc_base'][1]) < 1e-9, (s + 1, c['acc'], c['acc_base'])
        else:   # one image of 125 / 1000 changes an accuracy by at most 0.8 / 0.1 points
            assert max(abs(x - y) for x, y in zip(c['acc'][0], c['acc'][1])) <= 0.8 * len(c['flips']) + 1e-9
            assert abs(c['acc_base'][0] - c['acc_base'][1]) <= 0.1 * len(c['flips']) + 1e-9
        for k, v in g['sessions'][s]['bn'].items():
            u = rec['sessions'][s]['bn'][k].cpu()
            if 'num_batches_tracked' in k:
                assert int(u) == int(v), k
            else:
                assert ((u - v).norm() / (v.norm() + 1e-12)).item() < 2e-4, k
    n_flips = sum(len(c['flips']) for c in cmp)
    print("parity tier seed %d: %d of %d predictions differ from the oracle (all below the fp32 tie margin %.0e)" %
          (seed, n_flips, sum(c['n_scored'] for c in cmp), TIE_MARGIN))
    if n_flips == 0:
        assert rec['weighted'] == g['weighted'] and rec['novel'] == g['novel'] and rec['base'] == g['base']
        assert abs(rec['acc_novel_avg'] - g['acc_novel_avg']) < 1e-9 and abs(rec['acc_base_avg'] - g['acc_base_avg']) < 1e-9
    else:
        np.testing.assert_allclose(rec['weighted'], g['weighted'], atol=0.2 * n_flips)
        np.testing.assert_allclose(rec['novel'], g['novel'], atol=0.2 * n_flips)
        np.testing.assert_allclose(rec['base'], g['base'], atol=0.1 * n_flips)


def test_config2_throughput_tier_bf16(golden_dir, word_embed_dir):
    path = os.path.join(golden_dir, "config2_seed1.pt")
    if not os.path.exists(path):
        pytest.skip("fixture %s not generated" % path)
    g = torch.load(path, weights_only=False)
    rec = _run(g, word_embed_dir, "bf16")
    cmp = _compare(rec, g, "bf16 seed 1")
    total_flips = sum(len(c['flips']) for c in cmp)
    total = sum(c['n_scored'] for c in cmp)
    print("bf16 tier: %d of %d predictions differ from the fp32 oracle" % (total_flips, total))
    for s, c in enumerate(cmp):
        # bf16 operands: ~3e-3 feature error -> losses / weights inside north_star's 1e-3 .. few e-3 band
        # (measured on B200: loss 9e-5..2.7e-3, W 3e-3..1.8e-2, features 4e-3..7e-3, 1.0-2.3 % of the predictions flip, every
        # one of them at an oracle margin below 2.3e-2)
        assert c['loss_rel'] < 5e-3, (s + 1, c['loss_rel'])
        assert c['w_rel'] < 3e-2, (s + 1, c['w_rel'])
        assert c['feat_rel'] < 1e-2, (s + 1, c['feat_rel'])
        # a flipped prediction is only acceptable where the oracle itself was nearly tied
        for what, i, margin in c['flips']:
            assert margin < 5e-2, "session %d %s[%d] flipped although the oracle margin is %.3e" % (s + 1, what, i, margin)
        assert len(c['flips']) <= 0.03 * c['n_scored'], (s + 1, len(c['flips']), c['n_scored'])
    for a, b in zip(rec['weighted'], g['weighted']):
        assert abs(a - b) <= 1.0, (rec['weighted'], g['weighted'])
