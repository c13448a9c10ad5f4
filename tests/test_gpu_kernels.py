"""Kernel-level parity on the GPU through the C ABI (srb200.ops -> libsrb200.so): each case compares one kernel with a
plain fp32 / fp64 PyTorch statement of the same op on the same device.  Tolerances: bf16-output convolutions 2^-8 of the
output range (one bf16 rounding), fp32-output paths 2e-5, fp32 head / projection 1e-5 (north_star)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


@pytest.mark.parametrize("group", ["conv_basic", "conv_k", "conv_epi", "conv_train", "conv_x3", "head", "factor"])
def test_kernel_group(group):
    import diag
    assert getattr(diag, "group_" + group)()


def test_head_cluster_variant(monkeypatch):
    """The paper-size head has a second implementation that splits both reductions of an epoch over 8-CTA thread-block
    clusters with a DSMEM reduction (head_cluster.cu, opt-in with SRB_HEAD_CLUSTER=1: on par with head_small.cu, not faster).
    Same cases as the `head` group, same 1e-5 bar against fp64 autograd."""
    import diag
    monkeypatch.setenv("SRB_HEAD_CLUSTER", "1")
    assert diag.group_head()


def test_head_stress_config5():
    """BASELINE config 5 shapes: 1000 base + 100 novel classes, 100-shot, 512-d.  With 1000 base rows in 512-d the span
    is everything: the projector is the identity and the regulariser vanishes; n_base = 256 is the
    non-trivial projection."""
    import diag
    from srb200 import _lib as L
    assert diag.head_case("config5 P=I", 10000, 0, 1000, 0, 100, 512, L.SR_PULL_PROJECT, epochs=3)
    assert diag.head_case("config5 nb256", 10000, 0, 256, 0, 100, 512, L.SR_PULL_PROJECT, epochs=3)


@pytest.mark.parametrize("adam", [False, True])
def test_head_tensor_core_path_matches_simt_and_honours_the_stopping_rule(adam):
    """The tcgen05 head (csrc/head_tc.cu: both GEMMs as error-compensated 1x1 convolutions) against the fp32 SIMT
    head_kernel on the same large problem - memory rows, previous-novel anchors, fixed pullers, SGD and Adam - including
    the device-side stopping rule: max_novel_epochs = 4 inside a 10-epoch call stops after exactly 4 updates, and a
    chained second call does nothing.  Loss terms and (SGD) weights at 1e-5; under Adam a weight whose gradient is ~0 moves
    by +-lr on the sign of rounding noise (measured 3.7e-3 of max |W| after 4 steps), so only the losses are held tight."""
    import os
    from srb200 import ops, _lib as L
    dev = "cuda"
    res = {}
    for mode in ("tc", "simt"):
        if mode == "simt":
            os.environ["SRB_HEAD_SIMT"] = "1"
        else:
            os.environ.pop("SRB_HEAD_SIMT", None)
        try:
            g = torch.Generator(device=dev).manual_seed(5)
            Ns, Nm, nb, npv, nn_, d = 9000, 1000, 1000, 50, 50, 512
            Cn = nb + npv + nn_
            feat = (torch.randn(Ns + Nm, d, device=dev, generator=g) * 0.45 + 0.53).clamp_min(-0.11)
            ys = torch.randint(0, Cn, (Ns,), device=dev, generator=g)
            ym = torch.randint(0, Cn, (Nm,), device=dev, generator=g)
            W = (torch.rand(Cn, d, device=dev, generator=g) * 2 - 1) / d ** 0.5
            base = W[:nb].clone() + 0.01 * torch.randn(nb, d, device=dev, generator=g)
            reserve = W[nb:nb + npv].clone() + 0.01 * torch.randn(npv, d, device=dev, generator=g)
            pull = torch.randn(nn_, d, device=dev, generator=g) / d ** 0.5
            hs = ops.HeadSession(feat, Ns, 0, ys, W, nb, nn_, n_memory=Nm, memory_row0=Ns, labels_memory=ym, base_weight=base,
                                 reserve_weight=reserve, pull_mode=L.SR_PULL_FIXED, pull=pull, lmbd_base=0.2, lmbd_novel=0.1,
                                 gamma=1.0, adam=adam, lr=0.002, stable=False, target_train_loss=-1.0, min_novel_epochs=0,
                                 max_novel_epochs=4)
            hs.run(10, defer=True)
            hs.run(10, defer=True)            # chained on the device: must find the rule already fired
            tr = hs.collect()
            res[mode] = (hs.epochs, hs.stopped, tr.clone(), W.clone())
        finally:
            os.environ.pop("SRB_HEAD_SIMT", None)
    (e1, s1, t1, w1), (e2, s2, t2, w2) = res["tc"], res["simt"]
    assert e1 == e2 == 4 and s1 and s2
    rel_t = ((t1[:, :6] - t2[:, :6]).abs().max() / t2[:, :6].abs().max()).item()
    rel_w = ((w1 - w2).abs().max() / w2.abs().max()).item()
    print("tensor-core head vs SIMT head: loss terms rel %.2e, W rel %.2e, hits equal %s" %
          (rel_t, rel_w, bool((t1[:, 6:] == t2[:, 6:]).all())))
    assert rel_t < 1e-5 and rel_w < (1e-2 if adam else 1e-5)


def test_eval_logits_properties():
    """Scoring: predictions equal torch.argmax, hit counts / CE equal the reference's accuracy() + CrossEntropyLoss, the
    confusion matrix sums to n and its trace equals the top-1 hits."""
    from srb200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    n, d, Cn = 16384, 512, 1100
    X = torch.randn(n, d, device="cuda", generator=g)
    W = torch.randn(Cn, d, device="cuda", generator=g) / d ** 0.5
    y = torch.randint(0, Cn, (n,), device="cuda", generator=g)
    conf = torch.zeros(Cn, Cn, dtype=torch.int64, device="cuda")
    r = ops.eval_logits(X, W, y, conf)
    Z = X.double() @ W.double().t()
    assert (r["logits"].double() - Z).abs().max().item() < 1e-4
    Zf = r["logits"]
    assert (r["pred"].long() == Zf.argmax(1)).all()
    top5 = Zf.topk(5, 1).indices
    assert int(r["counts"][0]) == int((Zf.argmax(1) == y).sum())
    assert int(r["counts"][1]) == int((top5 == y[:, None]).any(1).sum())
    ce = torch.nn.functional.cross_entropy(Zf.double(), y, reduction="sum").item()
    assert abs(r["loss_sum"].item() - ce) < 1e-4 * ce
    assert int(conf.sum()) == n and int(conf.diag().sum()) == int(r["counts"][0])
    # the same call on the fp32 SIMT tiles (SRB_HEAD_SIMT=1): the tensor-core logits (error-compensated bf16x3) agree to 1e-5
    import os
    os.environ["SRB_HEAD_SIMT"] = "1"
    try:
        r2 = ops.eval_logits(X, W, y)
    finally:
        os.environ.pop("SRB_HEAD_SIMT", None)
    rel = ((r["logits"] - r2["logits"]).abs().max() / r2["logits"].abs().max()).item()
    print("eval logits, tensor-core vs SIMT: rel %.2e; predictions equal: %s" % (rel, bool((r["pred"] == r2["pred"]).all())))
    assert rel < 1e-5
    for fn, name in ((lambda: ops.eval_logits(X, W, y), "tensor-core"),):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print("query scoring 16384 x 1100 x 512 (%s): %.1f us per call" % (name, e0.elapsed_time(e1) * 1e3 / 5))


def test_backbone_linearity_in_last_residual():
    """Size-independent property at full batch size: eval features are deterministic and independent of batch
    composition (per-image independence of the eval-mode pass that the feature cache relies on)."""
    from models.util import create_model
    from srb200 import synthetic
    net = synthetic.init_model(create_model, synthetic.default_opt(1), 1).cuda().eval()
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(300, 3, 84, 84, device="cuda", generator=g)
    with torch.no_grad():
        f_all = net.features(x)
        f_part = net.features(x[37:150].contiguous())
        f_again = net.features(x)
    assert torch.equal(f_all, f_again)
    assert torch.equal(f_all[37:150], f_part)


@pytest.mark.parametrize("model,precision,tol", [("resnet18", "bf16", 5e-3), ("resnet12", "bf16", 5e-3),
                                                 ("resnet18", "bf16x3", 2.5e-4), ("resnet12", "bf16x3", 2.5e-4)])
def test_backbone_features_and_taps_vs_oracle(model, precision, tol):
    """Eval-mode features of both models in model_pool against the fp32 oracle, in both precision tiers (plain bf16
    tensor-core convolutions: rel-l2 below 5e-3, measured 2.9e-3; error-compensated bf16x3: below 2.5e-4, measured 1.1e-4), and the
    is_feat=True surface: [f0, f1, f2, f3, feat] with the reference's shapes."""
    from models.util import create_model
    from oracle import backbone as obb, init as oinit
    from srb200 import synthetic
    opt = synthetic.default_opt(1, model=model)
    net = synthetic.init_model(create_model, opt, 1).cuda().eval()
    net.set_conv_precision(precision)
    sd = oinit.init_state_dict(1, model=model)
    x = synthetic.make_world(1, n_sessions=1, n_base_batch=12).base_val_loader.batches[0][0]
    plan = obb.block_plan(model, True)
    with torch.no_grad():
        want = obb.features(sd, plan, x, False, obb.new_counters(plan))
        feats, logits = net(x.cuda(), is_feat=True)
        got = net.features(x.cuda())
    rel = ((got.cpu() - want).norm() / want.norm()).item()
    print("%s %s: feature rel-l2 vs fp32 oracle %.3e" % (model, precision, rel))
    assert rel < tol, rel
    assert [tuple(f.shape[1:]) for f in feats] == [(64, 42, 42), (160, 21, 21), (320, 10, 10), (640, 5, 5), (640,)]
    assert ((feats[-1].cpu() - want).norm() / want.norm()).item() < tol
    assert logits.shape == (12, 60)


@pytest.mark.gpu
def test_uint8_image_front_end_is_bit_exact():
    """sr_pack_input_u8 (uint8 HWC store -> ToTensor -> Normalize -> NHWC bf16, one kernel) against the reference's
    preprocessing done by torch on the CPU (x / 255, then (x - mean) / std, transform_cfg.py:8-10, 42-45) followed by
    sr_pack_input: identical bits, hence identical features.  Ragged batch, every byte value present."""
    import torch
    from dataset import transform_cfg
    from models.util import create_model
    from srb200 import ops, synthetic
    g = torch.Generator().manual_seed(7)
    x8 = torch.randint(0, 256, (37, 84, 84, 3), dtype=torch.uint8, generator=g)
    x8[0, :2].view(-1)[:256] = torch.arange(256).to(torch.uint8)
    t = x8.permute(0, 3, 1, 2).to(torch.float32).div(255)                       # ToTensor
    mean = torch.tensor(transform_cfg.mean, dtype=torch.float32).view(1, 3, 1, 1)
    std = torch.tensor(transform_cfg.std, dtype=torch.float32).view(1, 3, 1, 1)
    t = t.sub(mean).div(std).contiguous()                                       # Normalize
    want = ops.pack_input(t.cuda(), 16)
    got = ops.pack_input_u8(x8.cuda(), transform_cfg.mean, transform_cfg.std, 16)
    assert torch.equal(want.view(torch.int16), got.view(torch.int16))
    net = synthetic.init_model(create_model, synthetic.default_opt(1), 1).cuda().eval()
    with torch.no_grad():
        f_ref = net.features(t.cuda())
        f_u8 = net.features(x8.cuda())
    assert torch.equal(f_ref, f_u8)


@pytest.mark.gpu
def test_support_augmentation_on_device_matches_torchvision():
    """RandomCrop(84, padding=8) + RandomHorizontalFlip + ToTensor + Normalize of the reference's support transform
    (transform_cfg.py:32-40), applied by torchvision / PIL on the CPU, against the fused kernel fed with parameters that
    dataset.transform_cfg.draw_crop_flip draws from the same generator: identical bits, identical generator state."""
    import torch
    from dataset import transform_cfg
    from srb200 import ops
    support_tf = transform_cfg.transforms_test_options['A'][0]
    if support_tf is None:
        pytest.skip("torchvision / PIL not installed")
    g = torch.Generator().manual_seed(3)
    x8 = torch.randint(0, 256, (40, 84, 84, 3), dtype=torch.uint8, generator=g)
    torch.manual_seed(123)
    want_f = torch.stack([support_tf(img.numpy()) for img in x8])
    state_ref = torch.get_rng_state()
    torch.manual_seed(123)
    ij, flip = transform_cfg.draw_crop_flip(40)
    assert torch.equal(torch.get_rng_state(), state_ref)
    assert int(flip.sum()) not in (0, 40) and int(ij.min()) >= 0 and int(ij.max()) <= 16
    want = ops.pack_input(want_f.cuda().contiguous(), 16)
    got = ops.pack_input_u8(x8.cuda(), transform_cfg.mean, transform_cfg.std, 16, crop_ij=ij.cuda(), flip=flip.cuda(), pad=8)
    assert torch.equal(want.view(torch.int16), got.view(torch.int16))


def test_head_cta_budget_is_bit_identical():
    """sr_head_args.cta_budget (fewer, fatter CTAs so that several runs' cooperative head launches are co-resident) changes
    the launch shape only: loss trace and weights equal the unconstrained kernel's bit for bit, for SGD and Adam."""
    from srb200 import ops, _lib as L
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(5)
    for adam in (False, True):
        Ns, Nm, nb, npv, nn_, d = 185, 175, 60, 35, 5, 640
        Cn = nb + npv + nn_
        feat = (torch.randn(Ns + Nm, d, device=dev, generator=g) * 0.45 + 0.53).clamp_min(-0.11)
        ys = torch.randint(0, Cn, (Ns,), device=dev, generator=g)
        ym = torch.randint(0, Cn, (Nm,), device=dev, generator=g)
        W = (torch.rand(Cn, d, device=dev, generator=g) * 2 - 1) / d ** 0.5
        base, reserve = W[:nb].clone(), W[nb:nb + npv].clone()
        qt, q, _ = ops.subspace_factor(base.contiguous())
        got = []
        for budget in (0, 74, 49):
            hs = ops.HeadSession(feat, Ns, 0, ys, W.clone(), nb, nn_, n_memory=Nm, memory_row0=Ns, labels_memory=ym,
                                 base_weight=base, reserve_weight=reserve, pull_mode=L.SR_PULL_PROJECT, pull=qt, q_rows=q,
                                 lmbd_base=0.2, lmbd_novel=0.1, gamma=1.0, adam=adam, stable=False, target_train_loss=-1.0,
                                 min_novel_epochs=0, max_novel_epochs=10 ** 6, cta_budget=budget)
            tr = hs.run(200)
            got.append((tr.clone(), hs.weight.clone()))
        for tr, w in got[1:]:
            assert torch.equal(tr, got[0][0]) and torch.equal(w, got[0][1])

