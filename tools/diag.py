"""GPU bring-up diagnostics: each group runs in its own subprocess (a device-side trap poisons the CUDA context)
and prints max errors of the srb200 kernels against plain fp32/fp64 PyTorch on the same device.

    python tools/diag.py            # all groups
    python tools/diag.py conv_basic # one group, in-process
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "subspace-reg_b200"))

GROUPS = ["conv_basic", "conv_k", "conv_epi", "conv_train", "conv_x3", "head", "factor"]


def _ref_conv(act_nhwc, w_packed, taps):
    import torch
    import torch.nn.functional as F
    x = act_nhwc.float().permute(0, 3, 1, 2).contiguous()
    co, t, ci = w_packed.shape
    k = 3 if t == 9 else 1
    w = w_packed.float().reshape(co, k, k, ci).permute(0, 3, 1, 2).contiguous()
    return F.conv2d(x, w, padding=k // 2)  # NCHW fp32


def _report(name, got, want, tol):
    import torch
    got = got.float()
    want = want.float()
    err = (got - want).abs().max().item()
    scale = want.abs().max().item() + 1e-12
    bad = not (err <= tol * scale) or not torch.isfinite(got).all().item()
    print("%-58s max|err| %.3e  (ref max %.3e)  rel %.3e  %s" % (name, err, scale, err / scale, "FAIL" if bad else "ok"),
          flush=True)
    return not bad


def conv_case(B, H, cin, cout, taps, epi, residual=False, second=None, seed=0):
    import torch
    import torch.nn.functional as F
    from srb200 import ops, _lib as L
    g = torch.Generator(device="cuda").manual_seed(seed)
    dev = "cuda"
    cin_pad = (cin + 15) // 16 * 16
    act = torch.zeros(B, H, H, cin_pad, device=dev)
    act[..., :cin] = torch.randn(B, H, H, cin, device=dev, generator=g)
    act = act.to(torch.bfloat16)
    w = torch.zeros(cout, taps, cin_pad, device=dev)
    w[..., :cin] = torch.randn(cout, taps, cin, device=dev, generator=g) / (taps * cin) ** 0.5
    w = w.to(torch.bfloat16)
    shift = torch.randn(cout, device=dev, generator=g) * 0.1
    panels = [(act, w)]
    ref = _ref_conv(act, w, taps)
    if second is not None:
        cin2 = second
        cin2_pad = (cin2 + 15) // 16 * 16
        act2 = torch.zeros(B, H, H, cin2_pad, device=dev)
        act2[..., :cin2] = torch.randn(B, H, H, cin2, device=dev, generator=g)
        act2 = act2.to(torch.bfloat16)
        w2 = torch.zeros(cout, 1, cin2_pad, device=dev)
        w2[..., :cin2] = torch.randn(cout, 1, cin2, device=dev, generator=g) / cin2 ** 0.5
        w2 = w2.to(torch.bfloat16)
        panels.append((act2, w2))
        ref = ref + _ref_conv(act2, w2, 1)
    res = None
    if residual:
        res = torch.randn(B, H, H, cout, device=dev, generator=g).to(torch.bfloat16)
    name = "conv B%d %dx%d cin%d cout%d taps%d epi%d%s%s" % (B, H, H, cin, cout, taps, epi, " +res" if residual else "",
                                                            " +ds%d" % second if second else "")
    if epi == L.SR_EPI_RAW_STATS:
        stats = torch.zeros(2 * cout, dtype=torch.float64, device=dev)
        out = ops.conv(panels, cout, epilogue=epi, stats=stats)
        torch.cuda.synchronize()
        ok = _report(name + " raw", out.permute(0, 3, 1, 2), ref, 2e-5)
        s1 = ref.double().sum((0, 2, 3))
        s2 = (ref.double() ** 2).sum((0, 2, 3))
        ok &= _report(name + " sum", stats[:cout], s1, 1e-4 * max(1.0, (s2.max() ** 0.5 / (s1.abs().max() + 1e-9)).item()))
        ok &= _report(name + " sumsq", stats[cout:], s2, 1e-5)
        return ok
    y = ref + shift.view(1, -1, 1, 1)
    if residual:
        y = y + res.float().permute(0, 3, 1, 2)
    y = F.leaky_relu(y, 0.1)
    out = ops.conv(panels, cout, shift=shift, residual=res, epilogue=epi)
    torch.cuda.synchronize()
    if epi == L.SR_EPI_ACT:
        return _report(name, out.permute(0, 3, 1, 2), y, 6e-3)
    if epi == L.SR_EPI_ACT_POOL2:
        return _report(name, out.permute(0, 3, 1, 2), F.max_pool2d(y, 2), 6e-3)
    return _report(name, out, y.mean((2, 3)), 1e-5)



def _split(x):
    """fp32 -> error-compensated bf16 pair [2, ...] (hi = rn(x), lo = rn(x - hi))."""
    import torch
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return torch.stack([hi, lo], 0).contiguous()


def conv_case_x3(B, H, cin, cout, taps, epi, residual=False, second=None, seed=0, tol=2e-5):
    """Error-compensated (bf16x3) mode against an fp64 convolution of the fp32 operands."""
    import torch
    import torch.nn.functional as F
    from srb200 import ops, _lib as L
    g = torch.Generator(device="cuda").manual_seed(seed)
    dev = "cuda"

    def operand(cin_, taps_):
        cp = (cin_ + 15) // 16 * 16
        a = torch.zeros(B, H, H, cp, device=dev)
        a[..., :cin_] = torch.randn(B, H, H, cin_, device=dev, generator=g)
        w = torch.zeros(cout, taps_, cp, device=dev)
        w[..., :cin_] = torch.randn(cout, taps_, cin_, device=dev, generator=g) / (taps_ * cin_) ** 0.5
        k = 3 if taps_ == 9 else 1
        ref = F.conv2d(a.double().permute(0, 3, 1, 2), w.double().reshape(cout, k, k, cp).permute(0, 3, 1, 2), padding=k // 2)
        return _split(a), _split(w), ref
    a1, w1, ref = operand(cin, taps)
    panels = [(a1, w1)]
    if second is not None:
        a2, w2, r2 = operand(second, 1)
        panels.append((a2, w2))
        ref = ref + r2
    shift = torch.randn(cout, device=dev, generator=g) * 0.1
    res = res_pair = None
    if residual:
        res = torch.randn(B, H, H, cout, device=dev, generator=g)
        res_pair = _split(res)
    name = "conv x3 B%d %dx%d cin%d cout%d taps%d epi%d%s%s" % (B, H, H, cin, cout, taps, epi, " +res" if residual else "",
                                                               " +ds%d" % second if second else "")
    if epi == L.SR_EPI_RAW_STATS:
        stats = torch.zeros(2 * cout, dtype=torch.float64, device=dev)
        out = ops.conv(panels, cout, epilogue=epi, stats=stats)
        torch.cuda.synchronize()
        ok = _report(name + " raw", out.permute(0, 3, 1, 2), ref, tol)
        ok &= _report(name + " sumsq", stats[cout:], (ref ** 2).sum((0, 2, 3)), tol)
        return ok
    y = ref + shift.double().view(1, -1, 1, 1)
    if residual:
        y = y + (res_pair[0].double() + res_pair[1].double()).permute(0, 3, 1, 2)
    y = F.leaky_relu(y, 0.1)
    out = ops.conv(panels, cout, shift=shift, residual=res_pair, epilogue=epi)
    torch.cuda.synchronize()
    if epi == L.SR_EPI_ACT_AVG:
        return _report(name, out, y.mean((2, 3)), tol)
    got = (out[0].double() + out[1].double()).permute(0, 3, 1, 2)
    want = y if epi == L.SR_EPI_ACT else F.max_pool2d(y, 2)
    # the pair itself must be well formed: lo is at most half a bf16 ulp of hi
    ok = bool((out[1].float().abs() <= out[0].float().abs() * 2.0 ** -8 + 1e-30).all().item())
    if not ok:
        print(name + ": lo plane exceeds half an ulp of hi  FAIL", flush=True)
    return _report(name, got, want, tol) and ok


def group_conv_x3():
    from srb200 import _lib as L
    ok = True
    ok &= conv_case_x3(2, 84, 3, 64, 9, L.SR_EPI_ACT)                       # first layer (32-byte rows, tap reuse)
    ok &= conv_case_x3(2, 84, 64, 64, 9, L.SR_EPI_ACT_POOL2, second=3)      # layer1 conv3 + downsample: 6 internal panels
    ok &= conv_case_x3(2, 42, 160, 160, 9, L.SR_EPI_ACT)                    # ragged channel blocks
    ok &= conv_case_x3(7, 21, 320, 320, 9, L.SR_EPI_ACT_POOL2, second=160)  # floor pool 21 -> 10, N split
    ok &= conv_case_x3(13, 10, 320, 320, 9, L.SR_EPI_ACT, residual=True)    # identity residual pair
    ok &= conv_case_x3(23, 5, 640, 640, 9, L.SR_EPI_ACT_AVG, residual=True) # fused global average
    ok &= conv_case_x3(3, 42, 64, 160, 9, L.SR_EPI_RAW_STATS)               # train-mode raw output + statistics
    return ok

def group_conv_basic():
    from srb200 import _lib as L
    ok = True
    ok &= conv_case(2, 42, 64, 64, 1, L.SR_EPI_ACT)       # plain GEMM, SW128, one k-block
    ok &= conv_case(2, 42, 64, 64, 9, L.SR_EPI_ACT)       # 3x3 halo via TMA OOB
    ok &= conv_case(3, 84, 64, 64, 9, L.SR_EPI_ACT)       # two w-tiles
    ok &= conv_case(2, 42, 64, 160, 9, L.SR_EPI_ACT)      # N = 160
    ok &= conv_case(1, 84, 64, 64, 9, L.SR_EPI_ACT)       # a single image: most CTAs of the persistent grid idle
    return ok


def group_conv_k():
    from srb200 import _lib as L
    ok = True
    ok &= conv_case(2, 42, 160, 160, 9, L.SR_EPI_ACT)     # ragged channel blocks 64 + 64 + 32 (zero K steps skipped)
    ok &= conv_case(2, 84, 3, 64, 9, L.SR_EPI_ACT)        # KC = 16 (SW32), tall-box tap reuse
    ok &= conv_case(5, 21, 160, 320, 9, L.SR_EPI_ACT)     # box (21,3,2), N split 2
    ok &= conv_case(13, 10, 320, 640, 9, L.SR_EPI_ACT)    # box (10,2,6), ragged batch, N split 4
    ok &= conv_case(7, 10, 640, 640, 9, L.SR_EPI_ACT)     # 10 k-blocks / tap
    ok &= conv_case(11, 5, 640, 640, 9, L.SR_EPI_ACT, residual=True)
    return ok


def group_conv_epi():
    from srb200 import _lib as L
    ok = True
    ok &= conv_case(2, 84, 64, 64, 9, L.SR_EPI_ACT_POOL2, second=3)      # layer1 conv3 + downsample panel (KC16)
    ok &= conv_case(2, 42, 160, 160, 9, L.SR_EPI_ACT_POOL2, second=64)   # layer2
    ok &= conv_case(7, 21, 320, 320, 9, L.SR_EPI_ACT_POOL2, second=160)  # layer3.0 (floor pool 21 -> 10)
    ok &= conv_case(13, 10, 320, 320, 9, L.SR_EPI_ACT, residual=True)    # layer3.1
    ok &= conv_case(13, 10, 640, 640, 9, L.SR_EPI_ACT_POOL2, second=320) # layer4.0
    ok &= conv_case(23, 5, 640, 640, 9, L.SR_EPI_ACT_AVG, residual=True) # layer4.1 + avgpool
    return ok


def group_conv_train():
    from srb200 import _lib as L
    ok = True
    ok &= conv_case(3, 42, 64, 160, 9, L.SR_EPI_RAW_STATS)
    ok &= conv_case(2, 84, 64, 64, 9, L.SR_EPI_RAW_STATS)   # tall-box tap reuse + raw output / statistics
    ok &= conv_case(2, 84, 3, 64, 9, L.SR_EPI_RAW_STATS)
    ok &= conv_case(5, 21, 160, 320, 1, L.SR_EPI_RAW_STATS)
    ok &= conv_case(7, 5, 640, 640, 9, L.SR_EPI_RAW_STATS)
    return ok


def torch_head_reference(feat, ys, feat_m, ym, W, base, reserve, n_base, n_new, pull_mode, pull, lmbd_b, lmbd_n, gamma,
                         lr, mom, wd, epochs, adam=False):
    """fp64 autograd restatement of language_eval.py:252-295 on cached features."""
    import torch
    from srb200 import _lib as L
    W = W.double().clone().requires_grad_(True)
    opt = (torch.optim.Adam([W], lr=lr, weight_decay=5e-4) if adam else
           torch.optim.SGD([W], lr=lr, momentum=mom, weight_decay=wd))
    ce = torch.nn.CrossEntropyLoss()
    losses = []
    Cn = W.shape[0]
    for _ in range(epochs):
        loss = ce(feat.double() @ W.t(), ys)
        if feat_m is not None:
            loss = loss + ce(feat_m.double() @ W.t(), ym)
        if base is not None:
            loss = loss + lmbd_b * torch.norm(W[:n_base] - base.double())
        if reserve is not None:
            loss = loss + lmbd_n * torch.norm(W[n_base:n_base + reserve.shape[0]] - reserve.double())
        wn = W[Cn - n_new:]
        if pull_mode == L.SR_PULL_PROJECT:
            Q, _ = torch.linalg.qr(base.double().t())
            p = (wn @ Q) @ Q.t()
            loss = loss + gamma * torch.norm(p - wn) ** 2
        elif pull_mode == L.SR_PULL_FIXED:
            loss = loss + gamma * torch.norm(pull.double() - wn) ** 2
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    return W.detach(), losses


def head_case(name, Ns, Nm, n_base, n_prev, n_new, d, pull_mode, epochs=25, adam=False, seed=0):
    import torch
    from srb200 import ops, _lib as L
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(seed)
    Cn = n_base + n_prev + n_new
    feat = (torch.randn(Ns + Nm, d, device=dev, generator=g) * 0.45 + 0.53).clamp_min(-0.11)
    ys = torch.randint(0, Cn, (Ns,), device=dev, generator=g)
    ym = torch.randint(0, Cn, (Nm,), device=dev, generator=g) if Nm else None
    W = (torch.rand(Cn, d, device=dev, generator=g) * 2 - 1) / d ** 0.5
    base = W[:n_base].clone() + 0.01 * torch.randn(n_base, d, device=dev, generator=g)
    reserve = (W[n_base:n_base + n_prev].clone() + 0.01 * torch.randn(n_prev, d, device=dev, generator=g)) if n_prev else None
    pull = None
    q_rows = 0
    if pull_mode == L.SR_PULL_PROJECT:
        pull, q_rows, _ = ops.subspace_factor(base.contiguous())
    elif pull_mode == L.SR_PULL_FIXED:
        pull = torch.randn(n_new, d, device=dev, generator=g) / d ** 0.5
    Wd = W.clone()
    hs = ops.HeadSession(feat, Ns, 0, ys, Wd, n_base, n_new, n_memory=Nm, memory_row0=Ns, labels_memory=ym,
                         base_weight=base, reserve_weight=reserve, pull_mode=pull_mode, pull=pull, q_rows=q_rows,
                         lmbd_base=0.2, lmbd_novel=0.1, gamma=1.0, adam=adam, lr=0.002, momentum=0.9, weight_decay=5e-4,
                         stable=False, target_train_loss=-1.0, min_novel_epochs=0, max_novel_epochs=10 ** 6)
    t0 = time.time()
    tr = hs.run(epochs)
    torch.cuda.synchronize()
    dt = time.time() - t0
    Wref, losses = torch_head_reference(feat[:Ns], ys, feat[Ns:] if Nm else None, ym, W, base, reserve, n_base, n_new,
                                        pull_mode, pull, 0.2, 0.1, 1.0, 0.002, 0.9, 5e-4, epochs, adam=adam)
    ok = _report("head %s loss trace (%d epochs, %.1f ms)" % (name, tr.shape[0], dt * 1e3), tr[:, 0].cuda(),
                 torch.tensor(losses, device=dev), 1e-5)
    ok &= _report("head %s final W" % name, Wd, Wref, 1e-5)
    return ok


def group_head():
    from srb200 import _lib as L
    ok = True
    ok &= head_case("s1 project", 185, 0, 60, 0, 5, 640, L.SR_PULL_PROJECT)
    ok &= head_case("s3 project +M", 185, 50, 60, 10, 5, 640, L.SR_PULL_PROJECT)
    ok &= head_case("s8 fixed +M", 185, 175, 60, 35, 5, 640, L.SR_PULL_FIXED)
    ok &= head_case("s2 adam", 185, 25, 60, 5, 5, 640, L.SR_PULL_PROJECT, adam=True)
    ok &= head_case("stress-ish", 4000, 0, 256, 0, 100, 512, L.SR_PULL_PROJECT, epochs=5)
    return ok


def group_factor():
    import torch
    from srb200 import ops
    ok = True
    for (n, d) in [(60, 640), (5, 64), (256, 512)]:
        g = torch.Generator(device="cuda").manual_seed(n)
        Bm = torch.randn(n, d, device="cuda", generator=g) / d ** 0.5
        qt, q, ident = ops.subspace_factor(Bm)
        Q, _ = torch.linalg.qr(Bm.double().t())
        P_ref = Q @ Q.t()
        P = qt.double().t() @ qt.double()
        ok &= _report("factor n%d d%d projector" % (n, d), P, P_ref, 1e-6)
        ok &= _report("factor n%d d%d orthonormality" % (n, d), qt.double() @ qt.double().t(),
                      torch.eye(q, device="cuda", dtype=torch.float64), 1e-6)
    qt, q, ident = ops.subspace_factor(torch.randn(100, 64, device="cuda"))
    print("factor n>=d -> identity flag:", ident, q)
    return ok and ident and q == 64


def main():
    if len(sys.argv) > 1:
        import torch
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        ok = globals()["group_" + sys.argv[1]]()
        print("GROUP %s: %s" % (sys.argv[1], "PASS" if ok else "FAIL"), flush=True)
        sys.exit(0 if ok else 1)
    rc = 0
    for grp in GROUPS:
        print("==== %s ====" % grp, flush=True)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), grp], timeout=420)
            rc |= r.returncode != 0
        except subprocess.TimeoutExpired:
            print("GROUP %s: TIMEOUT" % grp, flush=True)
            rc = 1
    sys.exit(rc)


if __name__ == "__main__":
    main()
