"""Torch-tensor front end of the srb200 C ABI.

PyTorch is used here for device memory and streams only: every function checks its tensors (CUDA, contiguous,
dtype), passes raw pointers + the current stream to libsrb200.so and returns torch tensors that own the outputs.
No function in this module has a CPU or PyTorch-op fallback.
"""
import ctypes as C
import threading

import torch

from . import _lib as L

_checked_devices = set()
class _LaunchCounter(object):
    """Number of srb200 kernels enqueued through this module (bench.py's gpu_launches): one slot per host thread, so
    concurrent runs never lose an update; LAUNCHES[0] reads the total, LAUNCHES.add(n) adds to the caller's slot."""

    def __init__(self):
        self._slots = {}

    def __getitem__(self, i):
        return sum(self._slots.values())

    def add(self, n=1):
        k = threading.get_ident()
        self._slots[k] = self._slots.get(k, 0) + n


LAUNCHES = _LaunchCounter()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream():
    # the raw handle of torch's current stream (the private accessor skips building a Stream object: this runs ~100
    # times per train-mode forward, which is host-bound)
    if _raw_stream is not None:
        return C.c_void_p(_raw_stream(torch.cuda.current_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t, dtype=None, name="tensor"):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("srb200: %s must be a CUDA tensor (no CPU fallback)" % name)
    if not t.is_contiguous():
        raise RuntimeError("srb200: %s must be contiguous" % name)
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError("srb200: %s must be %s, got %s" % (name, dtype, t.dtype))
    dev = t.device.index if t.device.index is not None else torch.cuda.current_device()
    if dev not in _checked_devices:
        L.check(L.load().sr_check_device(dev), "sr_check_device")
        _checked_devices.add(dev)
    return C.c_void_p(t.data_ptr())


def _lo(t):
    """Low plane of an error-compensated pair tensor [2, ...] (None for a plain tensor)."""
    return None if t is None or t.dim() not in (5,) else t[1]


def _planes(shape, split, device, dtype=torch.bfloat16):
    """Output buffer: plain [shape] or an error-compensated pair [2, *shape] (plane 0 = hi, plane 1 = lo)."""
    return torch.empty(((2,) + tuple(shape)) if split else tuple(shape), dtype=dtype, device=device)


def pack_input(x_nchw, cpad=16, split=False):
    """NCHW fp32 -> NHWC bf16 with channels zero-padded to cpad.  split: error-compensated pair [2,B,H,W,cpad]."""
    B, Cc, H, W = x_nchw.shape
    y = _planes((B, H, W, cpad), split, x_nchw.device)
    rc = L.load().sr_pack_input(_ptr(x_nchw, torch.float32, "x"), _ptr(y[0] if split else y), _ptr(y[1]) if split else None,
                                B, Cc, H, W, cpad, _stream())
    L.check(rc, "sr_pack_input")
    LAUNCHES.add(1)
    return y


def bn_fold(gamma, beta, running_mean, running_var, eps=1e-5):
    Cc = gamma.numel()
    scale = torch.empty(Cc, dtype=torch.float32, device=gamma.device)
    shift = torch.empty_like(scale)
    rc = L.load().sr_bn_fold(_ptr(gamma, torch.float32), _ptr(beta, torch.float32), _ptr(running_mean, torch.float32),
                             _ptr(running_var, torch.float32), eps, _ptr(scale), _ptr(shift), Cc, _stream())
    L.check(rc, "sr_bn_fold")
    LAUNCHES.add(1)
    return scale, shift


def pack_input_u8(x_nhwc_u8, mean, std, cpad=16, crop_ij=None, flip=None, pad=0, split=False):
    """uint8 NHWC images -> [RandomCrop(padding=pad) at crop_ij, horizontal flip where flip] -> ToTensor +
    Normalize(mean, std) -> NHWC bf16 with channels zero-padded to cpad (one kernel).  crop_ij: CUDA int32 [B,2],
    flip: CUDA uint8 [B] (see dataset.transform_cfg.draw_crop_flip), or None."""
    B, H, W, Cc = x_nhwc_u8.shape
    y = _planes((B, H, W, cpad), split, x_nhwc_u8.device)
    m = (C.c_float * Cc)(*[float(v) for v in mean])
    s = (C.c_float * Cc)(*[float(v) for v in std])
    if crop_ij is not None and tuple(crop_ij.shape) != (B, 2):
        raise RuntimeError("srb200: crop_ij must be [batch, 2]")
    if flip is not None and flip.numel() != B:
        raise RuntimeError("srb200: flip must be [batch]")
    rc = L.load().sr_pack_input_u8(_ptr(x_nhwc_u8, torch.uint8, "x"), _ptr(y[0] if split else y),
                                   _ptr(y[1]) if split else None, B, Cc, H, W, m, s, cpad,
                                   _ptr(crop_ij, torch.int32, "crop_ij"), _ptr(flip, torch.uint8, "flip"), int(pad), _stream())
    L.check(rc, "sr_pack_input_u8")
    LAUNCHES.add(1)
    return y


def pack_weight(w_oihw, scale=None, cin_pad=None, out=None, split=False):
    """OIHW fp32 -> bf16 [cout, kh*kw, cin_pad] (optionally scaled per output channel); split: pair [2, cout, taps, cin_pad]."""
    co, ci, kh, kw = w_oihw.shape
    if cin_pad is None:
        cin_pad = (ci + 15) // 16 * 16
    if out is None:
        out = _planes((co, kh * kw, cin_pad), split, w_oihw.device)
    split = out.dim() == 4
    rc = L.load().sr_pack_weight(_ptr(w_oihw, torch.float32, "w"), _ptr(scale, torch.float32, "scale"),
                                 _ptr(out[0] if split else out), _ptr(out[1]) if split else None, co, ci, kh, kw, cin_pad,
                                 _stream())
    L.check(rc, "sr_pack_weight")
    LAUNCHES.add(1)
    return out


def conv(panels, cout, shift=None, residual=None, slope=0.1, epilogue=L.SR_EPI_ACT, stats=None, out=None):
    """panels: list of (act NHWC bf16 [B,H,W,cin_pad], packed weight bf16 [cout,taps,cin_pad]); or, error-compensated
    mode, of pair tensors (act [2,B,H,W,cin_pad], weight [2,cout,taps,cin_pad]) - then residual / bf16 outputs are pairs too."""
    split = panels[0][0].dim() == 5
    if split:
        if any(act.dim() != 5 or wgt.dim() != 4 for act, wgt in panels) or (residual is not None and residual.dim() != 5):
            raise RuntimeError("srb200: error-compensated conv needs pair tensors for every operand")
        lo_panels = [(act[1], wgt[1]) for act, wgt in panels]
        panels = [(act[0], wgt[0]) for act, wgt in panels]
        residual_lo = None if residual is None else residual[1]
        residual = None if residual is None else residual[0]
    act0 = panels[0][0]
    B, H, W, _ = act0.shape
    a = L.ConvArgs()
    a.batch, a.height, a.width, a.cout = B, H, W, cout
    a.n_panels = len(panels)
    for i, (act, wgt) in enumerate(panels):
        if act.shape[:3] != act0.shape[:3]:
            raise RuntimeError("srb200: conv panels must share batch and spatial size")
        if wgt.shape[0] != cout or wgt.shape[2] != act.shape[3]:
            raise RuntimeError("srb200: packed weight %s does not match activation %s / cout %d" %
                               (tuple(wgt.shape), tuple(act.shape), cout))
        a.panel[i].act = _ptr(act, torch.bfloat16, "act")
        a.panel[i].wgt = _ptr(wgt, torch.bfloat16, "wgt")
        a.panel[i].cin_pad = act.shape[3]
        a.panel[i].taps = wgt.shape[1]
        if split:
            a.panel[i].act_lo = _ptr(lo_panels[i][0], torch.bfloat16, "act_lo")
            a.panel[i].wgt_lo = _ptr(lo_panels[i][1], torch.bfloat16, "wgt_lo")
    a.shift = _ptr(shift, torch.float32, "shift")
    a.residual = _ptr(residual, torch.bfloat16, "residual")
    if split and residual is not None:
        a.residual_lo = _ptr(residual_lo, torch.bfloat16, "residual_lo")
    if residual is not None and tuple(residual.shape) != (B, H, W, cout):
        raise RuntimeError("srb200: residual shape mismatch")
    a.slope = slope
    a.epilogue = epilogue
    dev = act0.device
    if out is None:
        if epilogue == L.SR_EPI_ACT:
            out = _planes((B, H, W, cout), split, dev)
        elif epilogue == L.SR_EPI_ACT_POOL2:
            out = _planes((B, H // 2, W // 2, cout), split, dev)
        elif epilogue == L.SR_EPI_ACT_AVG:
            out = torch.empty((B, cout), dtype=torch.float32, device=dev)
        else:
            out = torch.empty((B, H, W, cout), dtype=torch.float32, device=dev)
    if split and out.dim() == 5:
        a.out, a.out_lo = _ptr(out[0]), _ptr(out[1])
    else:
        a.out = _ptr(out)
    a.stats = _ptr(stats, torch.float64, "stats")
    L.check(L.load().sr_conv(C.byref(a), _stream()), "sr_conv")
    LAUNCHES.add(1)
    return out


def bn_finalize(stats, count, running_mean, running_var, eps=1e-5, momentum=0.1):
    Cc = running_mean.numel()
    mean = torch.empty(Cc, dtype=torch.float32, device=stats.device)
    invstd = torch.empty_like(mean)
    rc = L.load().sr_bn_finalize(_ptr(stats, torch.float64), int(count), eps, momentum,
                                 _ptr(running_mean, torch.float32), _ptr(running_var, torch.float32), _ptr(mean),
                                 _ptr(invstd), Cc, _stream())
    L.check(rc, "sr_bn_finalize")
    LAUNCHES.add(1)
    return mean, invstd


def bn_apply(raw, mean, invstd, gamma, beta, res_raw=None, res_bn=None, res_act=None, lrelu=True, slope=0.1, pool=0,
             keep=None, keep_scale=1.0, split=False):
    """split: bf16 outputs (and res_act) are error-compensated pairs [2, ...]."""
    B, H, W, Cc = raw.shape
    a = L.BnApplyArgs()
    a.batch, a.height, a.width, a.channels = B, H, W, Cc
    a.raw = _ptr(raw, torch.float32, "raw")
    a.mean, a.invstd = _ptr(mean, torch.float32), _ptr(invstd, torch.float32)
    a.gamma, a.beta = _ptr(gamma, torch.float32), _ptr(beta, torch.float32)
    if res_raw is not None:
        a.res_raw = _ptr(res_raw, torch.float32, "res_raw")
        rm, ri, rg, rb = res_bn
        a.res_mean, a.res_invstd = _ptr(rm, torch.float32), _ptr(ri, torch.float32)
        a.res_gamma, a.res_beta = _ptr(rg, torch.float32), _ptr(rb, torch.float32)
    if res_act is not None and res_act.dim() == 5:
        a.res_act, a.res_act_lo = _ptr(res_act[0], torch.bfloat16, "res_act"), _ptr(res_act[1], torch.bfloat16, "res_act_lo")
    else:
        a.res_act = _ptr(res_act, torch.bfloat16, "res_act")
    a.lrelu = 1 if lrelu else 0
    a.slope = slope
    a.pool = pool
    a.keep = _ptr(keep, torch.uint8, "keep")
    if torch.is_tensor(keep_scale):          # device scalar (DropBlock's numel / kept from sr_dropblock_keep)
        a.keep_scale = 1.0
        a.keep_scale_dev = _ptr(keep_scale, torch.float32, "keep_scale")
    else:
        a.keep_scale = keep_scale
    if pool == -1:
        out = torch.empty((B, Cc), dtype=torch.float32, device=raw.device)
    elif pool == 2:
        out = _planes((B, H // 2, W // 2, Cc), split, raw.device)
    else:
        out = _planes((B, H, W, Cc), split, raw.device)
    if split and out.dim() == 5:
        a.out, a.out_lo = _ptr(out[0]), _ptr(out[1])
    else:
        a.out = _ptr(out)
    L.check(L.load().sr_bn_apply(C.byref(a), _stream()), "sr_bn_apply")
    LAUNCHES.add(1)
    return out


def backbone_eval(h, blocks, folded, slope=0.1):
    """The whole eval-mode backbone on a packed input in ONE library call (sr_backbone_eval): h NHWC bf16 [B,H,W,cin_pad]
    (or a pair [2,...]), blocks = the engine's block descriptors, folded = their folded-BN packed weights / shifts.
    -> fp32 [B, cout_last]."""
    split = h.dim() == 5
    hs = h[0] if split else h
    B, H, W, cin_pad = hs.shape
    dev = hs.device
    arr = (L.EvalBlock * len(blocks))()
    for i, (b, w) in enumerate(zip(blocks, folded)):
        e = arr[i]
        e.cout, e.pool, e.downsample = b['cout'], b['pool'], 1 if b['downsample'] else 0
        for k in ('w1', 'w2', 'w3') + (('wd',) if b['downsample'] else ()):
            t = w[k]
            setattr(e, k, _ptr(t[0] if split else t, torch.bfloat16, k).value)
            if split:
                setattr(e, k + '_lo', _ptr(t[1], torch.bfloat16, k + '_lo').value)
        e.s1, e.s2, e.s3 = (_ptr(w[k], torch.float32, k).value for k in ('s1', 's2', 's3'))
    a = L.BackboneEvalArgs()
    a.n_blocks, a.blocks = len(blocks), arr
    a.batch, a.height, a.width, a.cin_pad = B, H, W, cin_pad
    a.x = _ptr(hs, torch.bfloat16, "x")
    a.x_lo = _ptr(h[1], torch.bfloat16, "x_lo") if split else None
    a.slope = slope
    ws_bytes = int(L.load().sr_backbone_eval_workspace_bytes(C.byref(a)))
    ws = scratch('backbone_eval', (ws_bytes,), torch.uint8, dev)
    a.workspace, a.workspace_bytes = _ptr(ws), ws_bytes
    feat = torch.empty((B, blocks[-1]['cout']), dtype=torch.float32, device=dev)
    a.features = _ptr(feat)
    L.check(L.load().sr_backbone_eval(C.byref(a), _stream()), "sr_backbone_eval")
    LAUNCHES.add(3 * len(blocks) + (1 if blocks[-1]['pool'] == 2 else 0))
    return feat


_SCRATCH = {}


def scratch(tag, shape, dtype, device):
    """Grow-only device scratch buffer per (host thread, tag): the large temporaries of the train-mode pass are handed out
    from here instead of being allocated and freed on every call.  Their sizes change from forward to forward (support vs
    memory batches), which keeps a per-call allocation falling through PyTorch's cache to cudaMalloc - and on the GPU boxes
    used here an allocation that reaches the driver occasionally blocks the launching thread for 100-300 ms.  Stream-ordered
    reuse: a buffer is only ever written and read by launches on the calling thread's stream."""
    n = 1
    for v in shape:
        n *= int(v)
    nbytes = n * torch.empty((), dtype=dtype).element_size()
    key = (threading.get_ident(), tag, str(device))
    buf = _SCRATCH.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _SCRATCH[key] = buf
    return buf[:nbytes].view(dtype).view(shape)


def train_block(x, w, bn, stats, cin_pad, cout, downsample, eps, momentum, slope):
    """One BasicBlock of the train-mode pass up to its last BatchNorm in ONE library call (sr_train_block): conv1 -> BN ->
    LeakyReLU -> conv2 -> BN -> LeakyReLU -> conv3 [-> 1x1 downsample conv of x], batch statistics and running-stat EMAs
    included.  x: NHWC bf16 [B,H,W,cin_pad] or a pair [2,...]; w: [w1, w2, w3(, wd)] packed raw weights (pairs in the x3
    tier); bn: [(gamma, beta, running_mean, running_var)] per conv; stats: zeroed fp64 [n_convs * 2 * cout].
    -> (raw3, rawd | None, mean_invstd [n_convs, 2, cout])."""
    split = x.dim() == 5
    xs = x[0] if split else x
    B, H, W, _ = xs.shape
    dev = xs.device
    n_convs = 4 if downsample else 3
    a = L.TrainBlockArgs()
    a.batch, a.height, a.width, a.cin_pad, a.cout, a.downsample = B, H, W, cin_pad, cout, 1 if downsample else 0
    a.x = _ptr(xs, torch.bfloat16, "x")
    a.x_lo = _ptr(x[1], torch.bfloat16, "x_lo") if split else None
    for i in range(n_convs):
        wi = w[i]
        a.w[i] = _ptr(wi[0] if split else wi, torch.bfloat16, "w").value
        a.w_lo[i] = _ptr(wi[1], torch.bfloat16, "w_lo").value if split else None
        g, b_, rm, rv = bn[i]
        if i < 2:
            a.gamma[i] = _ptr(g, torch.float32, "gamma").value
            a.beta[i] = _ptr(b_, torch.float32, "beta").value
        a.running_mean[i] = _ptr(rm, torch.float32, "running_mean").value
        a.running_var[i] = _ptr(rv, torch.float32, "running_var").value
    a.eps, a.momentum, a.slope = eps, momentum, slope
    a.stats = _ptr(stats, torch.float64, "stats")
    mi = torch.empty((n_convs, 2, cout), dtype=torch.float32, device=dev)
    a.mean_invstd = _ptr(mi)
    raw12 = scratch('tb_raw12', (B, H, W, cout), torch.float32, dev)            # raw output of conv1, then of conv2
    raw3 = scratch('tb_raw3', (B, H, W, cout), torch.float32, dev)
    rawd = scratch('tb_rawd', (B, H, W, cout), torch.float32, dev) if downsample else None
    a.raw[0] = a.raw[1] = _ptr(raw12).value
    a.raw[2] = _ptr(raw3).value
    a.raw[3] = _ptr(rawd).value if downsample else None
    h = scratch('tb_h', ((2, 2) if split else (2,)) + (B, H, W, cout), torch.bfloat16, dev)   # h1 | h2 (pairs in the x3 tier)
    if split:
        a.h1, a.h1_lo, a.h2, a.h2_lo = _ptr(h[0, 0]), _ptr(h[1, 0]), _ptr(h[0, 1]), _ptr(h[1, 1])
    else:
        a.h1, a.h2 = _ptr(h[0]), _ptr(h[1])
    L.check(L.load().sr_train_block(C.byref(a), _stream()), "sr_train_block")
    LAUNCHES.add(2 * n_convs + 2)        # n convs + n finalizes + two BN-apply launches
    return raw3, rawd, mi


def subspace_factor(base_weight):
    """Orthonormal row basis Qt of span(rows of base_weight) -> (qt [q_rows, dim], q_rows, is_identity).

    Replaces torch.qr(base_weight.T) of LangPuller.get_projected_weight (reference resnet_language.py:92-97)."""
    n, d = base_weight.shape
    lib = L.load()
    ws_bytes = lib.sr_subspace_factor_workspace_bytes(n, d)
    ws = torch.empty(max(int(ws_bytes), 256), dtype=torch.uint8, device=base_weight.device)
    q = min(n, d)
    qt = torch.zeros((q, d), dtype=torch.float32, device=base_weight.device)
    info = torch.zeros(4, dtype=torch.int32, device=base_weight.device)
    rc = lib.sr_subspace_factor(_ptr(base_weight, torch.float32, "base_weight"), n, d, _ptr(qt), _ptr(info), _ptr(ws),
                                ws.numel(), _stream())
    L.check(rc, "sr_subspace_factor")
    LAUNCHES.add(4)
    info_h = info.cpu().tolist()
    if info_h[2] != 0:
        raise RuntimeError("srb200: base weights are rank deficient (Cholesky pivot %d <= 0)" % (info_h[2] - 1))
    return qt, info_h[0], bool(info_h[1])


_GPU_SHARE = [1]
_share_tls = threading.local()


def set_gpu_share(runs):
    """Tell the ops how many independent runs share this process's GPU (srb200.concurrent.SeedPool does).  With more than
    one, the paper-size head kernel is capped at SMs // runs CTAs (sr_head_args.cta_budget) so that the cooperative
    launches of all runs are resident together instead of queueing behind one another."""
    _GPU_SHARE[0] = max(1, int(runs))


def head_cta_budget():
    k = _GPU_SHARE[0]
    ls = getattr(_share_tls, "lockstep", None)
    if ls is not None:
        k = min(k, ls.n)      # a trailing group of fewer runs (one run: the device is its own again)
    if k <= 1:
        return 0
    return torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count // k


def bind_lockstep(ls):
    """Bind (or clear, with None) the rendezvous object of the calling thread's run (srb200.concurrent.Lockstep)."""
    _share_tls.lockstep = ls


def align_runs():
    """Called by the session driver right before it queues a head loop.  With several runs in flight on this GPU, wait
    until every run of the group has reached the same point and make this run's stream wait for the others' queued work:
    the K budgeted cooperative head launches then start together and are co-resident (3 x 47 CTAs), while the
    convolution phases - persistent kernels that want all 148 SMs - never run beside a head loop.  Free-running groups lose
    what the budget buys: a conv kernel next to ONE head loop runs its 148 CTAs in two waves on 101 SMs (measured: 365 ms
    per sweep with 3 free-running sweeps in flight, the same as one sweep alone).  No-op for a run that is alone."""
    ls = getattr(_share_tls, "lockstep", None)
    if ls is not None:
        ls.align()


def wait_stream_blocking():
    """Sleep (not spin) until everything queued on the current stream is done.  A plain .cpu() / synchronize() busy-waits on
    the launching thread - one host core per run for most of a sweep, which the mask generator threads of 8 ranks need."""
    ev = torch.cuda.Event(blocking=True)
    ev.record()
    ev.synchronize()


class HeadSession(object):
    """Device state of one session's fine-tuning (weight, optimiser state, counters) around sr_head_run."""

    def __init__(self, feat, n_support, support_row0, labels_support, weight, n_base, n_new, *, n_memory=0,
                 memory_row0=0, labels_memory=None, base_weight=None, reserve_weight=None, pull_mode=L.SR_PULL_NONE,
                 pull=None, q_rows=0, lmbd_base=0.0, lmbd_novel=0.0, gamma=0.0, adam=False, lr=0.002, momentum=0.9,
                 weight_decay=5e-4, stable=True, convergence_epsilon=1e-4, stable_epochs=10, target_train_loss=0.0,
                 min_novel_epochs=20, max_novel_epochs=1000, want_logits=False, cta_budget=None):
        dev = feat.device
        self.feat, self.weight = feat, weight
        self.labels_support, self.labels_memory = labels_support, labels_memory
        self.base_weight, self.reserve_weight, self.pull = base_weight, reserve_weight, pull
        Cn, d = weight.shape
        self.opt_state = torch.zeros((2 if adam else 1, Cn, d), dtype=torch.float32, device=dev)
        self._pending = []       # [(status, trace)] of launches whose results have not been read back yet
        self._posted = None      # (pinned status, pinned traces, event, #launches) queued by post()
        self.logits = torch.empty((n_support, Cn), dtype=torch.float32, device=dev) if want_logits else None
        a = L.HeadArgs()
        a.feat, a.dim = _ptr(feat, torch.float32, "feat"), d
        a.n_support, a.support_row0 = n_support, support_row0
        a.n_memory, a.memory_row0 = n_memory, memory_row0
        a.labels_support = _ptr(labels_support, torch.int64, "labels_support")
        a.labels_memory = _ptr(labels_memory, torch.int64, "labels_memory")
        a.weight, a.n_classes = _ptr(weight, torch.float32, "weight"), Cn
        a.opt_state = _ptr(self.opt_state)
        a.base_weight, a.n_base = _ptr(base_weight, torch.float32, "base_weight"), n_base
        a.reserve_weight = _ptr(reserve_weight, torch.float32, "reserve_weight")
        a.n_prev_novel = 0 if reserve_weight is None else reserve_weight.shape[0]
        a.n_new, a.pull_mode, a.pull, a.q_rows = n_new, pull_mode, _ptr(pull, torch.float32, "pull"), q_rows
        a.lmbd_base, a.lmbd_novel, a.gamma = lmbd_base, lmbd_novel, gamma
        a.optimizer = L.SR_OPT_ADAM if adam else L.SR_OPT_SGD
        a.lr, a.momentum, a.weight_decay = lr, momentum, weight_decay
        a.beta1, a.beta2, a.adam_eps = 0.9, 0.999, 1e-8
        a.stable, a.stable_epochs = (1 if stable else 0), stable_epochs
        a.min_novel_epochs, a.max_novel_epochs = min_novel_epochs, max_novel_epochs
        a.convergence_epsilon, a.target_train_loss = convergence_epsilon, target_train_loss
        a.logits_support = _ptr(self.logits)
        a.cta_budget = head_cta_budget() if cta_budget is None else int(cta_budget)
        ws_bytes = int(L.load().sr_head_workspace_bytes(C.byref(a)))
        self.workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        a.workspace, a.workspace_bytes = _ptr(self.workspace), ws_bytes
        self.args = a
        self.epochs = 0
        self.stable_count = 0
        self.prev_loss = 15.0  # language_eval.py:234
        self.stopped = False
        self.traces = []

    def run(self, max_epochs, feat=None, support_row0=None, memory_row0=None, defer=False):
        """Run up to max_epochs more epochs on the device; returns the [epochs, SR_TRACE_COLS] trace (host).

        `feat` / row offsets may change between calls (epoch 1 uses the train-mode features, later epochs the
        eval-mode cache); labels, weight and optimiser state carry over.

        defer=True enqueues the launch and returns None WITHOUT reading anything back: the next run() is chained to it on
        the device (sr_head_args.resume_status), and collect() reads every pending status / trace in one go.  A whole
        session (epoch 1, cache build, remaining epochs, scoring) is then queued without a host synchronisation."""
        a = self.args
        if feat is not None:
            self.feat = feat
            a.feat = _ptr(feat, torch.float32, "feat")
            if feat.shape[1] != a.dim:
                raise RuntimeError("srb200: feature width changed")
        if support_row0 is not None:
            a.support_row0 = support_row0
        if memory_row0 is not None:
            a.memory_row0 = memory_row0
        dev = self.feat.device
        trace = torch.zeros((max_epochs, L.SR_TRACE_COLS), dtype=torch.float32, device=dev)
        status = torch.zeros(8, dtype=torch.int32, device=dev)
        a.loss_trace = _ptr(trace)
        a.status = _ptr(status)
        a.resume_status = _ptr(self._pending[-1][0]) if self._pending else None
        a.max_epochs, a.epoch0, a.step0 = max_epochs, self.epochs, self.epochs
        a.stable_count0, a.prev_loss = self.stable_count, self.prev_loss
        L.check(L.load().sr_head_run(C.byref(a), _stream()), "sr_head_run")
        LAUNCHES.add(1)
        self._pending.append((status, trace))
        if defer:
            return None
        return self.collect()

    def post(self):
        """Queue the device->host copies of every pending launch's status / trace NOW (pinned buffers, one event), so that
        collect() waits for the head loop only - not for whatever the caller queues on the stream after this point (the
        session driver queues the query cache build and the scoring behind the head loop and reads the epoch count while
        they run)."""
        if not self._pending or self._posted is not None:
            return
        sts = torch.stack([st for st, _ in self._pending])
        sts_h = torch.empty(sts.shape, dtype=sts.dtype, pin_memory=True)
        sts_h.copy_(sts, non_blocking=True)
        traces_h = []
        for _, trace in self._pending:
            th = torch.empty(trace.shape, dtype=trace.dtype, pin_memory=True)
            th.copy_(trace, non_blocking=True)
            traces_h.append(th)
        ev = torch.cuda.Event(blocking=True)
        ev.record()
        self._posted = (sts_h, traces_h, ev, len(self._pending))

    def collect(self):
        """Read back every pending launch (the one host sync): -> concatenated trace of the epochs they ran."""
        if not self._pending:
            return torch.zeros((0, L.SR_TRACE_COLS), dtype=torch.float32)
        if self._posted is not None and self._posted[3] != len(self._pending):
            self._posted = None                      # launches were queued after post(): read everything the plain way
        if self._posted is not None:
            sts_h, traces_h, ev, _ = self._posted
            ev.synchronize()
            sts = sts_h.tolist()
            traces = traces_h
        else:
            wait_stream_blocking()
            sts = torch.stack([st for st, _ in self._pending]).cpu().tolist()
            traces = [trace for _, trace in self._pending]
        self._posted = None
        out = []
        for st, trace in zip(sts, traces):
            if st[3] != 0:
                raise RuntimeError("srb200: head kernel reported a grid-barrier timeout")
            n = st[0]
            self.epochs += n
            self.stopped = bool(st[1])
            self.stable_count = st[2]
            if n > 0:
                tr = trace[:n].cpu().clone() if trace.is_cuda else trace[:n].clone()
                self.prev_loss = float(tr[-1, 0])
                self.traces.append(tr)
                out.append(tr)
        self._pending = []
        return torch.cat(out, 0) if out else torch.zeros((0, L.SR_TRACE_COLS), dtype=torch.float32)

    def run_to_convergence(self, chunk=1 << 20):
        while not self.stopped:
            self.run(min(chunk, max(self.args.max_novel_epochs - self.epochs, 1)))
        return torch.cat(self.traces, 0)


def eval_logits(feat, weight, labels, confusion=None):
    """-> dict(logits, pred, top1, top5, loss_sum) for rows of feat scored against weight."""
    n, d = feat.shape
    Cn = weight.shape[0]
    dev = feat.device
    logits = torch.empty((n, Cn), dtype=torch.float32, device=dev)
    pred = torch.empty(n, dtype=torch.int32, device=dev)
    counts = torch.zeros(2, dtype=torch.int32, device=dev)
    loss_sum = torch.zeros(1, dtype=torch.float32, device=dev)
    a = L.EvalArgs()
    a.feat, a.weight = _ptr(feat, torch.float32, "feat"), _ptr(weight, torch.float32, "weight")
    a.labels = _ptr(labels, torch.int64, "labels")
    a.n, a.dim, a.n_classes = n, d, Cn
    a.logits, a.pred, a.counts, a.loss_sum = _ptr(logits), _ptr(pred), _ptr(counts), _ptr(loss_sum)
    a.confusion = _ptr(confusion, torch.int64, "confusion")
    a.conf_dim = 0 if confusion is None else confusion.shape[0]
    ws_bytes = int(L.load().sr_eval_workspace_bytes(n, d, Cn))     # > 0: large problem, logits GEMM on tcgen05
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev) if ws_bytes > 0 else None
    a.workspace, a.workspace_bytes = _ptr(ws), ws_bytes
    L.check(L.load().sr_eval_logits(C.byref(a), _stream()), "sr_eval_logits")
    LAUNCHES.add(2 if ws is None else 4)
    return {"logits": logits, "pred": pred, "counts": counts, "loss_sum": loss_sum}


# ---------------------------------------------------------------------------------------------------------------
# per-op surface (used by srb200.autograd behind the drop-in modules)
# ---------------------------------------------------------------------------------------------------------------
def semantic_pullers(novel_embeds, base_embeds, base_weight, temperature, mask=False):
    n, e = novel_embeds.shape
    b, d = base_weight.shape
    if base_embeds.shape != (b, e):
        raise RuntimeError("srb200: base_embeds %s does not match base_weight %s" % (tuple(base_embeds.shape), tuple(base_weight.shape)))
    out = torch.empty((n, d), dtype=torch.float32, device=base_weight.device)
    rc = L.load().sr_semantic_pullers(_ptr(novel_embeds, torch.float32, "novel_embeds"), _ptr(base_embeds, torch.float32, "base_embeds"),
                                      _ptr(base_weight, torch.float32, "base_weight"), n, b, e, d, temperature, 1 if mask else 0,
                                      _ptr(out), _stream())
    L.check(rc, "sr_semantic_pullers")
    LAUNCHES.add(1)
    return out


def linear_fwd(x, w, bias=None):
    n, k = x.shape
    m = w.shape[0]
    y = torch.empty((n, m), dtype=torch.float32, device=x.device)
    rc = L.load().sr_linear_fwd(_ptr(x, torch.float32, "x"), _ptr(w, torch.float32, "w"), _ptr(bias, torch.float32, "bias"), n, k, m,
                                _ptr(y), _stream())
    L.check(rc, "sr_linear_fwd")
    LAUNCHES.add(1)
    return y


def linear_bwd(dy, x, want_bias):
    n, m = dy.shape
    k = x.shape[1]
    dw = torch.empty((m, k), dtype=torch.float32, device=x.device)
    db = torch.empty(m, dtype=torch.float32, device=x.device) if want_bias else None
    rc = L.load().sr_linear_bwd(_ptr(dy, torch.float32, "dy"), _ptr(x, torch.float32, "x"), n, k, m, _ptr(dw), _ptr(db), _stream())
    L.check(rc, "sr_linear_bwd")
    LAUNCHES.add(1)
    return dw, db


def sqdist(a, b):
    out = torch.empty(1, dtype=torch.float32, device=a.device)
    rc = L.load().sr_sqdist(_ptr(a, torch.float32, "a"), _ptr(b, torch.float32, "b"), a.numel(), _ptr(out), _stream())
    L.check(rc, "sr_sqdist")
    LAUNCHES.add(1)
    return out


def diff_scale(a, b, scale, gout=None, sq=None):
    out = torch.empty_like(a)
    rc = L.load().sr_diff_scale(_ptr(a, torch.float32, "a"), _ptr(b, torch.float32, "b"), a.numel(), scale,
                                _ptr(gout, torch.float32, "gout"), _ptr(sq, torch.float32, "sq"), _ptr(out), _stream())
    L.check(rc, "sr_diff_scale")
    LAUNCHES.add(1)
    return out


def project_rows(x, qt, q_rows):
    n, d = x.shape
    out = torch.empty_like(x)
    rc = L.load().sr_project_rows(_ptr(x, torch.float32, "x"), _ptr(qt, torch.float32, "qt"), n, q_rows, d, _ptr(out), _stream())
    L.check(rc, "sr_project_rows")
    LAUNCHES.add(1)
    return out


def score_logits(logits, labels, confusion=None):
    """Top-1 / top-5 hit counts, argmax and summed CE of logits the caller already has."""
    n, Cn = logits.shape
    dev = logits.device
    pred = torch.empty(n, dtype=torch.int32, device=dev)
    counts = torch.zeros(2, dtype=torch.int32, device=dev)
    loss_sum = torch.zeros(1, dtype=torch.float32, device=dev)
    a = L.EvalArgs()
    a.labels = _ptr(labels, torch.int64, "labels")
    a.n, a.dim, a.n_classes = n, 0, Cn
    a.logits, a.pred, a.counts, a.loss_sum = _ptr(logits, torch.float32, "logits"), _ptr(pred), _ptr(counts), _ptr(loss_sum)
    a.confusion = _ptr(confusion, torch.int64, "confusion")
    a.conf_dim = 0 if confusion is None else confusion.shape[0]
    L.check(L.load().sr_score_logits(C.byref(a), _stream()), "sr_score_logits")
    LAUNCHES.add(1)
    return {"pred": pred, "counts": counts, "loss_sum": loss_sum}


def global_avg(x_nhwc):
    """bf16 NHWC [B,H,W,C] (or an error-compensated pair [2,B,H,W,C]) -> fp32 [B,C] (AdaptiveAvgPool2d(1))."""
    lo = _lo(x_nhwc)
    if lo is not None:
        x_nhwc = x_nhwc[0]
    B, H, W, Cc = x_nhwc.shape
    y = torch.empty((B, Cc), dtype=torch.float32, device=x_nhwc.device)
    rc = L.load().sr_global_avg(_ptr(x_nhwc, torch.bfloat16, "x"), _ptr(lo, torch.bfloat16, "x_lo"), _ptr(y), B, H, W, Cc, _stream())
    L.check(rc, "sr_global_avg")
    LAUNCHES.add(1)
    return y
