"""Keep-masks of the train-mode epoch drawn on the GPU from PyTorch's CPU generator state (csrc/mask.cu, csrc/mt_jump.cpp).

The draws are the reference's (F.dropout -> bernoulli_(1-p); Bernoulli(gamma).sample), in the same order and from the same
mt19937 stream, because the same generator later initialises the next session's classifier rows: sr_device_bernoulli reads
the generator's state on the host, produces the words on the device (jump-ahead walkers) and moves the host generator past
them.  Nothing but the 2.5 KB state crosses PCIe - the host path (srb200.host_rng) shipped 43 MB of masks per forward and
kept two to four host cores per run busy.  Used only after it has reproduced torch itself on a sample (`available`).
"""
import ctypes as C
import os
import threading

import torch

from . import _lib as L
from . import rng

JUMP_WORDS = 1 << 18
_lock = threading.Lock()
_host_table = None          # int32 [n, 624] (CPU), grows
_dev_tables = {}            # device index -> CUDA copy of (a prefix of) the host table
_ok = {}                    # device index -> self-check verdict


def _tables(n_polys, device):
    """-> (host table, device table) holding at least n_polys jump polynomials."""
    global _host_table
    with _lock:
        have = 0 if _host_table is None else _host_table.shape[0]
        if have < n_polys:
            n = max(n_polys, 2 * have, 64)
            t = torch.zeros((n, 624), dtype=torch.int32)
            L.check(L.load().sr_mt_jump_table(C.c_void_p(t.data_ptr()), n), "sr_mt_jump_table")
            _host_table = t
        idx = device.index if device.index is not None else torch.cuda.current_device()
        d = _dev_tables.get(idx)
        if d is None or d.shape[0] < _host_table.shape[0]:
            d = _host_table.to(device)
            torch.cuda.current_stream(device).synchronize()
            _dev_tables[idx] = d
        return _host_table, d


def region_words(kind, n):
    return 2 * n if kind == 0 else n


def draw(regions, workspace=None):
    """regions: [(kind, p, out)] in draw order, out = CUDA uint8 tensor (contiguous) whose numel is the number of elements
    (kind 2: out = number of words to skip).  Queues the launches on the current stream and advances this thread's CPU
    generator (srb200.rng) exactly as the draws would have.  -> the workspace tensor used (reusable)."""
    lib = L.load()
    arr = (L.MaskRegion * len(regions))()
    total = 0
    dev = None
    for i, (kind, p, out) in enumerate(regions):
        arr[i].kind = kind
        arr[i].p = float(p)
        if kind == 2:
            arr[i].n = int(out)
            arr[i].out = None
        else:
            if not (out.is_cuda and out.dtype == torch.uint8 and out.is_contiguous()):
                raise RuntimeError("srb200: mask outputs must be contiguous CUDA uint8 tensors")
            arr[i].n = out.numel()
            arr[i].out = out.data_ptr()
            dev = out.device
        total += region_words(kind, arr[i].n)
    if dev is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    n_polys = (total + JUMP_WORDS - 1) // JUMP_WORDS
    host_t, dev_t = _tables(n_polys, dev)
    need = int(lib.sr_device_bernoulli_workspace_bytes(total))
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(max(need, 1 << 20), dtype=torch.uint8, device=dev)
    state = rng.get_state()
    rc = lib.sr_device_bernoulli(C.c_void_p(state.data_ptr()), state.numel(), arr, len(regions), C.c_void_p(dev_t.data_ptr()),
                                 C.c_void_p(host_t.data_ptr()), host_t.shape[0], C.c_void_p(workspace.data_ptr()),
                                 workspace.numel(), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    L.check(rc, "sr_device_bernoulli")
    rng.set_state(state)
    return workspace


def dropblock_keep(seeds, block_size, keep, scale):
    """DropBlock._compute_block_mask on the device: seeds uint8 [B,C,hs,ws] -> keep uint8 [B,C,hs+bs-1,ws+bs-1] (1 = keep),
    scale[0] = numel / kept (fp32, as the reference computes it); scale: CUDA float32 [4]."""
    B, Cc, hs, ws = seeds.shape
    rc = L.load().sr_dropblock_keep(C.c_void_p(seeds.data_ptr()), B * Cc, hs, ws, block_size, C.c_void_p(keep.data_ptr()),
                                    C.c_void_p(scale.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    L.check(rc, "sr_dropblock_keep")


def _self_check(device):
    """The device path against torch's own CPU draws, on a sample that crosses a jump boundary and starts mid-block."""
    saved = torch.get_rng_state()
    ok = True
    try:
        with torch.cuda.device(device):
            torch.default_generator.manual_seed(4321)     # the CPU generator only: torch.manual_seed would reseed CUDA's too
            torch.rand(5)
            s0 = torch.get_rng_state()
            n0, n1 = 300000, 5001            # 600 k words: crosses two jump boundaries
            a0 = torch.empty(n0, dtype=torch.uint8).bernoulli_(0.9)
            a1 = torch.bernoulli(torch.tensor(0.0123).expand(n1)).to(torch.uint8)
            sa = torch.get_rng_state()
            torch.set_rng_state(s0)
            b0 = torch.empty(n0, dtype=torch.uint8, device=device)
            b1 = torch.empty(n1 + 1, dtype=torch.uint8, device=device)[:n1]
            gen = rng.generator()
            if gen is not torch.default_generator:       # (called from a thread with private generators: check on a copy)
                raise RuntimeError("self-check must run on the global generator")
            draw([(0, 0.9, b0), (1, 0.0123, b1)])
            sb = torch.get_rng_state()
            ok = bool(torch.equal(a0, b0.cpu()) and torch.equal(a1, b1.cpu()) and torch.equal(sa, sb))
    except Exception:
        ok = False
    torch.set_rng_state(saved)
    return ok


def available(device):
    """True if masks may be drawn on `device` (SRB_MASKS=host forces the host path; a failed self-check disables it)."""
    if os.environ.get("SRB_MASKS", "device") == "host":
        return False
    idx = device.index if device.index is not None else torch.cuda.current_device()
    with _lock:
        v = _ok.get(idx)
    if v is None:
        if rng.private():
            # worker threads bind private generators; the check uses the global one, which they must not touch
            return False
        v = _self_check(device)
        with _lock:
            _ok[idx] = v
    return v
