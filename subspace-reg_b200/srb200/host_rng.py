"""Keep-masks of the train-mode epoch from PyTorch's CPU generator.

The draws MUST be the ones the reference makes (F.dropout -> bernoulli_(1-p); Bernoulli(gamma).sample), in the same
order, because the same generator later initialises the next session's classifier rows.  Two equivalent producers:
torch itself (always correct, ~10 ns per element, serial) and sr_host_bernoulli (a vectorised replay of the same
mt19937 stream from torch.get_rng_state()).  The replay is used only after it has reproduced torch on a sample that
crosses several generator blocks; otherwise every call goes to torch.
"""
import ctypes as C
import threading
import time

import torch

from . import _lib as L
from . import rng

_replay_ok = None
WAIT_S = [0.0, 0.0, 0]   # diagnostics (tools/one_sweep.py): seconds the run's thread waited for pre-drawn masks, seconds it spent
                         # drawing masks itself, number of such in-line draws


def _torch_draw(shape, p, kind):
    g = rng.generator()
    if kind == 0:
        return torch.empty(shape, dtype=torch.uint8).bernoulli_(p, generator=g)
    probs = torch.tensor(p)                       # float32, like torch.distributions.Bernoulli(gamma).probs
    return torch.bernoulli(probs.expand(shape), generator=g).to(torch.uint8)


def _replay_draw(shape, p, kind, out=None):
    n = 1
    for s in shape:
        n *= int(s)
    state = rng.get_state()
    if out is None:
        out = torch.empty(shape, dtype=torch.uint8)
    ones = L.load().sr_host_bernoulli(C.c_void_p(state.data_ptr()), state.numel(), kind, float(p), n,
                                      C.c_void_p(out.data_ptr()))
    if ones < 0:
        raise RuntimeError("srb200: unexpected torch CPU generator state layout")
    rng.set_state(state)
    return out, int(ones)


def replay_into(state_blob, p, kind, out):
    """Draw out.numel() Bernoulli(p) values of `kind` from the generator state `state_blob` (a uint8 tensor like
    torch.get_rng_state(), advanced in place) WITHOUT touching torch's live generator.  -> number of ones, or -1."""
    return int(L.load().sr_host_bernoulli(C.c_void_p(state_blob.data_ptr()), state_blob.numel(), kind, float(p), out.numel(),
                                          C.c_void_p(out.data_ptr())))


def _self_check():
    saved = torch.get_rng_state()
    ok = True
    try:
        for kind, p in ((0, 0.9), (1, 0.0123), (0, 0.5), (1, 0.5)):
            torch.default_generator.manual_seed(1234 + kind)     # CPU generator only (torch.manual_seed reseeds CUDA's too)
            torch.rand(7)                          # start mid-block
            s0 = torch.get_rng_state()
            a = _torch_draw((3, 1777), p, kind)
            sa = torch.get_rng_state()
            torch.set_rng_state(s0)
            b, ones = _replay_draw((3, 1777), p, kind)
            sb = torch.get_rng_state()
            ok = ok and torch.equal(a, b) and torch.equal(sa, sb) and ones == int(a.sum())
    except Exception:
        ok = False
    torch.set_rng_state(saved)
    return ok


def replay_available():
    global _replay_ok
    if _replay_ok is None:
        _replay_ok = _self_check()
    return _replay_ok


def bernoulli_u8(shape, p, kind, out=None):
    """uint8 CPU tensor of draws (1 with probability p) + number of ones, consuming torch's CPU generator exactly like
    kind 0: tensor.bernoulli_(p)   kind 1: torch.bernoulli(torch.tensor(p).expand(shape))."""
    if replay_available():
        t0 = time.perf_counter()
        r = _replay_draw(shape, p, kind, out)
        WAIT_S[1] += time.perf_counter() - t0
        WAIT_S[2] += 1
        return r
    t = _torch_draw(shape, p, kind)
    if out is not None:
        out.copy_(t)
        t = out
    return t, int(t.sum())


def dropblock_keep(seeds, block_size, out):
    """DropBlock._compute_block_mask on host: seeds uint8 [B,C,hs,ws] -> out uint8 [B,C,hs+bs-1,ws+bs-1] (1 = keep);
    returns the number of kept positions."""
    B, Cc, hs, ws = seeds.shape
    assert seeds.is_contiguous() and out.is_contiguous() and out.numel() == B * Cc * (hs + block_size - 1) * (ws + block_size - 1)
    kept = L.load().sr_host_dropblock(C.c_void_p(seeds.data_ptr()), B * Cc, hs, ws, block_size, C.c_void_p(out.data_ptr()))
    if kept < 0:
        raise RuntimeError("srb200: sr_host_dropblock rejected its arguments")
    return int(kept)


def _numel(shape):
    n = 1
    for s in shape:
        n *= int(s)
    return n


class MaskPrefetch(object):
    """Pre-draws, on a host thread, the dropout masks of FUTURE train-mode forwards while the GPU is busy.

    The thread replays torch's stream on a private copy of the generator state taken now.  Consumers whose probability
    is not known yet (DropBlock's gamma depends on how many epochs the running session will take) and the next
    session's nn.Linear init only need to be SKIPPED: they use a fixed number of 32-bit draws.  Every pre-drawn mask
    remembers the generator state before and after it; `take` hands it out only if the live generator is exactly in
    the `before` state (then moves it to `after`), so a wrong guess about what happens in between costs nothing but
    the prefetch.  One instance may cover every remaining session of a run: the thread then works ahead continuously
    and `take` only waits for the mask it asks for."""

    def __init__(self, steps):
        """steps: list of ('skip', n_words) | ('draw', key, uint8 CPU buffer [shape], p) | ('hold', key, n_words): like
        'skip', but the generator states before / after the skipped words are kept under `key` so that somebody else can
        draw those words later, off the live generator (DropBlock seeds: their probability is known only when the
        session starts)."""
        self.ok = replay_available()
        self.results = {}
        self.holds = {}
        self.thread = None
        self.dead = False          # the live stream diverged from the guess (or the replay failed): nothing is usable
        self.finished = False
        self._cv = threading.Condition()
        self._pending = sum(1 for st in steps if st[0] == 'draw')
        if not self.ok:
            self.dead = True
            return
        self._state = rng.get_state().clone()
        self._steps = steps
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def alive(self):
        """True while there are pre-drawn (or still to be drawn) masks that the live stream may yet consume."""
        return self.thread is not None and not self.dead and self._pending > 0

    def _run(self):
        lib = L.load()
        blob = self._state
        try:
            for st in self._steps:
                if self.dead:
                    break
                if st[0] == 'skip':
                    if lib.sr_host_bernoulli(C.c_void_p(blob.data_ptr()), blob.numel(), 2, 0.0, int(st[1]), None) < 0:
                        raise RuntimeError("skip failed")
                elif st[0] == 'hold':
                    before = blob.clone()
                    if lib.sr_host_bernoulli(C.c_void_p(blob.data_ptr()), blob.numel(), 2, 0.0, int(st[2]), None) < 0:
                        raise RuntimeError("skip failed")
                    with self._cv:
                        self.holds[st[1]] = (before, blob.clone())
                        self._cv.notify_all()
                else:
                    _, key, buf, p = st
                    before = blob.clone()
                    ones = lib.sr_host_bernoulli(C.c_void_p(blob.data_ptr()), blob.numel(), 0, float(p), buf.numel(),
                                                 C.c_void_p(buf.data_ptr()))
                    if ones < 0:
                        raise RuntimeError("draw failed")
                    with self._cv:
                        self.results[key] = (before, blob.clone(), buf, int(ones))
                        self._cv.notify_all()
        except Exception:
            self.dead = True
        with self._cv:
            self.finished = True
            self._cv.notify_all()

    def hold_state(self, key):
        """-> (state before, state after) of a held region once the thread has passed it, or None."""
        if self.thread is None or self.dead:
            return None
        with self._cv:
            while key not in self.holds and not self.finished and not self.dead:
                self._cv.wait()
            return self.holds.get(key)

    def take(self, key, shape):
        """-> (uint8 buffer, ones) or None.  Advances torch's live generator exactly as the draw would have."""
        if self.thread is None or self.dead:
            return None
        t0 = time.perf_counter()
        with self._cv:
            while key not in self.results and not self.finished and not self.dead:
                self._cv.wait()
            ent = self.results.pop(key, None)
        WAIT_S[0] += time.perf_counter() - t0
        if ent is None:
            return None
        self._pending -= 1
        before, after, buf, ones = ent
        if tuple(buf.shape) != tuple(shape) or not torch.equal(rng.get_state(), before):
            self.dead = True           # the stream went somewhere else: everything drawn after this point is void too
            self.results = {}
            return None
        rng.set_state(after)
        return buf, ones
