"""Several independent runs (seeds) in flight on ONE GPU: a host thread + CUDA stream + private CPU generators per run.

A sweep alternates phases that cannot fill a B200 on their own - the persistent head kernel is latency-bound on ~85 SMs,
a train-mode pass pushes a few hundred images through ~65 small launches, the host draws masks in between - so the device
is shared by K runs whose phases interleave: while one run's head loop spins on its grid barriers another run's
convolutions take the idle SMs.  Runs share nothing (the reference runs one process per seed: scripts/continual/
slurm_subspace_reg.sh:7-8,25), so there is no data-path synchronisation between the threads; ctypes and PyTorch release
the GIL while they enqueue, and the numbers of a run do not depend on K (srb200.rng binds private generators).
"""
import contextlib
import gc
import queue
import threading

import torch

from . import device_rng
from . import host_rng
from . import ops
from . import rng


@contextlib.contextmanager
def frozen_heap():
    """Keep the long-lived objects of a multi-seed job (worlds, models) out of the cyclic garbage collector while sweeps
    run: every generation-2 collection otherwise re-scans them and stalls the launching thread for 100-400 ms (measured on
    a B200 box with 24 resident worlds: 9 of 24 sweeps hit; 2 of 24 inside this context).  The collector stays enabled for
    the sweeps' own garbage."""
    gc.collect()
    gc.freeze()
    try:
        yield
    finally:
        gc.unfreeze()


def prewarm_allocator(gigabytes, device=None):
    """Make PyTorch's caching allocator reserve `gigabytes` of HBM in ONE segment up front.  A sweep allocates and frees
    multi-hundred-MB activation buffers of session-dependent sizes; served from a fragmented cache they keep falling through
    to cudaMalloc (~9 calls per sweep), and on the B200 boxes used here a cudaMalloc / the copies and allocations queued
    behind it occasionally block the launching thread for 100-300 ms.  With one large cached segment to split, those
    requests are served without touching the driver (measured over 24 sweeps: stalls > 200 ms in 5 sweeps -> 1)."""
    dev = device if device is not None else "cuda"
    x = torch.empty(int(gigabytes * (1 << 30)), dtype=torch.uint8, device=dev)
    del x
    # Requests of <= 1 MB come from a separate pool of 2 MB segments (status blocks, traces, labels, statistics: a few
    # hundred per sweep); reserve some of those too, so that the small pool does not grow inside a sweep either.
    small = [torch.empty(1 << 19, dtype=torch.uint8, device=dev) for _ in range(256)]
    del small


class Lockstep(object):
    """Rendezvous of the n runs of one group (see ops.align_runs): align() returns once all n threads have called it, with
    the calling thread's CUDA stream waiting for everything the other n - 1 streams had queued at that point.  A run that
    fails or finishes early leaves the group with leave(); the others carry on unaligned."""

    def __init__(self, n, record=None, wait=None):
        """record() -> event of the calling thread's stream at this point; wait(event): make the calling thread's stream wait
        for it.  Defaults: CUDA events on the current stream (the CPU tests pass plain callables)."""
        self.n = n
        self.broken = n <= 1
        self._events = [None] * n
        self._barrier = threading.Barrier(n) if n > 1 else None
        self._record = record or self._cuda_record
        self._wait = wait or self._cuda_wait

    @staticmethod
    def _cuda_record():
        ev = torch.cuda.Event()
        ev.record()
        return ev

    @staticmethod
    def _cuda_wait(ev):
        torch.cuda.current_stream().wait_event(ev)

    def align(self, slot=None):
        if self.broken:
            return
        if slot is None:
            slot = getattr(_slot, "i", None)
        self._events[slot] = self._record()
        try:
            self._barrier.wait(timeout=120.0)           # every run has recorded its event
            for i, e in enumerate(self._events):
                if i != slot and e is not None:
                    self._wait(e)
            self._barrier.wait(timeout=120.0)           # every run has read the list (it is reused by the next align)
        except threading.BrokenBarrierError:
            self.broken = True

    def leave(self):
        self.broken = True
        if self._barrier is not None:
            self._barrier.abort()


_slot = threading.local()


class SeedPool(object):
    """K persistent worker threads, each with its own CUDA stream.  map(fn, items) runs fn(item) for every item, at most K
    at a time, and returns the results in order; exceptions are re-raised in the caller.  With lockstep=True (default) the
    items are run in groups of K whose session drivers rendezvous before every head loop (ops.align_runs)."""

    def __init__(self, workers, device=None, lockstep=True, prewarm_gb=0.0, group=None):
        self.workers = max(1, int(workers))
        self.lockstep = bool(lockstep) and self.workers > 1
        # runs per lockstep group (default: all workers form one group).  With workers = 2 x group two groups are in flight:
        # while one group's head loops spin on their grid barriers the other group's convolutions queue behind them.
        self.group = self.workers if not group else max(1, min(int(group), self.workers))
        # PyTorch's caching allocator keeps one pool PER STREAM: a segment reserved on the caller's stream (prewarm_allocator)
        # is of no use to the workers' streams, so each worker reserves its own before its first run
        self.prewarm_gb = float(prewarm_gb)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        host_rng.replay_available()              # the generator self-checks touch the global generator: do them here, once
        device_rng.available(self.device)
        ops.set_gpu_share(self.group)             # head launches of the runs of a group must fit on the device together
        self._q = queue.Queue()
        self._threads = [threading.Thread(target=self._work, daemon=True) for _ in range(self.workers)]
        for t in self._threads:
            t.start()

    def _work(self):
        torch.cuda.set_device(self.device)
        stream = torch.cuda.Stream(device=self.device)
        if self.prewarm_gb > 0:
            with torch.cuda.stream(stream):
                prewarm_allocator(self.prewarm_gb, self.device)
        while True:
            job = self._q.get()
            if job is None:
                return
            fn, item, out, idx, done, group, slot = job
            _slot.i = slot
            ops.bind_lockstep(group)
            try:
                with torch.cuda.stream(stream), rng.scope():
                    out[idx] = (True, fn(item))
                    ops.wait_stream_blocking()
            except BaseException as e:            # noqa: BLE001 - handed to the caller
                out[idx] = (False, e)
            finally:
                ops.bind_lockstep(None)
                if group is not None:
                    group.leave()                 # a finished (or failed) run must not keep the others waiting
            done.release()

    def map(self, fn, items):
        items = list(items)
        out = [None] * len(items)
        done = threading.Semaphore(0)
        torch.cuda.current_stream().synchronize()     # inputs prepared on the caller's stream are complete
        if self.lockstep:
            in_flight = []                                      # sizes of the groups dispatched and not yet finished
            max_groups = max(1, self.workers // self.group)
            for g0 in range(0, len(items), self.group):         # groups of <= `group` runs, at most workers / group at a time
                chunk = items[g0:g0 + self.group]
                if len(in_flight) == max_groups:
                    for _ in range(in_flight.pop(0)):           # (approximation: ANY `size` finished runs make room)
                        done.acquire()
                group = Lockstep(len(chunk))
                for j, it in enumerate(chunk):
                    self._q.put((fn, it, out, g0 + j, done, group, j))
                in_flight.append(len(chunk))
            for n in in_flight:
                for _ in range(n):
                    done.acquire()
        else:
            for i, it in enumerate(items):
                self._q.put((fn, it, out, i, done, None, 0))
            for _ in items:
                done.acquire()
        res = []
        for ok, v in out:
            if not ok:
                raise v
            res.append(v)
        return res

    def close(self):
        for _ in self._threads:
            self._q.put(None)
        for t in self._threads:
            t.join()
        ops.set_gpu_share(1)
