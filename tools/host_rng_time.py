"""Throughput of the host Bernoulli replay (sr_host_bernoulli) for different worker-thread counts."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "subspace-reg_b200"))
import torch  # noqa: E402

from srb200 import host_rng  # noqa: E402

print("replay ok", host_rng.replay_available(), "cpus", os.cpu_count())
shape = (200, 64, 42, 42)
for thr in ("0", "2", "3", "4", "6", "8"):
    os.environ["SRB_RNG_THREADS"] = thr
    torch.manual_seed(0)
    for k in (0, 1):
        best = 1e9
        for rep in range(3):
            t0 = time.perf_counter()
            out = host_rng.bernoulli_u8(shape, 0.9, k)
            best = min(best, time.perf_counter() - t0)
        n = out[0].numel()
        print("threads", thr, "kind", k, "%.1f ms  %.2f ns/elem" % (best * 1e3, best * 1e9 / n))
