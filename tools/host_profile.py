"""Host cost of one config-2 sweep: CPU time of the launching thread vs wall time, cProfile of one sweep, and a device
timeline (CUDA events against a common origin) of K sweeps in flight.  GPU."""
import cProfile
import io
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "subspace-reg_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from srb200 import synthetic  # noqa: E402
from srb200.concurrent import SeedPool, frozen_heap, prewarm_allocator  # noqa: E402

K = int(os.environ.get("K", "3"))
wdir = bench.word_embed_dir()
prewarm_allocator(24)


def mk(seed):
    return bench.prepare(bench.place_world(
        synthetic.make_world(seed, n_sessions=8, n_base_batch=1000, word_embed_path=wdir), 'gpu'))


for it in range(3):
    bench.run_sweeps([mk(100 + it)], None)
torch.cuda.synchronize()
with frozen_heap():
    for it in range(3):
        prep = mk(10 + it)
        torch.cuda.synchronize()
        w0, c0, p0 = time.perf_counter(), time.thread_time(), time.process_time()
        bench.run_sweeps([prep], None)
        torch.cuda.synchronize()
        w1, c1, p1 = time.perf_counter(), time.thread_time(), time.process_time()
        print("solo sweep %d: wall %.0f ms, launching-thread CPU %.0f ms, process CPU (all threads) %.0f ms" %
              (it, (w1 - w0) * 1e3, (c1 - c0) * 1e3, (p1 - p0) * 1e3), flush=True)
    prep = mk(20)
    pr = cProfile.Profile()
    pr.enable()
    bench.run_sweeps([prep], None)
    torch.cuda.synchronize()
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(45)
    print(s.getvalue(), flush=True)

    if K > 1:
        pool = SeedPool(K)
        bench.run_sweeps([mk(200 + i) for i in range(K)], pool)
        for rep in range(2):
            preps = [mk(300 + 10 * rep + i) for i in range(2 * K)]
            torch.cuda.synchronize()
            w0, p0 = time.perf_counter(), time.process_time()
            recs = bench.run_sweeps(preps, pool)
            torch.cuda.synchronize()
            w1, p1 = time.perf_counter(), time.process_time()
            print("%d sweeps, %d in flight: wall %.0f ms = %.0f ms per sweep; process CPU %.0f ms" %
                  (2 * K, K, (w1 - w0) * 1e3, (w1 - w0) * 1e3 / (2 * K), (p1 - p0) * 1e3), flush=True)
            for i, r in enumerate(recs):
                print("   sweep %d wall %.0f ms phases(ms) %s" % (i, r['wall_ms'], {k: round(v * 1e3) for k, v in r['phases'].items()}))
        pool.close()
