"""Where one train-mode forward spends its time: per-block mask draws (host) vs everything else."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "subspace-reg_b200"))
import torch  # noqa: E402

from models.util import create_model  # noqa: E402
from srb200 import synthetic  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 185
net = synthetic.init_model(create_model, synthetic.default_opt(1), 1).cuda().train()
for p_ in net.parameters():
    p_.requires_grad = False
for b_ in net._blocks():
    b_["mod"].num_batches_tracked += 3000   # DropBlock active, as in a real session
eng = net.engine()
x = torch.randn(B, 3, 84, 84, device="cuda")
times = {}
orig = eng.draw_mask


def timed(bi, *a, **k):
    t0 = time.perf_counter()
    r = orig(bi, *a, **k)
    times[bi] = times.get(bi, 0.0) + time.perf_counter() - t0
    return r


eng.draw_mask = timed
for rep in range(4):
    times.clear()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    f = net.features(x)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("forward B=%d: host %.1f ms, +sync %.1f ms; draw_mask per block (ms): %s" %
          (B, (t1 - t0) * 1e3, (t2 - t1) * 1e3, {k: round(v * 1e3, 1) for k, v in times.items()}), flush=True)

# ---- with the dropout masks prefetched (the session path): what is left on the critical path? ----
print("with prefetch:")
for rep in range(3):
    t0 = time.perf_counter()
    eng.start_mask_prefetch([(0, B)])
    eng._prefetch.thread.join()
    t1 = time.perf_counter()
    times.clear()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    f = net.features(x)
    t3 = time.perf_counter()
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    print("prefetch thread %.1f ms | forward host %.1f ms, +sync %.1f ms; draw_mask per block (ms): %s" %
          ((t1 - t0) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3, {k: round(v * 1e3, 1) for k, v in times.items()}), flush=True)
