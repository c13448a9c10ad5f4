"""BASELINE config 5 (1000 base + 100 novel classes, 100-shot, 512-d) through sr_head_run: a few fine-tune steps for ncu
launch lists / full captures of the tensor-core head (csrc/head_tc.cu).  STEPS / NBASE from the environment."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "subspace-reg_b200"))
import torch  # noqa: E402

from srb200 import ops, _lib as L  # noqa: E402

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(1)
n_base = int(os.environ.get("NBASE", "1000"))
steps = int(os.environ.get("STEPS", "5"))
N, n_new, d = 10000, 100, 512
Cn = n_base + n_new
X = (torch.randn(N, d, device=dev, generator=g) * 0.45 + 0.53).clamp_min(-0.11)
y = n_base + torch.arange(n_new, device=dev).repeat_interleave(N // n_new)
W = (torch.rand(Cn, d, device=dev, generator=g) * 2 - 1) / d ** 0.5
base = W[:n_base].clone().contiguous()
qt, q, _ = ops.subspace_factor(base)
hs = ops.HeadSession(X, N, 0, y, W, n_base, n_new, base_weight=base, pull_mode=L.SR_PULL_PROJECT, pull=qt, q_rows=q,
                     lmbd_base=0.2, gamma=1.0, stable=False, target_train_loss=-1.0, min_novel_epochs=0,
                     max_novel_epochs=10 ** 6)
hs.run(2)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
hs.run(steps, defer=True)      # queued only: the read-back is not part of a step
e1.record()
torch.cuda.synchronize()
tr = hs.collect()
print("config 5 (n_base %d): %.1f us per step over %d steps, last loss %.6f" % (n_base, e0.elapsed_time(e1) * 1e3 / steps, steps,
                                                                                float(tr[-1, 0])))
