"""CUDA-event timing of the eval-mode backbone pass (ms per N images)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "subspace-reg_b200"))
import torch  # noqa: E402

from models.util import create_model  # noqa: E402
from srb200 import synthetic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
net = synthetic.init_model(create_model, synthetic.default_opt(1), 1).cuda().eval()
x = torch.randn(n, 3, 84, 84, device="cuda")
eng = net.engine()
with torch.no_grad():
    for _ in range(3):
        eng.eval_features(x)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        eng.eval_features(x)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print("backbone %d images: %.3f ms/pass  %.0f img/s" % (n, ms, n / ms * 1e3))
