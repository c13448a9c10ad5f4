"""Eval-mode backbone pass on N synthetic images (ncu target: 18 conv_umma_kernel launches per pass)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "subspace-reg_b200"))
import torch  # noqa: E402

from models.util import create_model  # noqa: E402
from srb200 import synthetic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
net = synthetic.init_model(create_model, synthetic.default_opt(1), 1).cuda().eval()
x = torch.randn(n, 3, 84, 84, device="cuda")
with torch.no_grad():
    for _ in range(reps):
        f = net.engine().eval_features(x)
torch.cuda.synchronize()
print("ok", tuple(f.shape), float(f.mean()))
