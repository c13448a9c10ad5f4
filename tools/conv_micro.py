"""Time one conv layer shape with CUDA events (SRB_CONV_DBG experiments)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "subspace-reg_b200"))
import torch
from srb200 import ops, _lib as L
B, H, cin, cout = [int(v) for v in sys.argv[1:5]]
epi = int(sys.argv[5]) if len(sys.argv) > 5 else 0
act = torch.randn(B, H, H, cin, device="cuda").to(torch.bfloat16)
w = (torch.randn(cout, 9, cin, device="cuda") / (9 * cin) ** 0.5).to(torch.bfloat16)
shift = torch.zeros(cout, device="cuda")
for _ in range(3):
    ops.conv([(act, w)], cout, shift=shift, epilogue=epi)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(10):
    ops.conv([(act, w)], cout, shift=shift, epilogue=epi)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
fl = 2.0 * B * H * H * cout * 9 * cin
print("dbg=%s B%d %dx%d %d->%d epi%d: %.3f ms  %.0f TFLOP/s" % (os.environ.get("SRB_CONV_DBG", "0"), B, H, H, cin, cout, epi, ms, fl / ms / 1e9))
