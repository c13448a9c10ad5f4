"""Device time of the mask generator kernels for one train-mode forward (batch 125 / 220), alone on the GPU."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "subspace-reg_b200"))
import torch  # noqa: E402

from srb200 import device_rng  # noqa: E402

dev = torch.device("cuda", 0)
assert device_rng.available(dev)
for batch in (125, 220, 360):
    shapes = [(0, 0.9, (batch, 64, 42, 42)), (0, 0.9, (batch, 160, 21, 21)), (1, 0.0173, (batch, 320, 6, 6)),
              (1, 0.0173, (batch, 320, 6, 6)), (1, 0.004, (batch, 640, 1, 1)), (1, 0.5, (batch, 640, 1, 1))]
    outs = [torch.empty(s, dtype=torch.uint8, device=dev) for _, _, s in shapes]
    regions = [(k, p, o) for (k, p, _), o in zip(shapes, outs)]
    words = sum(device_rng.region_words(k, o.numel()) for k, _, o in regions)
    ws = None
    torch.manual_seed(1)
    for rep in range(3):
        ws = device_rng.draw(regions, ws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for rep in range(10):
        ws = device_rng.draw(regions, ws)
    e1.record()
    torch.cuda.synchronize()
    print("batch %d: %.1f M words, %d walkers: %.3f ms per forward" % (batch, words / 1e6, (words + (1 << 19) - 1) >> 19,
                                                                     e0.elapsed_time(e1) / 10), flush=True)
