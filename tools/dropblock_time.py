"""Host-side DropBlock mask cost: seed draw + block dilation (C helper vs the torch formulation)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "subspace-reg_b200"))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from srb200 import host_rng  # noqa: E402

host_rng.replay_available()
for (B, Cc, size, bs) in ((185, 320, 10, 5), (185, 640, 5, 3), (385, 320, 10, 5)):
    hs = size - (bs - 1)
    out = torch.empty(B, Cc, size, size, dtype=torch.uint8).pin_memory()
    for rep in range(3):
        t0 = time.perf_counter()
        seeds, n_seed = host_rng.bernoulli_u8((B, Cc, hs, hs), 0.02, 1)
        t1 = time.perf_counter()
        kept = host_rng.dropblock_keep(seeds, bs, out)
        t2 = time.perf_counter()
        left, right = int((bs - 1) / 2), int(bs / 2)
        padded = F.pad(seeds, (left, right, left, right))
        for i in range(bs):
            for j in range(bs):
                padded[:, :, i:i + hs, j:j + hs] = torch.maximum(padded[:, :, i:i + hs, j:j + hs], seeds)
        out.copy_(1 - padded)
        kept2 = int(out.sum())
        t3 = time.perf_counter()
    print("B%d C%d %dx%d bs%d: seeds %.2f ms | C helper %.2f ms | torch ops %.2f ms  (kept %d / %d)" %
          (B, Cc, size, size, bs, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, kept, kept2))
