"""cProfile of the host side of train-mode forwards (dropout masks prefetched, DropBlock masks drawn ahead)."""
import cProfile
import os
import pstats
import sys

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
    b_["mod"].num_batches_tracked += 3000
eng = net.engine()
x = torch.randn(B, 3, 84, 84, device="cuda")
for _ in range(3):
    net.features(x)
torch.cuda.synchronize()
N = 12
eng.start_mask_prefetch([(0, B)] * N)
eng._prefetch.thread.join()
pr = cProfile.Profile()
for i in range(N):
    nbt = next(iter(net.block_counters().values()))
    eng.start_dropblock_ahead([(B, nbt + 1)])
    eng._db_thread.join()
    torch.cuda.synchronize()
    pr.enable()
    net.features(x)
    pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)
