"""Small invocations of the round-2 kernels for compute-sanitizer (memcheck / racecheck): device mask generator, DropBlock,
the one-launch mapping fit, the budgeted head shape."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "subspace-reg_b200"))
import torch  # noqa: E402

from srb200 import device_rng, mapping, ops, _lib as L  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(3)
torch.rand(9)
s0 = torch.get_rng_state()
n0, n1 = 300001 * 2, 4001
want0 = torch.empty(n0, dtype=torch.uint8).bernoulli_(0.9)
want1 = torch.bernoulli(torch.tensor(0.02).expand(n1)).to(torch.uint8)
torch.set_rng_state(s0)
o0 = torch.empty(n0, dtype=torch.uint8, device=dev)
o1 = torch.empty(n1 + 1, dtype=torch.uint8, device=dev)[:n1]
device_rng.draw([(0, 0.9, o0), (1, 0.02, o1)])
assert torch.equal(o0.cpu(), want0) and torch.equal(o1.cpu(), want1)
seeds = (torch.rand(3, 320, 6, 6, device=dev) < 0.02).to(torch.uint8)
keep = torch.empty(3, 320, 10, 10, dtype=torch.uint8, device=dev)
scale = torch.zeros(4, device=dev)
device_rng.dropblock_keep(seeds, 5, keep, scale)
g = torch.Generator().manual_seed(0)
E = (torch.randn(60, 300, generator=g) * 0.3).cuda()
T = (torch.randn(60, 640, generator=g) * 0.05).cuda()
init = {'map.weight': torch.randn(640, 300, generator=g) * 0.02, 'map.bias': torch.zeros(640)}
mapping.fit_linear_map(E, T, epochs=3, init=init)
gd = torch.Generator(device=dev).manual_seed(0)
Ns, Nm, nb, npv, nn_, d = 185, 75, 60, 15, 5, 640
Cn = nb + npv + nn_
feat = torch.randn(Ns + Nm, d, device=dev, generator=gd)
ys = torch.randint(0, Cn, (Ns,), device=dev, generator=gd)
ym = torch.randint(0, Cn, (Nm,), device=dev, generator=gd)
W = (torch.rand(Cn, d, device=dev, generator=gd) * 2 - 1) / d ** 0.5
base, reserve = W[:nb].clone(), W[nb:nb + npv].clone()
qt, q, _ = ops.subspace_factor(base.contiguous())
hs = ops.HeadSession(feat, Ns, 0, ys, W.clone(), nb, nn_, n_memory=Nm, memory_row0=Ns, labels_memory=ym, base_weight=base,
                     reserve_weight=reserve, pull_mode=L.SR_PULL_PROJECT, pull=qt, q_rows=q, lmbd_base=0.2, lmbd_novel=0.1,
                     gamma=1.0, stable=False, target_train_loss=-1.0, min_novel_epochs=0, max_novel_epochs=10 ** 6, cta_budget=49)
hs.run(3)
torch.cuda.synchronize()
print("sanitize_new_kernels: ok")
