import sys
sys.path.insert(0, "subspace-reg_b200")
import torch
from srb200 import ops, _lib as L
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
Ns, Nm, nb, npv, nn_, d = 185, 175, 60, 35, 5, 640
Cn = nb + npv + nn_
feat = (torch.randn(Ns + Nm, d, device=dev, generator=g) * 0.45 + 0.53).clamp_min(-0.11)
ys = torch.randint(0, Cn, (Ns,), device=dev, generator=g)
ym = torch.randint(0, Cn, (Nm,), device=dev, generator=g)
W = (torch.rand(Cn, d, device=dev, generator=g) * 2 - 1) / d ** 0.5
base = W[:nb].clone(); reserve = W[nb:nb + npv].clone()
qt, q, _ = ops.subspace_factor(base.contiguous())
for rep in range(2):
    hs = ops.HeadSession(feat, Ns, 0, ys, W.clone(), nb, nn_, n_memory=Nm, memory_row0=Ns, labels_memory=ym, base_weight=base,
                         reserve_weight=reserve, pull_mode=L.SR_PULL_PROJECT, pull=qt, q_rows=q, lmbd_base=0.2, lmbd_novel=0.1,
                         gamma=1.0, stable=False, target_train_loss=-1.0, min_novel_epochs=0, max_novel_epochs=10 ** 6)
    hs.run(500)
torch.cuda.synchronize()
