"""One fused-head run of a given session shape (ncu target for head_small_kernel)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "subspace-reg_b200"))
import torch  # noqa: E402

from srb200 import ops, _lib as L  # noqa: E402

s = int(sys.argv[1]) if len(sys.argv) > 1 else 8
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 300
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
Ns, Nm, nb, npv, nn_, d = 185, 25 * (s - 1), 60, 5 * (s - 1), 5, 640
Cn = nb + npv + nn_
feat = (torch.randn(Ns + Nm, d, device=dev, generator=g) * 0.45 + 0.53).clamp_min(-0.11)
ys = torch.randint(0, Cn, (Ns,), device=dev, generator=g)
ym = torch.randint(0, Cn, (Nm,), device=dev, generator=g) if Nm else None
W = (torch.rand(Cn, d, device=dev, generator=g) * 2 - 1) / d ** 0.5
base = W[:nb].clone()
reserve = W[nb:nb + npv].clone() if npv else None
qt, q, _ = ops.subspace_factor(base.contiguous())
hs = ops.HeadSession(feat, Ns, 0, ys, W.clone(), nb, nn_, n_memory=Nm, memory_row0=Ns, labels_memory=ym, base_weight=base,
                     reserve_weight=reserve, pull_mode=L.SR_PULL_PROJECT, pull=qt, q_rows=q, lmbd_base=0.2, lmbd_novel=0.1,
                     gamma=1.0, stable=False, target_train_loss=-1.0, min_novel_epochs=0, max_novel_epochs=10 ** 6)
hs.run(epochs)
torch.cuda.synchronize()
print("ok", hs.epochs)
