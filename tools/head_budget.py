"""Paper-size head kernel under a CTA budget (sr_head_args.cta_budget): epoch time alone, and K budgeted head loops on K
streams of one GPU at the same time (the co-residency the budget exists for).  GPU."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "subspace-reg_b200"))
import torch  # noqa: E402

from srb200 import ops, _lib as L  # noqa: E402

EPOCHS = 2000


def problem(s, g, dev="cuda"):
    Ns, Nm, nb, npv, nn_, d = 185, 25 * (s - 1), 60, 5 * (s - 1), 5, 640
    Cn = nb + npv + nn_
    feat = (torch.randn(Ns + Nm, d, device=dev, generator=g) * 0.45 + 0.53).clamp_min(-0.11)
    ys = torch.randint(0, Cn, (Ns,), device=dev, generator=g)
    ym = torch.randint(0, Cn, (Nm,), device=dev, generator=g) if Nm else None
    W = (torch.rand(Cn, d, device=dev, generator=g) * 2 - 1) / d ** 0.5
    base = W[:nb].clone()
    reserve = W[nb:nb + npv].clone() if npv else None
    qt, q, _ = ops.subspace_factor(base.contiguous())

    def make(budget):
        return ops.HeadSession(feat, Ns, 0, ys, W.clone(), nb, nn_, n_memory=Nm, memory_row0=Ns, labels_memory=ym,
                               base_weight=base, reserve_weight=reserve, pull_mode=L.SR_PULL_PROJECT, pull=qt, q_rows=q,
                               lmbd_base=0.2, lmbd_novel=0.1, gamma=1.0, stable=False, target_train_loss=-1.0,
                               min_novel_epochs=0, max_novel_epochs=10 ** 6, cta_budget=budget)
    return make, Ns + Nm, Cn


def main():
    g = torch.Generator(device="cuda").manual_seed(0)
    for s in (1, 4, 8):
        make, n, c = problem(s, g)
        ref = None
        for budget in (0, 74, 49):
            for rep in range(2):
                hs = make(budget)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                tr = hs.run(EPOCHS)
                e1.record()
                torch.cuda.synchronize()
            if ref is None:
                ref = (tr, hs.weight.clone())
            dl = float((tr[:, 0] - ref[0][:, 0]).abs().max() / ref[0][:, 0].abs().max())
            dw = float((hs.weight - ref[1]).abs().max() / ref[1].abs().max())
            print("session %d (N=%d C=%d) budget %3d: %.2f us/epoch alone; vs budget 0: loss %.1e W %.1e" %
                  (s, n, c, budget, e0.elapsed_time(e1) * 1e3 / EPOCHS, dl, dw), flush=True)
        for k, budget in ((2, 74), (3, 49), (2, 0)):
            streams = [torch.cuda.Stream() for _ in range(k)]
            for rep in range(2):
                sessions = [make(budget) for _ in range(k)]
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for st, hs in zip(streams, sessions):
                    st.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(st):
                        hs.run(EPOCHS, defer=True)
                for st in streams:
                    torch.cuda.current_stream().wait_stream(st)
                e1.record()
                torch.cuda.synchronize()
                for hs in sessions:
                    hs.collect()
            ms = e0.elapsed_time(e1)
            print("    %d loops at once, budget %3d: %.2f ms for %d x %d epochs = %.2f us per epoch per loop" %
                  (k, budget, ms, k, EPOCHS, ms * 1e3 / (k * EPOCHS)), flush=True)


if __name__ == "__main__":
    main()
