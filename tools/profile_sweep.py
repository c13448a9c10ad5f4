"""Phase breakdown of one multi-session sweep + head-kernel epoch time per session shape (GPU)."""
import contextlib
import io
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "subspace-reg_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from srb200 import ops, synthetic, _lib as L  # noqa: E402


def head_timing():
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    for s in (1, 4, 8):
        Ns, Nm, nb, npv, nn_, d = 185, 25 * (s - 1), 60, 5 * (s - 1), 5, 640
        Cn = nb + npv + nn_
        feat = (torch.randn(Ns + Nm, d, device=dev, generator=g) * 0.45 + 0.53).clamp_min(-0.11)
        ys = torch.randint(0, Cn, (Ns,), device=dev, generator=g)
        ym = torch.randint(0, Cn, (Nm,), device=dev, generator=g) if Nm else None
        W = (torch.rand(Cn, d, device=dev, generator=g) * 2 - 1) / d ** 0.5
        base = W[:nb].clone()
        reserve = W[nb:nb + npv].clone() if npv else None
        qt, q, _ = ops.subspace_factor(base.contiguous())
        for rep in range(2):
            hs = ops.HeadSession(feat, Ns, 0, ys, W.clone(), nb, nn_, n_memory=Nm, memory_row0=Ns, labels_memory=ym,
                                 base_weight=base, reserve_weight=reserve, pull_mode=L.SR_PULL_PROJECT, pull=qt, q_rows=q,
                                 lmbd_base=0.2, lmbd_novel=0.1, gamma=1.0, stable=False, target_train_loss=-1.0,
                                 min_novel_epochs=0, max_novel_epochs=10 ** 6)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            hs.run(2000)
            e1.record()
            torch.cuda.synchronize()
        t = hs.workspace[:160].cpu().view(torch.int64)[4:16].tolist()
        print("head session %d (N=%d C=%d): %.2f us/epoch; CTA0 ns/epoch phase1 %.0f bar1 %.0f phase2 %.0f bar2 %.0f | loss CTA %.0f pull CTA %.0f" %
              (s, Ns + Nm, Cn, e0.elapsed_time(e1) * 1e3 / 2000, t[0] / 2000, t[1] / 2000, t[2] / 2000, t[3] / 2000, t[4] / 2000, t[5] / 2000), flush=True)
        print("    phase 1: W stream %.0f  logits->smem %.0f  softmax+dlogits %.0f | phase 2: prologue %.0f  DLt stream %.0f  update %.0f (ns/epoch)" %
              tuple(x / 2000 for x in t[6:12]), flush=True)
        print("    raw t_ns/epoch:", [round(x / 2000) for x in t], flush=True)


def sweep_timing():
    wdir = bench.word_embed_dir()
    for it in range(int(os.environ.get('SWEEPS', '2'))):
        t0 = time.perf_counter()
        world = synthetic.make_world(10 + it, n_sessions=8, n_base_batch=1000, word_embed_path=wdir)
        t1 = time.perf_counter()
        prep = bench.prepare(bench.place_world(world, 'gpu'))
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        rec = bench.run_sweeps([prep], None)[0]
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        print("sweep %d: make_world %.2fs prepare %.2fs sweep %.2fs; phases %s" % (it, t1 - t0, t2 - t1, t3 - t2, rec.get('phases')), flush=True)
        print("  epochs", [s['epochs'] for s in rec['sessions']], flush=True)


if __name__ == "__main__":
    head_timing()
    sweep_timing()
