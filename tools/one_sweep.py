"""One (or SWEEPS) full config-2 sweeps through the public API, for profilers."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "subspace-reg_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from srb200 import synthetic  # noqa: E402

wdir = bench.word_embed_dir()
for it in range(int(os.environ.get("SWEEPS", "1"))):
    world = synthetic.make_world(10 + it, n_sessions=8, n_base_batch=1000, word_embed_path=wdir,
                                 conv_precision=os.environ.get("PRECISION", "bf16"))
    prep = bench.prepare(bench.place_world(world, 'gpu'))
    rec = bench.run_sweeps([prep], None)[0]
    torch.cuda.synchronize()
    from srb200 import host_rng
    print("sweep %d: wall %.1f ms phases %s epochs %s | mask wait %.1f ms, inline draws %d (%.1f ms)" % (
        it, rec['wall_ms'], {k: round(v * 1e3, 1) for k, v in rec['phases'].items()}, [s['epochs'] for s in rec['sessions']],
        host_rng.WAIT_S[0] * 1e3, host_rng.WAIT_S[2], host_rng.WAIT_S[1] * 1e3), flush=True)
    host_rng.WAIT_S[:] = [0.0, 0.0, 0]
    ms = torch.cuda.memory_stats()
    import gc
    print("    allocator: cudaMalloc calls %d, cudaFree calls %d, retries %d, reserved %.1f GB; gc counts %s" % (
        ms.get('num_device_alloc', 0), ms.get('num_device_free', 0), ms.get('num_alloc_retries', 0),
        ms.get('reserved_bytes.all.current', 0) / 1e9, gc.get_count()), flush=True)
