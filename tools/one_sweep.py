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
    print("sweep %d: phases %s epochs %s" % (it, {k: round(v * 1e3, 1) for k, v in rec['phases'].items()},
                                              [s['epochs'] for s in rec['sessions']]), flush=True)
