"""Where does the launching thread sit when a sweep stalls?  Runs sweeps back to back and samples the main thread's Python
stack every 5 ms from a watchdog thread; stacks that do not move for > 40 ms are reported with their duration."""
import collections
import gc
import os
import sys
import threading
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "subspace-reg_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from srb200 import synthetic  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "nogc"
n = int(os.environ.get("SWEEPS", "24"))
wdir = bench.word_embed_dir()
worlds = [bench.prepare(bench.place_world(synthetic.make_world(10 + i, n_sessions=8, n_base_batch=1000, word_embed_path=wdir), 'gpu'))
          for i in range(n)]
if os.environ.get("PREWARM_GB"):
    x = torch.empty(int(os.environ["PREWARM_GB"]) << 30, dtype=torch.uint8, device="cuda")
    del x
if mode == "nogc":
    gc.collect()
    gc.freeze()
    gc.disable()

main_id = threading.main_thread().ident
stalls = collections.Counter()
stop = [False]


def watch():
    last, since = None, time.perf_counter()
    while not stop[0]:
        time.sleep(0.005)
        fr = sys._current_frames().get(main_id)
        if fr is None:
            continue
        key = tuple((f.f_code.co_filename.split('/')[-1], f.f_lineno) for f, _ in traceback.walk_stack(fr))[:6]
        now = time.perf_counter()
        if key != last:
            if last is not None and now - since > 0.04:
                stalls[(last, round((now - since) * 1e3, -1))] += 1
            last, since = key, now


th = threading.Thread(target=watch, daemon=True)
th.start()
out = []
for w in worlds:
    r = bench.run_sweeps([w], None)[0]
    out.append(round(r['wall_ms']))
stop[0] = True
th.join()
print(mode, out, 'cudaMalloc calls', torch.cuda.memory_stats().get('num_device_alloc'))
for (key, ms), c in sorted(stalls.items(), key=lambda kv: -kv[0][1])[:25]:
    print("%4d ms x%d  %s" % (ms, c, " <- ".join("%s:%d" % k for k in key[:5])))
