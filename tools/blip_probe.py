import os, sys, gc, time
sys.path.insert(0, "subspace-reg_b200"); sys.path.insert(0, "."); sys.path.insert(0, "tools")
import torch, bench
from srb200 import synthetic
mode = sys.argv[1]
wdir = bench.word_embed_dir()
worlds = [bench.prepare(bench.place_world(synthetic.make_world(10 + i, n_sessions=8, n_base_batch=1000, word_embed_path=wdir), 'gpu')) for i in range(24)]
if mode == "nogc":
    gc.collect(); gc.freeze(); gc.disable()
out = []
for w in worlds:
    r = bench.run_sweeps([w], None)[0]
    out.append((round(r['wall_ms']), round(r['phases']['train_pass'] * 1e3)))
print(mode, out)
