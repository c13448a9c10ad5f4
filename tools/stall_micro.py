import time, torch, threading
dev = "cuda"
src = torch.empty(21 << 20, dtype=torch.uint8, pin_memory=True)
dst = torch.empty(21 << 20, dtype=torch.uint8, device=dev)
big = torch.empty(2 << 30, dtype=torch.uint8, device=dev)
cs = torch.cuda.Stream()
def tail(name, ts):
    ts = sorted(ts)
    print("%-34s n=%d median %.3f ms  p99 %.3f  max %.3f ms; >20ms: %d" % (name, len(ts), ts[len(ts)//2]*1e3, ts[int(len(ts)*0.99)]*1e3, ts[-1]*1e3, sum(t > 0.02 for t in ts)))
# background GPU load: a long-running kernel stream like a sweep
def load():
    x = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
    while not stop[0]:
        for _ in range(20): y = x @ x
        torch.cuda.current_stream().synchronize()
stop = [False]
th = threading.Thread(target=load, daemon=True); th.start()
for rep in range(2):
    ts = []
    for i in range(1500):
        t0 = time.perf_counter()
        with torch.cuda.stream(cs):
            dst.copy_(src, non_blocking=True)
        ts.append(time.perf_counter() - t0)
        if i % 50 == 0: cs.synchronize()
        time.sleep(0.002)
    tail("H2D 21 MB pinned -> preallocated", ts)
    ts = []
    for i in range(1500):
        t0 = time.perf_counter()
        with torch.cuda.stream(cs):
            k = src.to(dev, non_blocking=True)
        ts.append(time.perf_counter() - t0)
        del k
        if i % 50 == 0: cs.synchronize()
        time.sleep(0.002)
    tail("H2D 21 MB .to() (allocates)", ts)
    ts = []
    for i in range(1500):
        t0 = time.perf_counter()
        a = torch.empty((185, 84, 84, 64), dtype=torch.float32, device=dev)
        ts.append(time.perf_counter() - t0)
        del a
        time.sleep(0.002)
    tail("torch.empty 334 MB + free", ts)
    ts = []
    for i in range(1500):
        t0 = time.perf_counter()
        big[:1024].zero_()
        ts.append(time.perf_counter() - t0)
        time.sleep(0.002)
    tail("tiny kernel launch", ts)
stop[0] = True
