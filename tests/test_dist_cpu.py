"""world_size-2 gloo test of the N>1 path: seed sharding + the single results all-reduce (no data-path collective)."""
import os
import socket
import sys

import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "subspace-reg_b200"))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from srb200 import dist as sdist
    r, w, _ = sdist.init(backend="gloo")
    seeds = list(range(1, 6))
    mine = sdist.shard_seeds(seeds, r, w)
    owned = {}
    for s in mine:
        conf = torch.zeros((100, 100), dtype=torch.int64)
        conf[s, s] = 10 * s
        owned[s] = dict(weighted=[s + 0.25 * i for i in range(9)], novel=[10.0 * s] * 8, base=[s / 8.0] * 8, confusion=conf)
    weighted, novel, base, conf = sdist.reduce_results(owned, seeds, 8, torch.device("cpu"))
    t = sdist.max_over_ranks(1.0 + rank, torch.device("cpu"))
    tot = sdist.sum_over_ranks(len(mine), torch.device("cpu"))
    sdist.barrier()
    q.put((rank, mine, weighted.tolist(), novel.tolist(), base.tolist(), conf.tolist(), t, tot))
    torch.distributed.destroy_process_group()


def test_seed_sharding_and_reduction_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[0][1] == [1, 3, 5] and res[1][1] == [2, 4]
    for rank, mine, weighted, novel, base, conf, t, tot in res:
        weighted, novel, base = (torch.tensor(x, dtype=torch.float64) for x in (weighted, novel, base))
        conf = torch.tensor(conf)
        assert weighted.shape == (5, 9) and novel.shape == (5, 8) and base.shape == (5, 8)
        for i, s in enumerate(range(1, 6)):
            assert torch.allclose(weighted[i], torch.tensor([s + 0.25 * j for j in range(9)], dtype=torch.float64))
            assert float(novel[i, 0]) == 10.0 * s and abs(float(base[i, 7]) - s / 8.0) < 1e-12
            assert int(conf[s, s]) == 10 * s
        assert int(conf.sum()) == 10 * 15 and t == 2.0 and tot == 5.0
    assert res[0][2] == res[1][2]
