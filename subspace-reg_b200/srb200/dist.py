"""Seed-parallel sharding over the GPUs of one node (one process per GPU, launched by torchrun).

The incremental-session path shards by SEED exactly like the reference's SLURM array
(scripts/continual/slurm_subspace_reg.sh:7-8,25): runs for different seeds share nothing, sessions inside a seed are
strictly sequential.  There is therefore no data-path collective; the only exchange is one all-reduce(SUM) at the end
over the per-seed accuracy tables (fp64, zero except for owned seeds) and the [100,100] int64 confusion counts.
NCCL over NVLink on the GPUs, gloo in the CPU tests.
"""
import os

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend=None):
    """Initialise torch.distributed from the torchrun environment (no-op for a single process)."""
    rank, world, local = env_rank_world()
    # Host threads of the mask generator (sr_host_bernoulli / sr_host_dropblock walkers): share the node's cores between
    # the ranks instead of letting every rank assume it owns the machine.
    if "SRB_RNG_THREADS" not in os.environ:
        try:
            cores = len(os.sched_getaffinity(0))
        except AttributeError:
            cores = os.cpu_count() or 1
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
        per_rank = cores // max(local_world, 1)
        # The launching thread and the (mostly waiting) prefetch / DropBlock threads share two cores; the generator walkers
        # get the rest, between 2 and 4: one walker tempers + compares ~1.1 G generator words per sweep at ~0.35 ns each,
        # i.e. needs as long as the GPU needs for the sweep - with a single walker (8 ranks on 32 cores under the old
        # `per_rank - 3` rule) the train-mode passes waited for their masks and the 8-GPU efficiency dropped to 0.92.
        os.environ["SRB_RNG_THREADS"] = str(max(2, min(4, per_rank - 2)))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend=backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_seeds(seeds, rank, world):
    """Round-robin deal: rank r owns seeds[r], seeds[r + world], ..."""
    return [s for i, s in enumerate(seeds) if i % world == rank]


def reduce_results(owned, seeds, n_sessions, device, n_classes=100):
    """owned: {seed: dict(weighted=[n_sessions+1], novel=[n_sessions], base=[n_sessions], confusion=int64 [C,C] or None)}.
    Returns (weighted [S, n_sessions+1], novel [S, n_sessions], base [S, n_sessions], confusion [C, C]) identical on
    every rank: one all_reduce(SUM) each over zero-filled buffers."""
    S = len(seeds)
    acc = torch.zeros((S, 3 * n_sessions + 1), dtype=torch.float64, device=device)
    conf = torch.zeros((n_classes, n_classes), dtype=torch.int64, device=device)
    for i, s in enumerate(seeds):
        if s in owned:
            r = owned[s]
            row = list(r['weighted']) + list(r['novel']) + list(r['base'])
            acc[i, :len(row)] = torch.tensor(row, dtype=torch.float64, device=device)
            if r.get('confusion') is not None:
                conf += r['confusion'].to(device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)
        dist.all_reduce(conf, op=dist.ReduceOp.SUM)
    return acc[:, :n_sessions + 1], acc[:, n_sessions + 1:2 * n_sessions + 1], acc[:, 2 * n_sessions + 1:], conf


def max_over_ranks(value, device):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
