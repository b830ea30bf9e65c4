"""Multi-GPU plumbing: one process per GPU under ``torch.distributed`` (NCCL on GPUs, gloo in CPU tests).

Two modes use it:
* **site-segment-parallel TDVP** of one chain (``pytdscf_b200/_mps_parallel.py``, the reference's ``MPSCoefParallel``):
  nearest-neighbour point-to-point exchange of environment blocks, bond matrices and centre tensors (``Comm`` there);
* **replicas** -- every rank propagates an independent wavefunction (the reference's semantics for Liouville-space /
  trajectory workloads, which its site-parallel path excludes, ``pytdscf/_mps_parallel.py:82-87``); no data-path
  collective, only the control-plane reductions below (MAX of times, SUM of work)."""
from __future__ import annotations

import os
from dataclasses import dataclass


@dataclass
class RankInfo:
    rank: int
    world: int
    local_rank: int
    dist: object | None  # torch.distributed when world > 1


def init_from_env(backend: str | None = None) -> RankInfo:
    """Read RANK / WORLD_SIZE / LOCAL_RANK (torchrun) and join the process group when WORLD_SIZE > 1."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1:
        return RankInfo(rank, world, local_rank, None)
    import torch
    import torch.distributed as dist

    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend, **kw)
    return RankInfo(rank, world, local_rank, dist)


def barrier(info: RankInfo):
    if info.dist is not None:
        info.dist.barrier()


def max_over_ranks(info: RankInfo, value: float, device=None) -> float:
    """MAX all-reduce of a scalar (timings are reported as the slowest rank's)."""
    if info.dist is None:
        return float(value)
    import torch

    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    info.dist.all_reduce(t, op=info.dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(info: RankInfo, value: float, device=None) -> float:
    if info.dist is None:
        return float(value)
    import torch

    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    info.dist.all_reduce(t, op=info.dist.ReduceOp.SUM)
    return float(t.item())


def aggregate_throughput(info: RankInfo, units_this_rank: float, seconds_this_rank: float, device=None) -> float:
    """Whole-job throughput of independent replicas: total units of all ranks / time of the slowest rank."""
    return sum_over_ranks(info, units_this_rank, device) / max_over_ranks(info, seconds_this_rank, device)


def finalize(info: RankInfo):
    if info.dist is not None and info.dist.is_initialized():
        info.dist.destroy_process_group()
