"""Multi-GPU plumbing: one process per GPU over torch.distributed (NCCL on NVLink 5 /
NVSwitch; gloo in CPU tests).  The reference is single-process; the path
shards naturally, so only two exchange steps exist:

  X1  all-gather of the sorted basis words + amplitudes before extraction (every rank
      searches the FULL basis for the neighbours of its own row block);
  X2  after annealing, all-gather of (best energy, rank) pairs, local argmin (NCCL has no
      MINLOC; ties -> lowest rank), broadcast of the winner's packed sign vector.

Row blocks and replica blocks are contiguous and deterministic, so the sharded result is
bit-identical to the single-GPU one (KAT-7).
"""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def init_from_env() -> Tuple[int, int, int]:
    """-> (rank, world_size, local_rank); initialises the process group when WORLD_SIZE > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group("gloo")
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return rank, world, local


def block(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [begin, begin+count) of `total` items owned by `rank`."""
    begin = total * rank // world
    end = total * (rank + 1) // world
    return begin, end - begin


def all_gather_blocks(local: torch.Tensor, total: int) -> torch.Tensor:
    """X1: concatenate every rank's contiguous block (sizes from block()) into [total]."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [block(total, r, world)[1] for r in range(world)]
    longest = max(sizes)
    if min(sizes) == longest:  # equal blocks: gather straight into place, no padding, no copy
        gathered = torch.empty(total, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(gathered, local.contiguous())
        return gathered
    padded = torch.zeros(longest, dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    gathered = torch.empty(world * longest, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, padded)
    parts = [gathered[r * longest: r * longest + sizes[r]] for r in range(world)]
    return torch.cat(parts)


def exclusive_offset(local_count: int, device) -> Tuple[int, int]:
    """Global indptr offset of this rank's row block: (sum of earlier ranks, grand total)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return 0, local_count
    world, rank = dist.get_world_size(), dist.get_rank()
    mine = torch.tensor([local_count], dtype=torch.int64, device=device)
    every = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(every, mine)
    every = every.cpu()
    return int(every[:rank].sum()), int(every.sum())


def reduce_best(best_energy: float, best_bits: torch.Tensor) -> Tuple[float, torch.Tensor, int]:
    """X2: global best replica -> (energy, bits, owner rank) on every rank."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return best_energy, best_bits, 0
    world = dist.get_world_size()
    mine = torch.tensor([best_energy], dtype=torch.float64, device=best_bits.device)
    every = torch.empty(world, dtype=torch.float64, device=best_bits.device)
    dist.all_gather_into_tensor(every, mine)
    every = every.cpu()
    owner = int(torch.nonzero(every == every.min())[0])
    bits = best_bits.clone()
    dist.broadcast(bits, src=owner)
    return float(every[owner]), bits, owner


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def max_over_ranks(value: float, device) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def sum_over_ranks(value: float, device) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t[0])
