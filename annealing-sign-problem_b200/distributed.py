"""Multi-GPU plumbing: one process per GPU over torch.distributed (NCCL on NVLink 5 /
NVSwitch; gloo in CPU tests).  The reference is single-process; the path
shards naturally, so only two exchange steps exist:

  X1  all-gather of the sorted basis words + amplitudes before extraction (every rank
      searches the FULL basis for the neighbours of its own row block);
  X2  after annealing, all-gather of (best energy, rank) pairs, local argmin (NCCL has no
      MINLOC; ties -> lowest rank), broadcast of the winner's packed sign vector.

Row blocks and replica blocks are contiguous and deterministic, so the sharded result is
bit-identical to the single-GPU one (KAT-7).

X1 has two implementations: ``all_gather_blocks`` (NCCL; gloo in CPU tests) and ``PeerBasis``
(the row blocks stay in NVLink peer memory and ONE kernel pulls + indexes them:
asp_gather_index, csrc/exchange_kernels.cuh -- the fast path on a B200 node).
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


class _RawDeviceMemory:
    """Zero-copy torch view of memory the C library allocated (``torch.as_tensor`` honours
    ``__cuda_array_interface__``)."""

    def __init__(self, address: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (address, False), "version": 2}


class PeerMemoryUnavailable(RuntimeError):
    """Peer memory could not be set up on at least one rank (raised on EVERY rank, so the caller's
    choice of exchange stays collective)."""


class PeerBasis:
    """X1 over NVLink peer memory: this rank's row block of the sorted basis (keys + amplitudes)
    lives in a buffer every other process of the node maps through CUDA IPC.

    Per exchange epoch (all calls enqueue on the current stream; ranks are ordered by epoch flags
    in peer memory, not by host barriers or NCCL calls):

        pb.begin_epoch()            # waits until every peer has finished READING my previous block
        pb.spins[:m] = ...; pb.psi[:m] = ...   # produce the block in place (or leave it)
        pb.publish()                # release: "my block is complete for this epoch" -> all peers
        full_spins, full_psi = pb.gather_index(op, shard_begin, num_rows, workspace)
                                    # ONE kernel: pull all blocks, private full copy, table + filter
        pb.release()                # "I have finished reading this epoch" -> all peers
        ... asp_extract_csr_indexed on the indexed workspace ...
    """

    FLAG_BYTES = 256  # ready[16] u64 at +0, done[16] u64 at +128
    MODES = {"ce": 0, "sm": 1, "tma": 2}
    DONE_SLOT = 16

    def __init__(self, capacity_rows: int, device=None, mode: str = "tma"):
        """mode (asp_set_gather_mode): "tma" = one persistent kernel, cp.async.bulk pulls chunks into shared
        memory while the threads index the chunk that landed; "sm" = one kernel, plain 16-byte loads;
        "ce" = copy engines pull, SMs index block by block behind them."""
        from ._lib import check, ffi, lib, require_cuda

        assert mode in self.MODES
        self.mode = mode

        self.device = device if device is not None else require_cuda()
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        if self.world > 16:
            raise ValueError("PeerBasis supports at most 16 ranks (one NVSwitch domain)")
        self.capacity = (int(capacity_rows) + 31) // 32 * 32
        nbytes = self.FLAG_BYTES + 16 * self.capacity
        # Every collective below is reached by EVERY rank whatever failed locally, so a rank that cannot
        # allocate or map peer memory makes all ranks raise together (callers fall back to NCCL in step).
        own = ffi.new("void **")
        handle = ffi.new("unsigned char[64]")
        error = None
        self._own, self._opened = 0, []
        try:
            check(lib().asp_peer_alloc(nbytes, own, handle))
            self._own = int(ffi.cast("uintptr_t", own[0]))
        except Exception as exc:  # noqa: BLE001
            error = "rank %d: %s" % (self.rank, exc)
        bases = [0] * self.world
        bases[self.rank] = self._own
        if self.world > 1:
            mine = (bytes(ffi.buffer(handle, 64)) if error is None else None, self.capacity, error)
            every = [None] * self.world
            dist.all_gather_object(every, mine)
            errors = [e for _, _, e in every if e]
            if not errors and any(cap != self.capacity for _, cap, _ in every):
                errors = ["every rank must pass the same capacity_rows"]
            if not errors:
                try:
                    for q, (h, _, _) in enumerate(every):
                        if q == self.rank:
                            continue
                        out = ffi.new("void **")
                        check(lib().asp_peer_open(ffi.from_buffer("unsigned char[]", h), out))
                        bases[q] = int(ffi.cast("uintptr_t", out[0]))
                        self._opened.append(bases[q])
                except Exception as exc:  # noqa: BLE001
                    error = "rank %d: %s" % (self.rank, exc)
                outcome = [None] * self.world
                dist.all_gather_object(outcome, error)
                errors = [e for e in outcome if e]
            if errors:
                for address in self._opened:
                    lib().asp_peer_close(ffi.cast("void *", address))
                self._opened = []
                barrier()
                if self._own:
                    lib().asp_peer_free(ffi.cast("void *", self._own))
                    self._own = 0
                raise PeerMemoryUnavailable("; ".join(errors))
        elif error is not None:
            raise PeerMemoryUnavailable(error)
        self._bases = bases
        raw = torch.as_tensor(_RawDeviceMemory(self._own, nbytes), device=self.device)
        self._raw = raw
        body = raw[self.FLAG_BYTES:]
        self.spins = body[: 8 * self.capacity].view(torch.int64)
        self.psi = body[8 * self.capacity:].view(torch.float64)
        self._flags = ffi.new("uint64_t *[]", [ffi.cast("uint64_t *", b) for b in bases])
        self._shard_spins = ffi.new("uint64_t const *[]", [ffi.cast("uint64_t const *", b + self.FLAG_BYTES) for b in bases])
        self._shard_psi = ffi.new("double const *[]", [ffi.cast("double const *", b + self.FLAG_BYTES + 8 * self.capacity) for b in bases])
        self.epoch = 0
        self._full = None

    # -- epoch protocol ------------------------------------------------------------------------
    def begin_epoch(self):
        from ._lib import check, ffi, lib, stream

        self.epoch += 1
        if self.epoch > 1:
            check(lib().asp_peer_wait(self.world, ffi.cast("uint64_t const *", self._own + 8 * self.DONE_SLOT), self.epoch - 1, stream()))

    def publish(self):
        from ._lib import check, lib, stream

        check(lib().asp_peer_signal(self.world, self._flags, self.rank, self.epoch, stream()))

    def release(self):
        from ._lib import check, lib, stream

        check(lib().asp_peer_signal(self.world, self._flags, self.DONE_SLOT + self.rank, self.epoch, stream()))

    def gather_index(self, operator, shard_begin, num_rows: int, workspace: torch.Tensor, slot: int = 0):
        """-> (full_spins int64 [n_total], full_psi f64 [n_total]); ``workspace`` is left indexed
        for asp_extract_csr_indexed(operator, n_total, ..., num_rows).  ``slot`` selects one of the
        private copies (a two-deep pipeline gathers into one while the other is being extracted)."""
        from ._lib import check, ffi, lib, ptr, stream

        begins = [int(b) for b in shard_begin]
        assert len(begins) == self.world + 1 and begins[0] == 0
        assert all(begins[q + 1] - begins[q] <= self.capacity for q in range(self.world))
        n_total = begins[-1]
        if self._full is None:
            self._full = {}
        if slot not in self._full or self._full[slot][0].shape[0] != n_total:
            self._full[slot] = (torch.empty(n_total, dtype=torch.int64, device=self.device),
                                torch.empty(n_total, dtype=torch.float64, device=self.device))
        full_spins, full_psi = self._full[slot]
        lib().asp_set_gather_mode(self.MODES[self.mode])
        check(lib().asp_gather_index(operator.handle, self.world, self.rank, ffi.new("uint64_t[]", begins), self._shard_spins,
                                     self._shard_psi, ffi.cast("uint64_t const *", self._own), self.epoch,
                                     ptr(full_spins, "uint64_t *"), ptr(full_psi, "double *"), num_rows,
                                     ptr(workspace, "void *"), workspace.numel(), stream()))
        return full_spins, full_psi

    def gather_blocks(self, shard_begin, slot: int = 0, engine: str = "ce"):
        """X1 without the index (asp_gather_blocks): -> (full_spins, full_psi).  engine "ce" (default): the copy
        engines pull the blocks -- no SM involved, so it overlaps completely with an extraction on another stream;
        "tma": one thread per CTA drives bulk copies (faster alone, but measured to slow a concurrent extraction
        down by more than it gains)."""
        from ._lib import check, ffi, lib, ptr, stream

        begins = [int(b) for b in shard_begin]
        assert len(begins) == self.world + 1 and begins[0] == 0
        assert all(begins[q + 1] - begins[q] <= self.capacity for q in range(self.world))
        n_total = begins[-1]
        if self._full is None:
            self._full = {}
        if slot not in self._full or self._full[slot][0].shape[0] != n_total:
            self._full[slot] = (torch.empty(n_total, dtype=torch.int64, device=self.device),
                                torch.empty(n_total, dtype=torch.float64, device=self.device))
        full_spins, full_psi = self._full[slot]
        lib().asp_set_gather_mode(0 if engine == "ce" else 2)
        check(lib().asp_gather_blocks(self.world, self.rank, ffi.new("uint64_t[]", begins), self._shard_spins, self._shard_psi,
                                      ffi.cast("uint64_t const *", self._own), self.epoch, ptr(full_spins, "uint64_t *"),
                                      ptr(full_psi, "double *"), stream()))
        return full_spins, full_psi

    def close(self):
        """Unmap the peers' blocks, then (after a host barrier: nobody may still map it) free ours."""
        from ._lib import ffi, lib

        if self._own == 0:
            return
        torch.cuda.synchronize()
        self.spins = self.psi = self._raw = self._full = None
        for address in self._opened:
            lib().asp_peer_close(ffi.cast("void *", address))
        self._opened = []
        barrier()
        lib().asp_peer_free(ffi.cast("void *", self._own))
        self._own = 0


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
