"""NVLink peer-memory probe (one rank per GPU under torchrun): how fast can a rank PULL from / PUSH to
peer-mapped buffers, with the copy engines (cudaMemcpyAsync through torch.copy_) and with SM
loads/stores (a torch elementwise kernel), to one peer and to all peers at once.  Guides the design
of the X1 exchange."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))

from annealing_sign_problem_b200 import distributed as D  # noqa: E402
from annealing_sign_problem_b200._lib import check, ffi, lib  # noqa: E402


def main():
    rank, world, local = D.init_from_env()
    dev = torch.device("cuda", local)
    nbytes = 160_000_000
    own = ffi.new("void **")
    handle = ffi.new("unsigned char[64]")
    check(lib().asp_peer_alloc(nbytes * world, own, handle))
    every = [None] * world
    dist.all_gather_object(every, bytes(ffi.buffer(handle, 64)))
    views = []
    for q, h in enumerate(every):
        if q == rank:
            address = int(ffi.cast("uintptr_t", own[0]))
        else:
            out = ffi.new("void **")
            check(lib().asp_peer_open(ffi.from_buffer("unsigned char[]", h), out))
            address = int(ffi.cast("uintptr_t", out[0]))
        views.append(torch.as_tensor(D._RawDeviceMemory(address, nbytes * world), device=dev).view(torch.int64))
    words = nbytes // 8
    src = torch.arange(words, dtype=torch.int64, device=dev)
    local_dst = torch.empty(words * world, dtype=torch.int64, device=dev)
    streams = [torch.cuda.Stream(device=dev) for _ in range(world)]
    peers = [(rank + 1 + k) % world for k in range(world - 1)]

    def timed(fn, reps=5):
        best = 1e9
        for _ in range(reps):
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best

    def fan(op):
        def run():
            cur = torch.cuda.current_stream()
            for k, p in enumerate(peers):
                streams[k].wait_stream(cur)
                with torch.cuda.stream(streams[k]):
                    op(p)
                cur.wait_stream(streams[k])
        return run

    def seq(op):
        def run():
            for p in peers:
                op(p)
        return run

    mine = slice(rank * words, (rank + 1) * words)
    ops = {
        "CE push": lambda p: views[p][mine].copy_(src),
        "CE pull": lambda p: local_dst[p * words:(p + 1) * words].copy_(views[p][p * words:(p + 1) * words]),
        "SM push": lambda p: torch.add(src, 1, out=views[p][mine]),
        "SM pull": lambda p: torch.add(views[p][p * words:(p + 1) * words], 1, out=local_dst[p * words:(p + 1) * words]),
    }
    rows = []
    for name, op in ops.items():
        one = timed(lambda: op(peers[0]))
        allseq = timed(seq(op))
        allfan = timed(fan(op))
        rows.append("%s: one peer %.0f GB/s | all %d peers one after the other %.0f GB/s | all at once (streams) %.0f GB/s" % (
            name, nbytes / one / 1e6, world - 1, nbytes * (world - 1) / allseq / 1e6, nbytes * (world - 1) / allfan / 1e6))
    gathered = [None] * world
    dist.all_gather_object(gathered, rows)
    if rank == 0:
        for r in (0, world - 1):
            print("rank %d\n  " % r + "\n  ".join(gathered[r]))
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
