#!/usr/bin/env python
"""Host-link probe behind the end-to-end numbers at N > 1: every rank copies a pinned buffer device -> host (and host ->
device) at the same time as all the others; per-rank and total GB/s, with the NUMA node and CPU affinity each process
runs on.  Shows whether the end-to-end efficiency of bench.py at 4 and 8 GPUs is the box (shared host links / one NUMA
node) or the code.

    torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/probe_pcie.py [--mbytes 600]
"""
import argparse
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--mbytes", type=int, default=600)
    p.add_argument("--reps", type=int, default=5)
    args = p.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.mbytes * 1_000_000 // 8
    d = torch.zeros(n, dtype=torch.float64, device="cuda")
    h = torch.zeros(n, dtype=torch.float64).pin_memory()
    out = {}
    for name, dst, src in (("d2h", h, d), ("h2d", d, h)):
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.reps):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        out[name] = args.reps * n * 8 / (time.perf_counter() - t0) / 1e9
    # both directions at once (full duplex)
    h2 = torch.zeros(n, dtype=torch.float64).pin_memory()
    d2 = torch.zeros(n, dtype=torch.float64, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        with torch.cuda.stream(s1):
            h.copy_(d, non_blocking=True)
        with torch.cuda.stream(s2):
            d2.copy_(h2, non_blocking=True)
    torch.cuda.synchronize()
    out["duplex_each_way"] = args.reps * n * 8 / (time.perf_counter() - t0) / 1e9
    try:
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        node = open("/sys/bus/pci/devices/0000:%02x:00.0/numa_node" % bus).read().strip()
    except Exception:  # noqa: BLE001
        node = "?"
    out.update({"rank": rank, "gpu_numa_node": node, "cpus": sorted(os.sched_getaffinity(0))[:4] + ["..."], "ncpus": len(os.sched_getaffinity(0))})
    every = [None] * world
    if world > 1:
        dist.all_gather_object(every, out)
    else:
        every = [out]
    if rank == 0:
        print(json.dumps({"world": world, "mbytes": args.mbytes, "total_d2h_gbs": sum(e["d2h"] for e in every), "total_h2d_gbs": sum(e["h2d"] for e in every),
                          "total_duplex_each_way_gbs": sum(e["duplex_each_way"] for e in every), "per_rank": every}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
