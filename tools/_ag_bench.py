import os, sys, time
sys.path.insert(0, "/root/repo")
import torch, torch.distributed as dist
from annealing_sign_problem_b200 import distributed as D
rank, world, local = D.init_from_env()
dev = torch.device("cuda", local)
n = 10_000_000
mine_s = torch.arange(n, dtype=torch.int64, device=dev) + rank * n
mine_p = torch.rand(n, dtype=torch.float64, device=dev)
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
out_s = torch.empty(world * n, dtype=torch.int64, device=dev)
out_p = torch.empty(world * n, dtype=torch.float64, device=dev)
t1 = timeit(lambda: dist.all_gather_into_tensor(out_s, mine_s))
t2 = timeit(lambda: (dist.all_gather_into_tensor(out_s, mine_s), dist.all_gather_into_tensor(out_p, mine_p)))
t3 = timeit(lambda: (D.all_gather_blocks(mine_s, world * n), D.all_gather_blocks(mine_p, world * n)))
pack = torch.empty(2 * n, dtype=torch.int64, device=dev)
out2 = torch.empty(world * 2 * n, dtype=torch.int64, device=dev)
def packed():
    pack[:n] = mine_s; pack[n:] = mine_p.view(torch.int64)
    dist.all_gather_into_tensor(out2, pack)
t4 = timeit(packed)
if rank == 0:
    print("world %d: one all-gather 80MB/rank %.3f ms; two (preallocated) %.3f ms; two via all_gather_blocks %.3f ms; one packed 160MB/rank incl. pack copies %.3f ms" % (world, t1, t2, t3, t4))
dist.destroy_process_group()
