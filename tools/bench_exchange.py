"""X1 alone, ranks aligned by a host barrier before every sample: device time of
gather_index_kernel (pull over NVLink peer memory + private copy + index) against the two NCCL
all-gathers + the local index pass it replaces.  Run under torchrun, one rank per GPU:

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_exchange.py [rows_per_gpu]
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))

import annealing_sign_problem_b200 as asp  # noqa: E402
from annealing_sign_problem_b200 import common, synthetic  # noqa: E402
from annealing_sign_problem_b200 import distributed as D  # noqa: E402
from annealing_sign_problem_b200._lib import ffi, lib  # noqa: E402


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    rank, world, local = D.init_from_env()
    dev = torch.device("cuda", local)
    cfg = asp.ls.load_config(asp.ls.system_path("heisenberg_kagome_36"))
    cfg["basis"]["symmetries"], cfg["basis"]["spin_inversion"] = [], None
    op = asp.ls.Operator.load_from_yaml(cfg["hamiltonian"], asp.ls.SpinBasis.load_from_yaml(cfg["basis"]))
    n = rows * world
    # any sorted unique keys do: uniform sector states, same on every rank
    spins = synthetic.random_sector_states(36, 18, n, 5, dev)
    dist.broadcast(spins, src=0)
    psi = synthetic.synthetic_amplitudes(n, 5, device=dev)
    bounds = [D.block(n, r, world)[0] for r in range(world)] + [n]
    begin, mine = bounds[rank], bounds[rank + 1] - bounds[rank]
    pb = D.PeerBasis(rows + 64, dev)
    pb.spins[:mine] = spins[begin:begin + mine]
    pb.psi[:mine] = psi[begin:begin + mine]
    my_s, my_p = spins[begin:begin + mine].clone(), psi[begin:begin + mine].clone()
    need = int(lib().asp_extract_csr_workspace_bytes(op.handle, n, mine))
    ws = torch.empty(need, dtype=torch.uint8, device=dev)
    out = {"peer": [], "peer_sm": [], "peer_tma": [], "nccl": [], "nccl_index": []}
    for it in range(25):
        for mode, key in (("ce", "peer"), ("sm", "peer_sm"), ("tma", "peer_tma")):
            pb.mode = mode
            dist.barrier()
            torch.cuda.synchronize()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            pb.begin_epoch()
            pb.publish()
            e[0].record()
            fs, fp = pb.gather_index(op, bounds, mine, ws)
            e[1].record()
            pb.release()
            torch.cuda.synchronize()
            out[key].append(e[0].elapsed_time(e[1]))
            if it == 0:
                assert torch.equal(fs, spins) and torch.equal(fp, psi)
        lib().asp_set_gather_mode(1)
        dist.barrier()
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        gs = D.all_gather_blocks(my_s, n)
        gp = D.all_gather_blocks(my_p, n)
        e[1].record()
        # the local pass that follows an all-gather: one block = the whole array (copy + index)
        common.check(lib().asp_gather_index(op.handle, 1, 0, ffi.new("uint64_t[]", [0, n]),
                                            ffi.new("uint64_t const *[]", [common.ptr(gs, "uint64_t const *")]),
                                            ffi.new("double const *[]", [common.ptr(gp, "double const *")]), ffi.NULL, 0,
                                            common.ptr(fs, "uint64_t *"), common.ptr(fp, "double *"), mine, common.ptr(ws, "void *"),
                                            need, common.stream()))
        e[2].record()
        torch.cuda.synchronize()
        out["nccl"].append(e[0].elapsed_time(e[1]))
        out["nccl_index"].append(e[1].elapsed_time(e[2]))
    pb.close()
    res = {k: (float(np.median(v[3:])), float(np.min(v[3:]))) for k, v in out.items()}
    remote_gb = (n - mine) * 16 / 1e9
    every = [None] * world
    dist.all_gather_object(every, res)
    if rank == 0:
        for r, x in enumerate(every):
            print("rank %d: copy-engine pull + block index median %.3f ms (min %.3f) = %.0f GB/s pulled | SM gather+index kernel %.3f ms "
                  "(min %.3f) = %.0f GB/s | TMA gather+index kernel %.3f ms (min %.3f) = %.0f GB/s | nccl all-gather x2 %.3f ms (min %.3f) + local copy+index %.3f ms" % (
                      r, x["peer"][0], x["peer"][1], remote_gb / (x["peer"][0] * 1e-3), x["peer_sm"][0], x["peer_sm"][1],
                      remote_gb / (x["peer_sm"][0] * 1e-3), x["peer_tma"][0], x["peer_tma"][1], remote_gb / (x["peer_tma"][0] * 1e-3), x["nccl"][0], x["nccl"][1], x["nccl_index"][0]))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
