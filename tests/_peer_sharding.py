"""Run under torchrun with >= 2 ranks, one GPU each: the exchange step X1 over NVLink peer memory
(distributed.PeerBasis: CUDA-IPC mapped row blocks, epoch flags, asp_gather_index) against the NCCL
all-gather + asp_extract_csr path and against the unsharded single-GPU build -- bit for bit, over
several epochs in which every rank REWRITES its block (so stale reads or a broken flag protocol
show up as wrong keys)."""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))

import annealing_sign_problem_b200 as asp  # noqa: E402
from annealing_sign_problem_b200 import common, synthetic  # noqa: E402
from annealing_sign_problem_b200 import distributed as D  # noqa: E402
from annealing_sign_problem_b200._lib import lib  # noqa: E402


def main():
    rank, world, local = D.init_from_env()
    assert world >= 2
    dev = torch.device("cuda", local)
    cfg = asp.ls.load_config(asp.ls.system_path("heisenberg_kagome_36"))
    cfg["basis"]["symmetries"], cfg["basis"]["spin_inversion"] = [], None
    op = asp.ls.Operator.load_from_yaml(cfg["hamiltonian"], asp.ls.SpinBasis.load_from_yaml(cfg["basis"]))
    n_max = 400_000
    pb = D.PeerBasis(n_max // world + 64, dev)
    for epoch, (n_want, seed) in enumerate([(200_001, 3), (399_990, 4), (77_777, 5), (300_000, 6)]):
        import torch.distributed as dist

        spins = synthetic.cluster_closed_states(op, n_want, seed, dev)
        size = torch.tensor([spins.shape[0]], dtype=torch.int64, device=dev)
        dist.broadcast(size, src=0)
        n = int(size[0])
        if rank != 0:
            spins = torch.empty(n, dtype=torch.int64, device=dev)
        dist.broadcast(spins, src=0)  # rank 0's basis is THE basis
        psi = synthetic.synthetic_amplitudes(n, seed, device=dev)
        dist.broadcast(psi, src=0)
        bounds = [D.block(n, r, world)[0] for r in range(world)] + [n]
        begin, rows = bounds[rank], bounds[rank + 1] - bounds[rank]
        # the reference result: unsharded build of my rows on this GPU
        ref = common.extract_csr_device(op, spins, psi, begin, rows)
        # NCCL path
        full_s = D.all_gather_blocks(spins[begin:begin + rows].clone(), n)
        full_p = D.all_gather_blocks(psi[begin:begin + rows].clone(), n)
        assert torch.equal(full_s, spins) and torch.equal(full_p, psi)
        # peer-memory path
        pb.mode = ["ce", "sm", "tma", "tma"][epoch]  # every implementation of asp_gather_index
        pb.begin_epoch()
        pb.spins[:rows] = spins[begin:begin + rows]
        pb.psi[:rows] = psi[begin:begin + rows]
        pb.spins[rows:] = -1  # poison the tail: nothing past the block may be read
        pb.publish()
        need = int(lib().asp_extract_csr_workspace_bytes(op.handle, n, rows))
        workspace = torch.empty(need, dtype=torch.uint8, device=dev)
        got_s, got_p = pb.gather_index(op, bounds, rows, workspace)
        pb.release()
        assert torch.equal(got_s, spins), "epoch %d rank %d: gathered keys differ" % (epoch, rank)
        assert torch.equal(got_p, psi), "epoch %d rank %d: gathered amplitudes differ" % (epoch, rank)
        out = common.extract_csr_indexed_device(op, got_s, got_p, begin, rows, workspace, capacity=int(ref[1].numel()))
        for a, b in zip(out, ref):
            assert torch.equal(a, b), "epoch %d rank %d: CSR differs" % (epoch, rank)
        # copy-engine gather without index (asp_gather_blocks) + the ordinary extraction: what a pipeline over
        # independent extractions runs on its exchange stream / compute stream
        pb.begin_epoch()
        pb.publish()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            cp_s, cp_p = pb.gather_blocks(bounds, slot=1, engine="ce" if epoch % 2 else "tma")
            pb.release()
        torch.cuda.current_stream().wait_stream(side)
        assert torch.equal(cp_s, spins) and torch.equal(cp_p, psi), "epoch %d rank %d: gather_blocks differs" % (epoch, rank)
        out = common.extract_csr_device(op, cp_s, cp_p, begin, rows)
        for a, b in zip(out, ref):
            assert torch.equal(a, b), "epoch %d rank %d: CSR after gather_blocks differs" % (epoch, rank)
    pb.close()
    D.barrier()
    if rank == 0:
        print("PEER_OK")
    import torch.distributed as dist

    dist.destroy_process_group()


if __name__ == "__main__":
    main()
