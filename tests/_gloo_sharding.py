"""Run under torchrun with 2 ranks on CPU (gloo): checks the partition + exchange plumbing
of annealing-sign-problem_b200/distributed.py (no GPU, no kernels)."""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))

from annealing_sign_problem_b200 import distributed as D  # noqa: E402
from annealing_sign_problem_b200 import synthetic  # noqa: E402


def main():
    rank, world, _ = D.init_from_env()
    assert world == 2
    n = 10001
    spins = synthetic.random_sector_states(20, 10, n, seed=3)  # same on every rank
    psi = synthetic.synthetic_amplitudes(n, seed=3)
    begin, count = D.block(n, rank, world)
    assert sum(D.block(n, r, world)[1] for r in range(world)) == n
    full_spins = D.all_gather_blocks(spins[begin:begin + count].clone(), n)
    full_psi = D.all_gather_blocks(psi[begin:begin + count].clone(), n)
    assert torch.equal(full_spins, spins) and torch.equal(full_psi, psi)
    off, total = D.exclusive_offset(100 + 7 * rank, "cpu")
    assert total == 207 and off == (0 if rank == 0 else 100)
    bits = torch.full((5,), rank + 10, dtype=torch.int64)
    e, b, owner = D.reduce_best(-1.0 - rank, bits)
    assert owner == 1 and e == -2.0 and torch.equal(b, torch.full((5,), 11, dtype=torch.int64))
    e, b, owner = D.reduce_best(-3.0, bits)  # tie -> lowest rank
    assert owner == 0 and torch.equal(b, torch.full((5,), 10, dtype=torch.int64))
    assert D.max_over_ranks(float(rank), "cpu") == 1.0 and D.sum_over_ranks(1.5, "cpu") == 3.0
    # peer memory cannot exist without a CUDA device: EVERY rank must learn that together (the error
    # travels through the same collectives a successful set-up uses), so callers can switch to the
    # all-gather exchange in step
    try:
        D.PeerBasis(1000, device=torch.device("cpu"))
        raise SystemExit("PeerBasis must not come up without CUDA")
    except D.PeerMemoryUnavailable as exc:
        assert "rank 0" in str(exc) and "rank 1" in str(exc), str(exc)
    D.barrier()
    if rank == 0:
        print("SHARDING_OK")


if __name__ == "__main__":
    main()
