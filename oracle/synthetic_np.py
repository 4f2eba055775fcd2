"""numpy twin of annealing-sign-problem_b200/synthetic.py for the CPU legs of bench.py (test infrastructure,
like everything under oracle/): the reference arm must not touch the product's CUDA code, not even to make
its inputs.  Same construction as synthetic.cluster_closed_states -- random seeds of the U(1) sector, their
full batched_apply shell (every neighbour of a seed is present), a random part of the second shell -- the
shape make_hamiltonian_extension (annealing_sign_problem/common.py:516-522) produces; different random stream."""
import numpy as np

U64 = np.uint64


def random_sector_states(number_spins: int, hamming_weight: int, n: int, rng) -> np.ndarray:
    """n distinct uniformly random words with `hamming_weight` bits among `number_spins`, ascending."""
    out = np.zeros(0, dtype=U64)
    while out.shape[0] < n:
        m = int((n - out.shape[0]) * 1.05) + 64
        up = np.argsort(rng.random((m, number_spins)), axis=1)[:, :hamming_weight].astype(U64)
        words = np.bitwise_or.reduce(U64(1) << up, axis=1) if hamming_weight else np.zeros(m, dtype=U64)
        out = np.unique(np.concatenate([out, words]))
    if out.shape[0] > n:
        out = np.sort(rng.choice(out, size=n, replace=False))
    return out


def cluster_closed_states(op_np, n: int, seed: int, interior_fraction: float = 1.0 / 3.0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    basis = op_np.basis
    bonds = sum(len(sites) for _, sites in op_np.terms)
    d = max(2.0, (2 * bonds + 1) / 4.0)  # candidates per row, as in synthetic.cluster_closed_states
    m = max(1, int(n * interior_fraction / (d + 1.0)))
    seeds = random_sector_states(basis.number_spins, int(basis.hamming_weight), m, rng)
    shell1, _, _ = op_np.apply_u64(seeds)
    interior = np.unique(np.concatenate([seeds, shell1]))
    if interior.shape[0] >= n:
        return np.sort(rng.choice(interior, size=n, replace=False))
    shell2, _, _ = op_np.apply_u64(interior)
    boundary = np.setdiff1d(np.unique(shell2), interior, assume_unique=True)
    need = n - interior.shape[0]
    if boundary.shape[0] > need:
        boundary = rng.choice(boundary, size=need, replace=False)
    pool = np.unique(np.concatenate([interior, boundary]))
    if pool.shape[0] < n:
        extra = np.setdiff1d(random_sector_states(basis.number_spins, int(basis.hamming_weight), n, rng), pool, assume_unique=True)
        pool = np.unique(np.concatenate([pool, extra[: n - pool.shape[0]]]))
    return pool


def synthetic_amplitudes(n: int, seed: int, sigma: float = 2.0) -> np.ndarray:
    """psi_i = +-exp(sigma z_i), z ~ N(0,1), uniform sign, L2-normalised (common.py:181)."""
    rng = np.random.default_rng(seed)
    psi = np.where(rng.random(n) < 0.5, -1.0, 1.0) * np.exp(sigma * rng.standard_normal(n))
    return psi / np.linalg.norm(psi)
