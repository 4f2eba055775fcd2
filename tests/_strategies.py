"""hypothesis strategies shared by the CPU and GPU property tests: random two-site operators (reference YAML schema)
and random sampled subsets with hits, misses, duplicate candidates, cancelling couplings and zero amplitudes."""
import numpy as np
from hypothesis import strategies as st

from oracle.operator_np import OperatorNP

HEIS = [[1, 0, 0, 0], [0, -1, 2, 0], [0, 2, -1, 0], [0, 0, 0, 1]]


@st.composite
def symmetric_matrix(draw):
    kind = draw(st.integers(0, 2))
    if kind == 0:
        scale = draw(st.sampled_from([1.0, -1.0, 0.5, 1.1, -0.25]))
        return (scale * np.array(HEIS, dtype=np.float64)).tolist()
    m = np.zeros((4, 4))
    for a in range(4):
        for b in range(a, 4):
            m[a, b] = m[b, a] = draw(st.sampled_from([0.0, 0.0, 1.0, -1.0, 2.0, 0.5]))
    return m.tolist()


@st.composite
def problems(draw):
    n_spins = draw(st.integers(2, 40))
    hw = draw(st.one_of(st.none(), st.integers(0, n_spins)))
    terms = []
    for _ in range(draw(st.integers(1, 3))):
        sites = []
        for _ in range(draw(st.integers(1, 6))):
            i = draw(st.integers(0, n_spins - 1))
            j = draw(st.integers(0, n_spins - 2))
            sites.append([i, j if j < i else j + 1])
        terms.append({"matrix": draw(symmetric_matrix()), "sites": sites})
    cfg = {"basis": {"number_spins": n_spins, "hamming_weight": hw, "symmetries": []},
           "hamiltonian": {"name": "random", "terms": terms}}
    seed = draw(st.integers(0, 2 ** 31 - 1))
    m = draw(st.sampled_from([0, 1, 2, 5, 33, 200]))
    return cfg, seed, m


def random_subset(cfg, seed, m):
    rng = np.random.default_rng(seed)
    n_spins, hw = cfg["basis"]["number_spins"], cfg["basis"]["hamming_weight"]
    if hw is None:
        words = rng.integers(0, 1 << n_spins, size=m, dtype=np.uint64) if m else np.zeros(0, dtype=np.uint64)
    else:
        words = np.zeros(m, dtype=np.uint64)
        for k in range(m):
            up = rng.choice(n_spins, size=hw, replace=False)
            words[k] = np.bitwise_or.reduce(np.uint64(1) << up.astype(np.uint64)) if hw else np.uint64(0)
    # a cluster: add the neighbours of a few states so that there ARE hits
    spins = np.unique(words)
    if spins.shape[0]:
        shell, _, _ = OperatorNP.from_config(cfg).apply_u64(spins[: max(1, spins.shape[0] // 4)])
        spins = np.unique(np.concatenate([spins, shell[rng.random(shell.shape[0]) < 0.7]]))
    psi = rng.standard_normal(spins.shape[0])
    psi[rng.random(spins.shape[0]) < 0.05] = 0.0
    return spins, psi


