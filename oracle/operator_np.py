"""numpy restatement of the neighbour generator the reference calls (TEST INFRASTRUCTURE).

The reference obtains the Hamiltonian's off-diagonal images from the third-party C
library ``lattice_symmetries`` (pinned ``=0.8.3`` in conda-annealing.yml:9, ``>=0.8.2``
in setup.py:33).  It is NOT vendored under /root/reference, so this file restates the
library's *published* semantics and anchors them on the reference's own call sites:

* ``Operator.batched_apply(x[m,8] u64) -> (spins[T,8] u64, coeffs[T] c128, counts[m])``
  -- annealing_sign_problem/common.py:85-106 (``_batched_apply``), :96 the call itself.
* ``hamiltonian.basis.number_spins`` -- common.py:86.
* The diagonal term is returned as one extra "neighbour" ``s' = s`` (common.py:445-446,
  :956-957 strip it with ``setdiag(0)``; the KAT ``s^T J s == E0`` at
  experiments/full_hilbert_space.py:143-145 needs it).
* YAML schema ``basis: {number_spins, hamming_weight, spin_inversion, symmetries:
  [{permutation, sector}]}``, ``hamiltonian: {terms: [{matrix 4x4, sites [[i,j],..]}]}``
  -- physical_systems/*.yaml, loaded by common.py:782-787 (``load_hamiltonian``).

PARITY UNPINNED at this boundary: no reference test fixes the order of the returned
neighbours, where the diagonal sits, or whether duplicates are merged.  Everything that
consumes this output is therefore compared in canonical (row-sorted, merged) form.

Conventions chosen here (all shipped matrices are symmetric under site swap and
transposition, so the reference cannot distinguish them):

* local two-site state index ``a = 2*bit(s, sites[0]) + bit(s, sites[1])``;
* ``coeff(s -> s') = matrix[a][a']`` (row = the basis state the row of J belongs to);
* per input state: off-diagonal images in (term, bond, a') order with zero
  coefficients skipped, then ONE diagonal entry (always emitted, even when 0);
* symmetrised bases (``spin_inversion`` and/or permutation generators, sector 0 or
  real characters): image -> orbit representative ``min_g g(s')``, coefficient scaled by
  ``chi(g) * norm(rep)/norm(s)`` with ``norm(x)^2 = sum_g chi(g)[g x == x] / |G|``.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

U64 = np.uint64


def load_config(path: str) -> dict:
    """Parse a physical-system file (reference YAML schema; JSON with the same schema)."""
    with open(path, "r") as f:
        if path.endswith(".json"):
            return json.load(f)
        import yaml

        return yaml.load(f, Loader=yaml.SafeLoader)


def _popcount(x: np.ndarray) -> np.ndarray:
    x = x.astype(U64)
    return np.bitwise_count(x).astype(np.int64)


def _permute_bits(x: np.ndarray, perm: Sequence[int]) -> np.ndarray:
    """new bit k = old bit perm[k]."""
    out = np.zeros_like(x, dtype=U64)
    for k, src in enumerate(perm):
        out |= ((x >> U64(src)) & U64(1)) << U64(k)
    return out


def _close_group(generators: List[Tuple[Tuple[int, ...], int]], n: int):
    """Closure of permutation generators -> list of (perm, character) with real characters.

    ``sector`` k of a generator with periodicity p has character exp(-2 pi i k / p); only
    k = 0 (chi = +1) and 2k = p (chi = -1) give the real coefficients the reference
    requires (common.py:97-98 raises otherwise).
    """
    identity = tuple(range(n))
    if not generators:
        return [(identity, 1.0)]

    def compose(p, q):  # apply q first, then p: new[k] = old[q[p[k]]]
        return tuple(q[p[k]] for k in range(n))

    gens = []
    for perm, sector in generators:
        perm = tuple(int(v) for v in perm)
        period, cur = 1, perm
        while cur != identity:
            cur = compose(perm, cur)
            period += 1
        if sector == 0:
            chi = 1.0
        elif 2 * sector == period:
            chi = -1.0
        else:
            raise NotImplementedError("complex characters are outside the reference's hot path")
        gens.append((perm, chi))
    group = {identity: 1.0}
    frontier = [identity]
    while frontier:
        new = []
        for g in frontier:
            for p, chi in gens:
                h = compose(p, g)
                c = group[g] * chi
                if h not in group:
                    group[h] = c
                    new.append(h)
                elif group[h] != c:
                    raise ValueError("inconsistent sectors: the chosen characters do not form a representation")
        frontier = new
    return sorted(group.items())


@dataclass
class SpinBasisNP:
    number_spins: int
    hamming_weight: Optional[int] = None
    spin_inversion: Optional[int] = None
    symmetries: List[dict] = field(default_factory=list)

    def __post_init__(self):
        gens = [(tuple(s["permutation"]), int(s.get("sector", 0))) for s in (self.symmetries or [])]
        self.group = _close_group(gens, self.number_spins)
        self.mask = U64((1 << self.number_spins) - 1) if self.number_spins < 64 else U64(0xFFFFFFFFFFFFFFFF)
        self._states = None

    @classmethod
    def from_config(cls, cfg: dict) -> "SpinBasisNP":
        return cls(
            number_spins=int(cfg["number_spins"]),
            hamming_weight=cfg.get("hamming_weight"),
            spin_inversion=cfg.get("spin_inversion"),
            symmetries=cfg.get("symmetries") or [],
        )

    @property
    def is_symmetrised(self) -> bool:
        return bool(self.spin_inversion) or len(self.group) > 1

    @property
    def group_order(self) -> int:
        return len(self.group) * (2 if self.spin_inversion else 1)

    def state_info(self, x: np.ndarray):
        """-> (representative u64, character f64, norm f64) per state (ls_get_state_info)."""
        x = np.asarray(x, dtype=U64)
        rep = x.copy()
        chi_rep = np.ones(x.shape, dtype=np.float64)
        stab = np.zeros(x.shape, dtype=np.float64)
        inversions = [(False, 1.0)]
        if self.spin_inversion:
            inversions.append((True, float(self.spin_inversion)))
        for perm, chi in self.group:
            y0 = _permute_bits(x, perm) if perm != tuple(range(self.number_spins)) else x
            for flip, chi_f in inversions:
                y = (~y0) & self.mask if flip else y0
                c = chi * chi_f
                stab += np.where(y == x, c, 0.0)
                better = y < rep
                rep = np.where(better, y, rep)
                chi_rep = np.where(better, c, chi_rep)
        norm = np.sqrt(np.maximum(stab, 0.0) / self.group_order)
        return rep, chi_rep, norm

    def build(self):
        """Enumerate the representatives of the whole sector (small systems only)."""
        n = self.number_spins
        if n > 26:
            raise ValueError("full-basis enumeration is for the small (<=26 spin) systems only")
        x = np.arange(1 << n, dtype=U64)
        if self.hamming_weight is not None:
            x = x[_popcount(x) == int(self.hamming_weight)]
        if self.is_symmetrised:
            rep, _, norm = self.state_info(x)
            x = x[(rep == x) & (norm > 0)]
        self._states = np.ascontiguousarray(x)
        return self

    @property
    def states(self) -> np.ndarray:
        if self._states is None:
            self.build()
        return self._states

    @property
    def number_states(self) -> int:
        return int(self.states.shape[0])


class OperatorNP:
    """Duck-type of ``ls.Operator`` as the reference uses it (common.py:85-106)."""

    def __init__(self, basis: SpinBasisNP, terms: List[dict]):
        self.basis = basis
        self.terms = []
        for t in terms:
            m = np.asarray(t["matrix"], dtype=np.float64)
            if m.shape != (4, 4):
                raise ValueError("only two-site terms (4x4 matrices) occur on the reference's hot path")
            sites = [(int(a), int(b)) for a, b in t["sites"]]
            self.terms.append((m, sites))

    @classmethod
    def from_config(cls, cfg: dict) -> "OperatorNP":
        basis = SpinBasisNP.from_config(cfg["basis"])
        return cls(basis, cfg["hamiltonian"]["terms"])

    @classmethod
    def load(cls, path: str) -> "OperatorNP":
        return cls.from_config(load_config(path))

    # -- the call the reference makes ------------------------------------------------
    def batched_apply(self, x):
        x = np.asarray(x, dtype=U64)
        if x.ndim == 2:
            x = x[:, 0]
        spins, coeffs, counts = self.apply_u64(np.ascontiguousarray(x))
        out = np.zeros((spins.shape[0], 8), dtype=U64)
        out[:, 0] = spins
        return out, coeffs.astype(np.complex128), counts

    def apply_u64(self, x: np.ndarray):
        """-> (other_spins[T] u64, other_coeffs[T] f64, other_counts[m] i64), row-major."""
        m = x.shape[0]
        cols_s, cols_c, cols_ok = [], [], []
        diag = np.zeros(m, dtype=np.float64)
        for mat, sites in self.terms:
            for (i, j) in sites:
                bi = (x >> U64(i)) & U64(1)
                bj = (x >> U64(j)) & U64(1)
                a = (bi * U64(2) + bj).astype(np.int64)
                diag = diag + mat[a, a]
                cleared = x & ~((U64(1) << U64(i)) | (U64(1) << U64(j)))
                for ap in range(4):
                    c = mat[a, ap]
                    ok = (c != 0.0) & (a != ap)
                    if not ok.any():
                        continue
                    img = cleared | (U64(ap >> 1) << U64(i)) | (U64(ap & 1) << U64(j))
                    cols_s.append(img)
                    cols_c.append(c)
                    cols_ok.append(ok)
        cols_s.append(x)
        cols_c.append(diag)
        cols_ok.append(np.ones(m, dtype=bool))
        S = np.stack(cols_s, axis=1)
        C = np.stack(cols_c, axis=1)
        OK = np.stack(cols_ok, axis=1)
        if self.basis.is_symmetrised:
            _, _, norm_x = self.basis.state_info(x)
            flat = S[OK]
            rep, chi, norm = self.basis.state_info(flat)
            rows = np.broadcast_to(np.arange(m)[:, None], S.shape)[OK]
            scale = chi * norm / norm_x[rows]
            S = S.copy()
            C = C.copy()
            S[OK] = rep
            C[OK] = C[OK] * scale
            # images outside the symmetry sector (norm 0) carry no weight
            dead = np.zeros_like(OK)
            dead[OK] = norm == 0
            OK = OK & ~dead
        counts = OK.sum(axis=1).astype(np.int64)
        return np.ascontiguousarray(S[OK]), np.ascontiguousarray(C[OK]), counts

    # -- dense/sparse matrix in the (possibly symmetrised) basis: used for ED KATs ------
    def to_sparse(self):
        import scipy.sparse

        states = self.basis.states
        n = states.shape[0]
        spins, coeffs, counts = self.apply_u64(states)
        rows = np.repeat(np.arange(n), counts)
        cols = np.searchsorted(states, spins)
        cols = np.clip(cols, 0, n - 1)
        found = states[cols] == spins
        return scipy.sparse.csr_matrix((coeffs[found], (rows[found], cols[found])), shape=(n, n))


def ground_state(op: OperatorNP, k: int = 2, seed: int = 0):
    """(E0, psi0, E1) by Lanczos; for the 16/18-spin known-answer tests."""
    import scipy.sparse.linalg

    h = op.to_sparse()
    rng = np.random.default_rng(seed)
    v0 = rng.standard_normal(h.shape[0])
    w, v = scipy.sparse.linalg.eigsh(h, k=k, which="SA", tol=1e-13, v0=v0)
    order = np.argsort(w)
    return float(w[order[0]]), np.ascontiguousarray(v[:, order[0]]), float(w[order[1]])


SYSTEMS_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "annealing-sign-problem_b200", "systems")


def system_path(name: str) -> str:
    """Committed JSON restatement of physical_systems/<name>.yaml (tests/golden/make_systems.py)."""
    return os.path.join(SYSTEMS_DIR, name + ".json")
