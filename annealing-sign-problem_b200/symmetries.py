"""Host-side mirror of the ``lattice_symmetries`` objects the reference's hot path touches
(``ls.SpinBasis`` / ``ls.Operator``: annealing_sign_problem/common.py:782-787 load them from
YAML, :86 reads ``basis.number_spins``, :96 calls ``batched_apply``).  The bond list is
compiled once into an ``asp_operator`` (include/asp_b200.h) so that neighbour generation
runs fused with the search on the GPU instead of materialising candidate lists.
"""
from __future__ import annotations

import json
import os
from typing import List, Optional, Sequence

import numpy as np
import torch

from ._lib import check, ffi, lib, ptr, require_cuda, stream

SYSTEMS_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "systems")


def load_config(path: str) -> dict:
    """Reference YAML schema (physical_systems/*.yaml) or the same schema as JSON."""
    with open(path, "r") as f:
        if path.endswith(".json"):
            return json.load(f)
        import yaml

        return yaml.load(f, Loader=yaml.SafeLoader)


def system_path(name: str) -> str:
    return os.path.join(SYSTEMS_DIR, name + ".json")


def _compose(p: Sequence[int], q: Sequence[int]) -> tuple:
    return tuple(q[k] for k in p)


def _group_from_generators(generators, number_spins: int):
    """All elements of the permutation group with their (real) characters."""
    identity = tuple(range(number_spins))
    elements = {identity: 1.0}
    gens = []
    for g in generators:
        perm = tuple(int(v) for v in g["permutation"])
        if sorted(perm) != list(identity):
            raise ValueError("not a permutation: {}".format(perm))
        sector = int(g.get("sector", 0))
        period, cur = 1, perm
        while cur != identity:
            cur = _compose(perm, cur)
            period += 1
        if sector == 0:
            chi = 1.0
        elif 2 * sector == period:
            chi = -1.0
        else:
            raise ValueError("expected all Hamiltonian matrix elements to be real: sector {} of a period-{} "
                             "symmetry has a complex character".format(sector, period))
        gens.append((perm, chi))
    frontier = [identity]
    while frontier:
        nxt = []
        for e in frontier:
            for perm, chi in gens:
                h = _compose(perm, e)
                if h not in elements:
                    elements[h] = elements[e] * chi
                    nxt.append(h)
        frontier = nxt
    return elements


class SpinBasis:
    def __init__(self, number_spins: int, hamming_weight: Optional[int] = None,
                 spin_inversion: Optional[int] = None, symmetries: Optional[List[dict]] = None):
        if not 1 <= int(number_spins) <= 64:
            raise ValueError("only works with up to 64 bits")  # common.py:86
        self.number_spins = int(number_spins)
        self.hamming_weight = None if hamming_weight is None else int(hamming_weight)
        self.spin_inversion = int(spin_inversion) if spin_inversion else 0
        self.symmetries = list(symmetries or [])
        self._group = _group_from_generators(self.symmetries, self.number_spins)
        self._states = None

    @staticmethod
    def load_from_yaml(cfg: dict) -> "SpinBasis":
        return SpinBasis(cfg["number_spins"], cfg.get("hamming_weight"), cfg.get("spin_inversion"),
                         cfg.get("symmetries") or [])

    @property
    def group_elements(self):
        """Non-identity permutations and their characters."""
        identity = tuple(range(self.number_spins))
        return [(p, c) for p, c in sorted(self._group.items()) if p != identity]

    @property
    def is_symmetrised(self) -> bool:
        return self.spin_inversion != 0 or len(self._group) > 1

    def build(self, representatives=None):
        if representatives is not None:
            self._states = np.ascontiguousarray(representatives, dtype=np.uint64)
            return self
        if self.is_symmetrised and len(self._group) > 1:
            raise NotImplementedError("enumerating a permutation-symmetrised sector: pass representatives")
        if self.number_spins > 30:
            raise ValueError("full-basis enumeration is limited to 30 spins; pass representatives")
        dev = require_cuda()
        x = torch.arange(1 << self.number_spins, dtype=torch.int64, device=dev)
        if self.hamming_weight is not None:
            pop = torch.zeros_like(x)
            for b in range(self.number_spins):
                pop += (x >> b) & 1
            x = x[pop == self.hamming_weight]
        if self.spin_inversion:
            x = x[((x >> (self.number_spins - 1)) & 1) == 0]  # representative = min(s, ~s)
        self._states = x.cpu().numpy().view(np.uint64)
        return self

    @property
    def states(self) -> np.ndarray:
        if self._states is None:
            self.build()
        return self._states

    @property
    def number_states(self) -> int:
        return int(self.states.shape[0])

    def states_device(self) -> torch.Tensor:
        """The sorted basis on the device (kept: batched_index is called once per cluster)."""
        if getattr(self, "_states_dev", None) is None or self._states_dev_of is not self._states:
            dev = require_cuda()
            self._states_dev = torch.from_numpy(self.states.view(np.int64)).to(dev)
            self._states_dev_of = self._states
        return self._states_dev

    def batched_index_device(self, spins: torch.Tensor) -> torch.Tensor:
        """Positions of ``spins`` (int64 bit patterns, CUDA) in the basis; raises when one is absent
        (lattice_symmetries raises too; the reference's callers rely on it: common.py:283, :817)."""
        from ._lib import check, ffi, lib, ptr, stream

        states = self.states_device()
        spins = spins.contiguous()
        out = torch.empty(spins.shape[0], dtype=torch.int64, device=states.device)
        missing = ffi.new("uint64_t *")
        check(lib().asp_batched_index(states.shape[0], ptr(states, "uint64_t *"), spins.shape[0], ptr(spins, "uint64_t *"),
                                      ptr(out, "int64_t *"), missing, stream()))
        if int(missing[0]) != 0:
            raise ValueError("%d state(s) not in the basis" % int(missing[0]))
        return out

    def batched_index(self, spins) -> np.ndarray:
        spins = np.ascontiguousarray(np.asarray(spins, dtype=np.uint64).reshape(-1))
        dev = require_cuda()
        return self.batched_index_device(torch.from_numpy(spins.view(np.int64)).to(dev)).cpu().numpy()

    def index(self, spin: int) -> int:
        return int(self.batched_index(np.array([spin], dtype=np.uint64))[0])


class Operator:
    def __init__(self, basis: SpinBasis, terms: List[dict]):
        self.basis = basis
        self.terms = []
        for t in terms:
            m = np.asarray(t["matrix"])
            if np.iscomplexobj(m):
                if not np.allclose(m.imag, 0, atol=1e-6):
                    raise ValueError("expected all Hamiltonian matrix elements to be real")  # common.py:97-98
                m = m.real
            m = np.ascontiguousarray(m, dtype=np.float64)
            if m.shape != (4, 4):
                raise ValueError("only two-site (4x4) terms are supported")
            sites = np.ascontiguousarray(t["sites"], dtype=np.uint32).reshape(-1, 2)
            self.terms.append((m, sites))
        self._handle = None

    @staticmethod
    def load_from_yaml(cfg: dict, basis: SpinBasis) -> "Operator":
        return Operator(basis, cfg["terms"])

    @staticmethod
    def load(path: str) -> "Operator":
        cfg = load_config(path)
        return Operator.load_from_yaml(cfg["hamiltonian"], SpinBasis.load_from_yaml(cfg["basis"]))

    # -- C-ABI handle -----------------------------------------------------------------
    @property
    def handle(self):
        if self._handle is None:  # host-side compilation needs no device; kernels do
            mats = np.ascontiguousarray(np.stack([m for m, _ in self.terms]).reshape(-1), dtype=np.float64) \
                if self.terms else np.zeros(0, dtype=np.float64)
            offsets = np.zeros(len(self.terms) + 1, dtype=np.uint32)
            offsets[1:] = np.cumsum([s.shape[0] for _, s in self.terms])
            sites = np.ascontiguousarray(np.concatenate([s for _, s in self.terms]).reshape(-1), dtype=np.uint32) \
                if self.terms else np.zeros(0, dtype=np.uint32)
            group = self.basis.group_elements
            perms = np.ascontiguousarray([p for p, _ in group], dtype=np.uint32).reshape(-1)
            chars = np.ascontiguousarray([c for _, c in group], dtype=np.float64)
            out = ffi.new("asp_operator **")
            hw = -1 if self.basis.hamming_weight is None else self.basis.hamming_weight
            check(lib().asp_operator_create(
                out, self.basis.number_spins, hw, self.basis.spin_inversion, len(self.terms),
                ffi.cast("double *", mats.ctypes.data), ffi.cast("uint32_t *", offsets.ctypes.data),
                ffi.cast("uint32_t *", sites.ctypes.data), len(group),
                ffi.cast("uint32_t *", perms.ctypes.data) if len(group) else ffi.NULL,
                ffi.cast("double *", chars.ctypes.data) if len(group) else ffi.NULL))
            self._handle = ffi.gc(out[0], lib().asp_operator_destroy)
        return self._handle

    @property
    def max_candidates(self) -> int:
        return int(lib().asp_operator_max_candidates(self.handle))

    @property
    def is_sorted_emitter(self) -> bool:
        return bool(lib().asp_operator_is_sorted_emitter(self.handle))

    # -- the reference-facing calls -----------------------------------------------------
    def batched_apply_device(self, spins: torch.Tensor):
        """spins: int64/uint64 CUDA tensor [m] -> (other_spins[T] int64, coeffs[T] f64, counts[m] int64)."""
        dev = require_cuda()
        spins = spins.to(dev).contiguous()
        if spins.dtype == torch.uint64:
            spins = spins.view(torch.int64)
        m = spins.shape[0]
        cap = m * self.max_candidates
        other = torch.empty(cap, dtype=torch.int64, device=dev)
        coeffs = torch.empty(cap, dtype=torch.float64, device=dev)
        counts = torch.empty(m, dtype=torch.int64, device=dev)
        total = ffi.new("uint64_t *")
        check(lib().asp_operator_apply_dev(self.handle, m, ptr(spins, "uint64_t *"), ptr(other, "uint64_t *"),
                                           ptr(coeffs, "double *"), ptr(counts, "int64_t *"), cap, total, stream()))
        t = int(total[0])
        return other[:t], coeffs[:t], counts

    def batched_apply(self, x):
        """``ls.Operator.batched_apply`` contract (common.py:96): x[m,8] (or [m]) uint64 ->
        (spins[T,8] uint64, coeffs[T] complex128, counts[m])."""
        x = np.asarray(x, dtype=np.uint64)
        if x.ndim == 2:
            if x.shape[1] != 8:
                raise ValueError("'spins' has wrong shape: {}; expected (?, 8)".format(x.shape))
            x = np.ascontiguousarray(x[:, 0])
        elif x.ndim != 1:
            raise ValueError("'spins' has wrong shape: {}; expected a 2D array".format(x.shape))
        dev = require_cuda()
        other, coeffs, counts = self.batched_apply_device(torch.from_numpy(x.view(np.int64)).to(dev))
        out = np.zeros((other.shape[0], 8), dtype=np.uint64)
        out[:, 0] = other.cpu().numpy().view(np.uint64)
        return out, coeffs.cpu().numpy().astype(np.complex128), counts.cpu().numpy()

    def apply(self, spin: int):
        s, c, _ = self.batched_apply(np.array([spin], dtype=np.uint64))
        return s, c


def batched_index(basis: SpinBasis, spins) -> np.ndarray:
    """``ls.batched_index(basis, spins)`` as the reference calls it (common.py:817)."""
    return basis.batched_index(spins)
