"""Host-side mirror of the ``ising_glass_annealer`` API the reference uses (``import
ising_glass_annealer as sa``): ``sa.Hamiltonian(exchange, field)`` (common.py:204),
``.energy(bits)`` (experiments/full_hilbert_space.py:144), ``sa.anneal(...)``
(common.py:242-248, full_hilbert_space.py:212-218), ``sa.signs_to_bits`` /
``sa.bits_to_signs`` (common.py:205, 224-225).  All compute runs in libasp_b200.so.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import scipy.sparse
import torch

from ._lib import check, ffi, lib, ptr, require_cuda, stream


def signs_to_bits(signs) -> np.ndarray:
    """Packed LSB-first bits, 1 <=> sign > 0 (cbits/build_matrix.c:67-76 semantics; 0 -> bit 0)."""
    dev = require_cuda()
    s = torch.as_tensor(np.ascontiguousarray(signs, dtype=np.float64)).to(dev)
    return signs_to_bits_device(s).cpu().numpy().view(np.uint64)


def signs_to_bits_device(psi: torch.Tensor) -> torch.Tensor:
    n = psi.shape[0]
    out = torch.empty((n + 63) // 64, dtype=torch.int64, device=psi.device)
    check(lib().asp_extract_signs_dev(n, ptr(psi.contiguous(), "double *"), ptr(out, "uint64_t *"), stream()))
    return out


def bits_to_signs(bits, count: int) -> np.ndarray:
    """``(bits[i//64] >> (i%64)) & 1`` -> +1/-1 (annealing_sign_problem/train.py:247-252)."""
    bits = np.ascontiguousarray(bits, dtype=np.uint64)
    i = np.arange(count, dtype=np.uint64)
    b = (bits[(i >> np.uint64(6)).astype(np.int64)] >> (i & np.uint64(63))) & np.uint64(1)
    return 2.0 * b.astype(np.float64) - 1.0


class Hamiltonian:
    """Container with the reference's attribute surface: ``.exchange`` (a real scipy sparse
    matrix), ``.field``, ``.shape``, ``.energy(bits)``.  Keeps a device-resident CSR copy."""

    def __init__(self, exchange, field, _device_csr=None):
        self.exchange = exchange
        self.field = np.ascontiguousarray(field, dtype=np.float64)
        self.shape = exchange.shape
        self._dev = _device_csr  # (indptr i64, indices i32, data f64, field f64 | None)

    @property
    def size(self) -> int:
        return int(self.shape[0])

    def device_csr(self):
        if self._dev is None:
            dev = require_cuda()
            csr = scipy.sparse.csr_matrix(self.exchange)
            csr.sum_duplicates()
            csr.sort_indices()
            # the annealer and the greedy solver take dE from row p alone and colour the graph by row adjacency: both
            # are only right for a symmetric exchange matrix (E = s^T J s with J = J^T, as make_ising_model builds it)
            if csr.shape[0] != csr.shape[1] or (csr != csr.T).nnz != 0:
                raise ValueError("'exchange' must be a symmetric matrix")
            fld = torch.from_numpy(self.field).to(dev) if self.field.any() else None
            self._dev = (
                torch.from_numpy(csr.indptr.astype(np.int64)).to(dev),
                torch.from_numpy(csr.indices.astype(np.int32)).to(dev),
                torch.from_numpy(np.ascontiguousarray(csr.data, dtype=np.float64)).to(dev),
                fld,
            )
        return self._dev

    def energy(self, bits) -> float:
        dev = require_cuda()
        b = torch.from_numpy(np.ascontiguousarray(bits, dtype=np.uint64).view(np.int64)).to(dev)
        return float(self.energies_device(b.reshape(1, -1))[0])

    def energies_device(self, bits: torch.Tensor) -> torch.Tensor:
        """bits: [R, ceil(n/64)] int64 CUDA tensor -> energies [R] f64."""
        indptr, indices, data, fld = self.device_csr()
        bits = bits.contiguous()
        out = torch.empty(bits.shape[0], dtype=torch.float64, device=bits.device)
        check(lib().asp_energy(self.size, ptr(indptr, "int64_t *"), ptr(indices, "int32_t *"), ptr(data, "double *"),
                               ptr(fld, "double *"), bits.shape[0], ptr(bits, "uint64_t *"), ptr(out, "double *"), stream()))
        return out


def default_betas(hamiltonian: Hamiltonian, number_sweeps: int, beta0: Optional[float] = None,
                  beta1: Optional[float] = None) -> np.ndarray:
    """Geometric ladder beta0 -> beta1 (the reference leaves both to the annealer,
    common.py:242-248).  Auto ends: hot = largest single-flip barrier accepted with
    probability 1/2, cold = smallest non-zero barrier accepted with probability 1/100 (capped at
    1e4 x hot), then max(8, number_sweeps//50) zero-temperature sweeps (beta = inf)."""
    indptr, indices, data, fld = hamiltonian.device_csr()
    n = hamiltonian.size
    rows = torch.repeat_interleave(torch.arange(n, device=data.device), indptr[1:] - indptr[:-1])
    off = rows != indices.to(torch.int64)
    a = data.abs()[off]
    row_sum = torch.zeros(n, dtype=torch.float64, device=data.device).index_add_(0, rows[off], a)
    f = fld.abs() if fld is not None else torch.zeros(n, dtype=torch.float64, device=data.device)
    max_de = float((4.0 * row_sum + 2.0 * f).max()) if n else 1.0
    cands = []
    nz = a[a > 0]
    if nz.numel():
        cands.append(4.0 * float(nz.min()))
    fz = f[f > 0]
    if fz.numel():
        cands.append(2.0 * float(fz.min()))
    min_de = min(cands) if cands else 1.0
    if max_de <= 0:
        max_de = 1.0
    b0 = np.log(2.0) / max_de if beta0 is None else float(beta0)
    # cold end: the smallest barrier, but at most 4 decades above the hot end -- amplitudes
    # span many decades and a ladder reaching 1/min|J| would spend its sweeps frozen
    b1 = min(np.log(100.0) / min_de, b0 * 1e4) if beta1 is None else float(beta1)
    # final zero-temperature sweeps: a fiftieth of the run, but at least 8 (the spins of tiny amplitude only settle
    # there; with 2 such sweeps a 100-sweep run of kagome_16 never reached accuracy > 0.995) and at most a quarter
    quench = min(max(number_sweeps // 50, 8), number_sweeps // 4) if beta1 is None else 0
    ladder = number_sweeps - quench
    if ladder <= 1:
        betas = np.full(number_sweeps, b1, dtype=np.float64)
    else:
        t = np.arange(ladder, dtype=np.float64) / (ladder - 1)
        betas = np.concatenate([b0 * (b1 / b0) ** t, np.full(quench, np.inf)])
    return np.ascontiguousarray(betas, dtype=np.float64)


def energy_scale(hamiltonian: Hamiltonian) -> float:
    """Power of two that turns energy differences into integers: every single-flip change stays
    below 2^50 (the kernel rounds with the 1.5 * 2^52 trick, exact below 2^51) and every running sum
    well inside int64."""
    _, _, data, fld = hamiltonian.device_csr()
    w = float(data.abs().sum()) + (float(fld.abs().sum()) if fld is not None else 0.0)
    return float(2.0 ** np.floor(np.log2(2.0 ** 50 / (2.0 * w + 1.0))))


class AnnealPlan:
    """Coloured + relabelled model on the device (asp_sa_plan)."""

    def __init__(self, hamiltonian: Hamiltonian):
        require_cuda()
        self.hamiltonian = hamiltonian
        indptr, indices, data, fld = hamiltonian.device_csr()
        out = ffi.new("asp_sa_plan **")
        check(lib().asp_sa_plan_create(out, hamiltonian.size, ptr(indptr, "int64_t *"), ptr(indices, "int32_t *"),
                                       ptr(data, "double *"), ptr(fld, "double *"), stream()))
        self.handle = ffi.gc(out[0], lib().asp_sa_plan_destroy)
        np_, nc, nnz = ffi.new("uint64_t *"), ffi.new("uint32_t *"), ffi.new("uint64_t *")
        check(lib().asp_sa_plan_info(self.handle, np_, nc, nnz))
        self.n_padded, self.num_classes, self.nnz = int(np_[0]), int(nc[0]), int(nnz[0])

    def export(self):
        """Relabelled model on the host (what the CPU oracle is run on in the tests)."""
        order = np.empty(self.n_padded, dtype=np.int32)
        class_ptr = np.empty(self.num_classes + 1, dtype=np.int64)
        indptr = np.empty(self.n_padded + 1, dtype=np.int64)
        indices = np.empty(max(self.nnz, 1), dtype=np.int32)
        data = np.empty(max(self.nnz, 1), dtype=np.float64)
        field = np.empty(self.n_padded, dtype=np.float64)
        check(lib().asp_sa_plan_export(
            self.handle, ffi.cast("int32_t *", order.ctypes.data), ffi.cast("int64_t *", class_ptr.ctypes.data),
            ffi.cast("int64_t *", indptr.ctypes.data), ffi.cast("int32_t *", indices.ctypes.data),
            ffi.cast("double *", data.ctypes.data), ffi.cast("double *", field.ctypes.data)))
        return dict(order=order, class_ptr=class_ptr, indptr=indptr, indices=indices[:self.nnz],
                    data=data[:self.nnz], field=field)

    def greedy_device(self):
        """-> (bits [ceil(n/64)] int64 CUDA, energy f64 0-d CUDA, merge rounds, descent sweeps)."""
        dev = require_cuda()
        n = self.hamiltonian.size
        bits = torch.zeros((n + 63) // 64, dtype=torch.int64, device=dev)
        energy = torch.zeros(1, dtype=torch.float64, device=dev)
        rounds, sweeps = ffi.new("uint32_t *"), ffi.new("uint32_t *")
        check(lib().asp_greedy_solve(self.handle, ptr(bits, "uint64_t *"), ptr(energy, "double *"), rounds, sweeps, stream()))
        return bits, energy[0], int(rounds[0]), int(sweeps[0])

    def anneal_device(self, repetitions: int, betas: np.ndarray, seed: int, x0: Optional[torch.Tensor] = None,
                      escale: Optional[float] = None, replica_offset: int = 0, want_energies: bool = True):
        """-> (best_bits [R, ceil(n/64)] int64 CUDA, energies [R] f64 CUDA)."""
        dev = require_cuda()
        n = self.hamiltonian.size
        betas = np.ascontiguousarray(betas, dtype=np.float64)
        if escale is None:
            escale = energy_scale(self.hamiltonian)
        bits = torch.empty((repetitions, (n + 63) // 64), dtype=torch.int64, device=dev)
        energies = torch.empty(repetitions, dtype=torch.float64, device=dev)
        check(lib().asp_sa_anneal(self.handle, repetitions, replica_offset, betas.shape[0],
                                  ffi.cast("double *", betas.ctypes.data), int(seed) & 0xFFFFFFFFFFFFFFFF,
                                  ptr(x0, "uint64_t *"), float(escale), ptr(bits, "uint64_t *"),
                                  ptr(energies, "double *") if want_energies else ffi.NULL, stream()))
        return bits, energies


def anneal(hamiltonian: Hamiltonian, x0=None, seed: Optional[int] = None, number_sweeps: int = 5120,
           beta0: Optional[float] = None, beta1: Optional[float] = None, repetitions: int = 1,
           only_best: bool = True):
    """``sa.anneal`` (common.py:242-248).  ``only_best=True`` -> (bits, energy) of the best
    replica (ties: lowest replica index); ``False`` -> (bits[R, words], energies[R])."""
    dev = require_cuda()
    if seed is None:
        seed = int(np.random.SeedSequence().entropy & 0xFFFFFFFFFFFFFFFF)
    plan = getattr(hamiltonian, "_plan", None)
    if plan is None:
        plan = AnnealPlan(hamiltonian)
        hamiltonian._plan = plan
    betas = default_betas(hamiltonian, number_sweeps, beta0, beta1)
    x0_dev = None
    if x0 is not None:
        x0_dev = torch.from_numpy(np.ascontiguousarray(x0, dtype=np.uint64).view(np.int64)).to(dev)
    bits, energies = plan.anneal_device(repetitions, betas, seed, x0_dev)
    if only_best:
        best = int(torch.nonzero(energies == energies.min())[0])  # first minimum
        return bits[best].cpu().numpy().view(np.uint64), float(energies[best])
    return bits.cpu().numpy().view(np.uint64), energies.cpu().numpy()


def greedy_solve(hamiltonian: Hamiltonian):
    """``sa.greedy_solve`` (common.py:249-250) -> (bits, energy): strongest couplings first (clusters
    merge with the joining edge satisfied), then local descent until no flip lowers the energy
    (asp_greedy_solve; algorithm and its two deviations from common.py:298-438 in csrc/greedy.cu)."""
    dev = require_cuda()
    plan = getattr(hamiltonian, "_plan", None)
    if plan is None:
        plan = AnnealPlan(hamiltonian)
        hamiltonian._plan = plan
    bits, energy, _, _ = plan.greedy_device()
    return bits.cpu().numpy().view(np.uint64), float(energy)
