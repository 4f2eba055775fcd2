"""Synthetic inputs of the BASELINE shapes: the ED ``.h5`` ground states
of the reference (common.py:771-780) are not available offline, so benchmarks and tests
draw random sampled subsets and log-normal amplitudes.  Works on CPU or CUDA tensors.
"""
from __future__ import annotations

import numpy as np
import torch

_SIGN_BIT = -0x8000000000000000


def _sorted_unique_unsigned(x: torch.Tensor) -> torch.Tensor:
    return torch.unique(x ^ _SIGN_BIT, sorted=True) ^ _SIGN_BIT


def random_sector_states(number_spins: int, hamming_weight: int, n: int, seed: int, device="cpu") -> torch.Tensor:
    """Flavour (U): n distinct uniformly random words of the U(1) sector, ascending."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    out = torch.zeros(0, dtype=torch.int64, device=device)
    want = n
    while out.shape[0] < n:
        m = int((want - out.shape[0]) * 1.05) + 64
        words = torch.zeros(m, dtype=torch.int64, device=device)
        for start in range(0, m, 1 << 20):
            cnt = min(1 << 20, m - start)
            keys = torch.rand(cnt, number_spins, generator=gen, device=device)
            up = torch.argsort(keys, dim=1)[:, :hamming_weight]
            w = (torch.ones_like(up, dtype=torch.int64) << up).sum(dim=1)
            words[start:start + cnt] = w
        out = _sorted_unique_unsigned(torch.cat([out, words]))
    if out.shape[0] > n:
        keep = torch.randperm(out.shape[0], generator=gen, device=device)[:n]
        out = out[torch.sort(keep).values]
    return out.contiguous()


def _setdiff_sorted(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Elements of sorted-unique a (unsigned order) that are not in sorted-unique b."""
    if b.numel() == 0:
        return a
    ka, kb = a ^ _SIGN_BIT, b ^ _SIGN_BIT
    idx = torch.searchsorted(kb, ka).clamp_(max=kb.shape[0] - 1)
    return a[kb[idx] != ka]


def cluster_closed_states(operator, n: int, seed: int, device="cuda", interior_fraction: float = 1.0 / 3.0) -> torch.Tensor:
    """Flavour (C): random seeds, their full batched_apply shell (the "interior": every
    neighbour of a seed is present), then a random part of the second shell up to n states --
    the shape make_hamiltonian_extension (common.py:516-522) produces after ORDER=2
    extensions.  About `interior_fraction` of the rows keep all their neighbours, so roughly a
    third of all candidates are hits (the regime of SURVEY.md section 8d).  Needs CUDA."""
    basis = operator.basis
    # candidates per row: one of a bond's two exchange moves applies to an antiparallel pair,
    # about half of the bonds are antiparallel at half filling
    d = max(2.0, operator.max_candidates / 4.0)
    m = max(1, int(n * interior_fraction / (d + 1.0)))
    seeds = random_sector_states(basis.number_spins, basis.hamming_weight, m, seed, device)
    shell1, _, _ = operator.batched_apply_device(seeds)
    interior = _sorted_unique_unsigned(torch.cat([seeds, shell1]))
    gen = torch.Generator(device=device)
    gen.manual_seed(seed + 1)
    if interior.shape[0] >= n:
        keep = torch.randperm(interior.shape[0], generator=gen, device=device)[:n]
        return interior[torch.sort(keep).values].contiguous()
    boundary = torch.zeros(0, dtype=torch.int64, device=device)
    for start in range(0, interior.shape[0], 1 << 21):  # bounded candidate buffers
        shell2, _, _ = operator.batched_apply_device(interior[start:start + (1 << 21)])
        boundary = _sorted_unique_unsigned(torch.cat([boundary, _setdiff_sorted(_sorted_unique_unsigned(shell2), interior)]))
        del shell2
    need = n - interior.shape[0]
    if boundary.shape[0] > need:
        keep = torch.randperm(boundary.shape[0], generator=gen, device=device)[:need]
        boundary = boundary[torch.sort(keep).values]
    pool = _sorted_unique_unsigned(torch.cat([interior, boundary]))
    if pool.shape[0] < n:  # tiny sectors: top up with uniform states
        extra = _setdiff_sorted(random_sector_states(basis.number_spins, basis.hamming_weight, n, seed + 2, device), pool)
        pool = _sorted_unique_unsigned(torch.cat([pool, extra[: n - pool.shape[0]]]))
    return pool.contiguous()


def representative_cluster_states(operator, n: int, seed: int, device="cuda") -> torch.Tensor:
    """Cluster-closed subset of a SYMMETRISED basis: only images of batched_apply are kept (they are orbit
    representatives by construction, random words of the sector are not): first shell of random seeds, then a
    random part of the second shell up to n states.  Needs CUDA."""
    basis = operator.basis
    d = max(2.0, operator.max_candidates / 4.0)
    m = max(8, int(n / (d * d)) + 1)
    seeds = random_sector_states(basis.number_spins, basis.hamming_weight, m, seed, device)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed + 1)
    pool = _sorted_unique_unsigned(operator.batched_apply_device(seeds)[0])
    while pool.shape[0] < n:
        grown = torch.zeros(0, dtype=torch.int64, device=device)
        for start in range(0, pool.shape[0], 1 << 18):  # bounded candidate buffers (a group orbit per candidate)
            shell, _, _ = operator.batched_apply_device(pool[start:start + (1 << 18)])
            grown = _sorted_unique_unsigned(torch.cat([grown, shell]))
            if pool.shape[0] + grown.shape[0] > 4 * n:
                break
        new = _setdiff_sorted(grown, pool)
        if new.shape[0] == 0:
            break
        need = n - pool.shape[0]
        if new.shape[0] > need:
            new = new[torch.sort(torch.randperm(new.shape[0], generator=gen, device=device)[:need]).values]
        pool = _sorted_unique_unsigned(torch.cat([pool, new]))
    return pool.contiguous()


def synthetic_amplitudes(n: int, seed: int, sigma: float = 2.0, device="cpu") -> torch.Tensor:
    """psi_i = +-exp(sigma z_i), z ~ N(0,1), uniform sign, L2-normalised (common.py:181)."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    z = torch.randn(n, generator=gen, device=device, dtype=torch.float64)
    sign = torch.where(torch.rand(n, generator=gen, device=device) < 0.5, -1.0, 1.0).to(torch.float64)
    psi = sign * torch.exp(sigma * z)
    return (psi / torch.linalg.norm(psi)).contiguous()


def sk_config(number_spins: int, seed: int) -> dict:
    """SK-shaped operator: N(0,1) coupling x Heisenberg matrix on every pair i<j, exactly the
    construction of physical_systems/generate_sk.py:24-29 (different random stream)."""
    rng = np.random.default_rng(seed)
    base = np.array([[1, 0, 0, 0], [0, -1, 2, 0], [0, 2, -1, 0], [0, 0, 0, 1]], dtype=np.float64)
    terms = []
    for i in range(number_spins - 1):
        for j in range(i + 1, number_spins):
            terms.append({"matrix": (rng.normal() * base).tolist(), "sites": [[i, j]]})
    return {"basis": {"number_spins": number_spins, "hamming_weight": number_spins // 2, "symmetries": []},
            "hamiltonian": {"name": "Sherrington-Kirkpatrick", "terms": terms}}
