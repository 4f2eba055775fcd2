"""On-disk formats either side of the path.

* ground states: the reference reads ``/hamiltonian/eigenvectors``, ``/hamiltonian/eigenvalues``,
  ``/basis/representatives`` from HDF5 (annealing_sign_problem/common.py:771-780);
* extracted models: ``dump_ising_model_to_hdf5`` writes ``elements f64, indices int32, indptr int32,
  field f64, energy, signs u64`` (common.py:750-768) -- int32 index arrays, i.e. nnz < 2^31, the
  same limit scipy applies to the matrix ``make_ising_model`` returns;
* results: one CSV line per cluster (experiments/sampled_connected_components.py:672-693, 804-831).

HDF5 needs ``h5py`` (absent from this image: the functions then raise ImportError); the same
datasets under the same names are also read from / written to ``.npz`` so the pipeline runs offline.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

from . import annealer as sa

_KEYS = ("/hamiltonian/eigenvectors", "/hamiltonian/eigenvalues", "/basis/representatives")


def _h5py():
    try:
        import h5py  # noqa: F401

        return h5py
    except ImportError as exc:
        raise ImportError("reading/writing .h5 files needs h5py, which this environment lacks; "
                          "use the .npz form (same dataset names)") from exc


def load_ground_state(filename: str) -> Tuple[np.ndarray, float, np.ndarray]:  # common.py:771-780
    if filename.endswith(".npz"):
        f = np.load(filename)
        vectors, values, reps = (f[k] for k in _KEYS)
    else:
        with _h5py().File(filename, "r") as f:
            vectors, values, reps = (np.asarray(f[k]) for k in _KEYS)
    ground_state = np.asarray(vectors, dtype=np.float64).squeeze()
    if ground_state.ndim > 1:
        ground_state = ground_state[0, :]
    energy = float(np.asarray(values).reshape(-1)[0])
    return ground_state, energy, np.asarray(reps, dtype=np.uint64)


def save_ground_state(filename: str, ground_state, energy: float, representatives) -> None:
    """Writes what load_ground_state reads (the reference gets these files from an ED code)."""
    data = {_KEYS[0]: np.asarray(ground_state, dtype=np.float64).reshape(1, -1),
            _KEYS[1]: np.asarray([energy], dtype=np.float64),
            _KEYS[2]: np.asarray(representatives, dtype=np.uint64)}
    if filename.endswith(".npz"):
        np.savez(filename, **data)
    else:
        with _h5py().File(filename, "w") as out:
            for k, v in data.items():
                out[k] = v


def ising_model_datasets(model, ground_state) -> dict:
    """The datasets of dump_ising_model_to_hdf5 (common.py:750-768).  ``energy`` is s^T J s + h.s of
    the exact signs -- equal to <psi|H|psi> when psi is an eigenvector on the full basis
    (full_hilbert_space.py:143-145), which is what the reference stores via ``expectation``."""
    matrix = model.ising_hamiltonian.exchange.tocsr()
    if matrix.nnz >= 2 ** 31 or matrix.shape[0] >= 2 ** 31:
        raise ValueError("the model dump stores int32 indices/indptr: needs nnz < 2^31")
    signs = sa.signs_to_bits(np.sign(ground_state))
    return {"elements": np.asarray(matrix.data, dtype=np.float64), "indices": np.asarray(matrix.indices, dtype=np.int32),
            "indptr": np.asarray(matrix.indptr, dtype=np.int32), "field": np.asarray(model.ising_hamiltonian.field, dtype=np.float64),
            "energy": np.float64(model.ising_hamiltonian.energy(signs)), "signs": signs}


def dump_ising_model_to_hdf5(model, ground_state, filename: str) -> None:
    data = ising_model_datasets(model, ground_state)
    if filename.endswith(".npz"):
        np.savez(filename, **data)
        return
    with _h5py().File(filename, "w") as out:
        for k, v in data.items():
            out[k] = v
