"""CPU (-m "not gpu") tests of the drop-in boundary: the C-ABI library builds, loads and
exports every symbol include/asp_b200.h declares; the host-side mirror validates inputs the
way the reference does; the product fails loudly without a GPU and never touches oracle/."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.fixture(scope="module")
def asp():
    import annealing_sign_problem_b200 as mod
    from annealing_sign_problem_b200.build_extension import build

    build()
    return mod


def test_library_exports_every_declared_symbol(asp):
    from annealing_sign_problem_b200._lib import lib

    header = open(os.path.join(ROOT, "include", "asp_b200.h")).read()
    names = sorted(set(re.findall(r"\b(asp_[a-z0-9_]+)\s*\(", header)))
    assert len(names) >= 25
    handle = lib()
    for name in names:
        assert hasattr(handle, name), name
    out = subprocess.check_output(["nm", "-D", "--defined-only", os.path.join(ROOT, "annealing-sign-problem_b200", "libasp_b200.so")]).decode()
    exported = set(re.findall(r" T (asp_[a-z0-9_]+)", out))
    assert set(names) <= exported
    assert handle.asp_version() >= 100


REFERENCE_CDEF = """
    typedef struct ls_bits512 { uint64_t words[8]; } ls_bits512;
    uint64_t build_matrix(uint64_t num_spins, ls_bits512 const spins[], int64_t const *counts, double const *psi,
                          ls_bits512 const *other_spins, double const *other_coeffs, int64_t const *other_counts,
                          double const *other_psi, uint32_t *row_indices, uint32_t *col_indices, double *elements,
                          double *field);
    void extract_signs(uint64_t num_spins, double const *psi, uint64_t *signs);
"""


def reference_cdef():
    """The declarations the reference binds (annealing_sign_problem/build_extension.py:5-21): read from the reference
    itself where it is present (this container), else the equivalent text above (the GPU box has no /root/reference)."""
    path = "/root/reference/annealing_sign_problem/build_extension.py"
    if os.path.exists(path):
        m = re.search(r'ffibuilder\.cdef\(\s*"""(.*?)"""', open(path).read(), flags=re.S)
        if m:
            return m.group(1)
    return REFERENCE_CDEF


def test_reference_cdef_binds_the_library_unchanged(asp):
    """Drop-in at the reference's own FFI seam: its cdef, fed to ffi.dlopen of OUR library, finds both symbols."""
    from cffi import FFI

    ffi = FFI()
    ffi.cdef(reference_cdef())
    handle = ffi.dlopen(os.path.join(ROOT, "annealing-sign-problem_b200", "libasp_b200.so"))
    assert handle.build_matrix is not None and handle.extract_signs is not None
    assert ffi.sizeof("ls_bits512") == 64


def test_library_is_built_for_sm_100a():
    lib_path = os.path.join(ROOT, "annealing-sign-problem_b200", "libasp_b200.so")
    out = subprocess.run(["cuobjdump", "--list-elf", lib_path], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def test_no_cpu_fallback(asp):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from annealing_sign_problem_b200._lib import AspError

    op = asp.load_hamiltonian(asp.ls.system_path("j1j2_square_4x4"))
    with pytest.raises(AspError):
        asp.make_ising_model(np.array([255, 510], dtype=np.uint64), op, log_psi=np.zeros(2))
    with pytest.raises(AspError):
        asp.sa.signs_to_bits(np.ones(4))
    with pytest.raises(AspError):
        asp.compute_accuracy_and_overlap(np.zeros(1, dtype=np.uint64), np.zeros(1, dtype=np.uint64), number_spins=3)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "annealing-sign-problem_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "oracle/" not in text or f.endswith((".cu", ".cuh")), f


def test_argument_validation_matches_the_reference(asp):
    op = asp.load_hamiltonian(asp.ls.system_path("heisenberg_kagome_16"))
    with pytest.raises(ValueError):  # common.py:140-141
        asp.make_ising_model(np.zeros(3, dtype=np.uint64), op)
    with pytest.raises(ValueError):  # common.py:142-143
        asp.make_ising_model(np.zeros(3, dtype=np.uint64), op, log_psi=np.zeros(3), external_field=True)
    with pytest.raises(ValueError):  # common.py:217-218
        asp.compute_accuracy_and_overlap(np.zeros(1, dtype=np.uint64), np.zeros(1, dtype=np.uint64))
    assert np.array_equal(asp.binary_search(np.array([1, 5, 9]), np.array([9, 1])), [2, 0])
    with pytest.raises(AssertionError):  # common.py:544-548
        asp.binary_search(np.array([1, 5, 9]), np.array([4]))
    assert np.array_equal(asp.sa.bits_to_signs(np.array([0b101], dtype=np.uint64), 4), [1, -1, 1, -1])


def test_symmetry_groups_and_system_files(asp):
    for name, order in [("heisenberg_kagome_36", 144), ("heisenberg_pyrochlore_2x2x2", 384), ("j1j2_square_4x4", 1)]:
        op = asp.load_hamiltonian(asp.ls.system_path(name))
        assert len(op.basis.group_elements) + 1 == order
    bad = {"number_spins": 4, "hamming_weight": 2, "symmetries": [{"permutation": [1, 2, 3, 0], "sector": 1}]}
    with pytest.raises(ValueError):  # complex character: the reference rejects non-real coefficients
        asp.ls.SpinBasis.load_from_yaml(bad)
    with pytest.raises(ValueError):
        asp.ls.Operator(asp.ls.SpinBasis(4), [{"matrix": np.eye(3), "sites": [[0, 1]]}])


def test_operator_compiles_to_sorted_moves_on_the_host(asp):
    """Bond list -> delta-sorted XOR moves: exchange terms on distinct pairs give a sorted,
    duplicate-free emitter (2 moves per bond + the diagonal); symmetrised bases do not."""
    for name, bonds in [("heisenberg_kagome_36", 72), ("j1j2_square_4x4", 64), ("sk_32_1", 496)]:
        cfg = asp.ls.load_config(asp.ls.system_path(name))
        cfg["basis"]["symmetries"], cfg["basis"]["spin_inversion"] = [], None
        op = asp.ls.Operator.load_from_yaml(cfg["hamiltonian"], asp.ls.SpinBasis.load_from_yaml(cfg["basis"]))
        assert op.max_candidates == 2 * bonds + 1
        assert op.is_sorted_emitter
    assert not asp.load_hamiltonian(asp.ls.system_path("heisenberg_kagome_18")).is_sorted_emitter
    assert not asp.load_hamiltonian(asp.ls.system_path("heisenberg_kagome_36")).is_sorted_emitter  # Benes networks self-check
    # the same bond twice merges into one move; a single-site-like flip from two bonds is not sorted
    basis = asp.ls.SpinBasis(4)
    heis = [[1, 0, 0, 0], [0, -1, 2, 0], [0, 2, -1, 0], [0, 0, 0, 1]]
    twice = asp.ls.Operator(basis, [{"matrix": heis, "sites": [[0, 1], [0, 1]]}])
    assert twice.max_candidates == 3 and twice.is_sorted_emitter
    flip_j = [[0, 1, 0, 0], [1, 0, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]]  # sigma^x on the second site
    clash = asp.ls.Operator(basis, [{"matrix": flip_j, "sites": [[0, 2], [1, 2]]}])
    assert not clash.is_sorted_emitter


def test_two_rank_sharding_plan_over_gloo(tmp_path):
    """world_size-2 gloo run of the row-block / replica partition used by bench.py --gpus N."""
    script = os.path.join(ROOT, "tests", "_gloo_sharding.py")
    port = 29500 + os.getpid() % 2000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), script]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "SHARDING_OK" in out.stdout


def test_error_convention_of_the_exchange_and_sampling_entry_points(asp):
    """ADDITION to the reference (which has no error channel): every `int` entry point answers a bad argument with
    ASP_ERR_ARG and a message in asp_last_error() before it touches the device -- checked here without a GPU for the
    entry points of the exchange step X1 and of the sampling front-end."""
    from annealing_sign_problem_b200._lib import ffi, lib

    L = lib()
    NULL = ffi.NULL
    two = ffi.new("uint64_t *[]", [ffi.cast("uint64_t *", 256), ffi.cast("uint64_t *", 512)])
    keys = ffi.new("uint64_t const *[]", [ffi.cast("uint64_t const *", 256), ffi.cast("uint64_t const *", 512)])
    amps = ffi.new("double const *[]", [ffi.cast("double const *", 256), ffi.cast("double const *", 512)])
    begin = ffi.new("uint64_t[]", [0, 10, 20])

    def message():
        return ffi.string(L.asp_last_error()).decode()

    assert L.asp_peer_signal(0, two, 0, 1, NULL) == L.ASP_ERR_ARG and "1..16" in message()
    assert L.asp_peer_signal(17, two, 0, 1, NULL) == L.ASP_ERR_ARG
    assert L.asp_peer_wait(2, NULL, 1, NULL) == L.ASP_ERR_ARG
    assert L.asp_peer_alloc(0, ffi.new("void **"), ffi.new("unsigned char[64]")) == L.ASP_ERR_ARG
    assert L.asp_peer_open(NULL, ffi.new("void **")) == L.ASP_ERR_ARG
    assert L.asp_peer_close(NULL) == L.ASP_OK and L.asp_peer_free(NULL) == L.ASP_OK  # like free(NULL)
    assert L.asp_gather_blocks(0, 0, begin, keys, amps, NULL, 0, ffi.cast("uint64_t *", 256), ffi.cast("double *", 256), NULL) == L.ASP_ERR_ARG
    assert L.asp_gather_blocks(2, 0, begin, keys, amps, NULL, 0, NULL, NULL, NULL) == L.ASP_ERR_ARG and "NULL output" in message()
    assert L.asp_gather_blocks(2, 0, NULL, keys, amps, NULL, 0, ffi.cast("uint64_t *", 256), ffi.cast("double *", 256), NULL) == L.ASP_ERR_ARG
    assert L.asp_gather_index(NULL, 2, 0, begin, keys, amps, NULL, 0, ffi.cast("uint64_t *", 256), ffi.cast("double *", 256), 10, NULL, 0,
                              NULL) == L.ASP_ERR_ARG and "operator is NULL" in message()
    assert L.asp_extract_csr_indexed(NULL, 20, NULL, NULL, 0, 10, NULL, 0, 0, NULL, NULL, NULL, NULL, NULL) == L.ASP_ERR_ARG
    assert L.asp_extract_host_i32(NULL, 20, NULL, NULL, 0, 10, 0, NULL, NULL, NULL, NULL) == L.ASP_ERR_ARG
    assert L.asp_extract_indexed_to_host_i32(NULL, 20, NULL, NULL, 0, 10, NULL, 0, 0, NULL, NULL, NULL, NULL, NULL) == L.ASP_ERR_ARG
    assert L.asp_sample_indices(0, NULL, 2.0, 0, NULL, NULL, NULL, NULL) == L.ASP_ERR_ARG and "empty distribution" in message()
    assert L.asp_sample_indices(5, ffi.cast("double *", 256), 2.0, 3, NULL, NULL, NULL, NULL) == L.ASP_ERR_ARG
    assert L.asp_batched_index(5, NULL, 0, NULL, NULL, NULL, NULL) == L.ASP_ERR_ARG
    assert L.asp_batched_index(5, ffi.cast("uint64_t *", 256), 3, NULL, NULL, NULL, NULL) == L.ASP_ERR_ARG
    L.asp_set_gather_mode(7)  # unknown modes select the default, never an undefined path
    L.asp_set_gather_mode(2)
