"""CPU oracle for the Ising-extraction + annealing hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or as the
timed CPU baseline.  The product package (``annealing-sign-problem_b200/``) never
imports this package and fails loudly when its CUDA library is missing.

Contents (each module cites the reference file:line it restates):

* ``operator_np``   numpy restatement of ``lattice_symmetries.Operator.batched_apply``
                    (third-party, un-vendored, pinned =0.8.3; PARITY UNPINNED at this
                    boundary, anchored on the reference call sites common.py:85-106).
* ``extract_port.c``  C restatement of ``cbits/build_matrix.c`` (+ canonical CSR).
* ``live_path``     numpy/scipy restatement of ``common.py:make_ising_model``, of the reductions and of the
                    sampling front-end (``sample_indices``, ``batched_index``, ``log_coeff``); PINNED by golden
                    vectors the reference itself produced (``tests/golden/make_golden.py``).
* ``anneal_port.c`` C restatement of the replica Metropolis annealer
                    (``ising_glass_annealer.anneal``; third-party, un-vendored, pinned
                    =0.4.1.2; PARITY UNPINNED, anchored on outcomes: E0, bit layout).
* ``greedy_port.c`` Kruskal restatement of ``ising_glass_annealer.greedy_solve`` from the Python the
                    reference preserves at common.py:298-438 (PARITY UNPINNED).
* ``synthetic_np``  numpy generator of the CPU legs' inputs (the reference arm of bench.py uses nothing else).
* ``_ref/``         the reference's own ``cbits/build_matrix.c`` compiled where it lies
                    (git-ignored; built by ``oracle/Makefile``).
"""
