#!/usr/bin/env python
"""The reference's Figure-2 experiment (experiments/full_hilbert_space.py:205-246) on the GPU: for a full-basis model
built from the exact ground state, 1024 repetitions per number of sweeps, probabilities of accuracy > 0.995,
overlap > 0.995 and relative energy error <= 1e-12 (full_hilbert_space.py:168-185) -- beside the published columns
(tests/golden/published_sa_statistics.json, from the reference's experiments/*.csv).

    python tools/figure2_ladder.py [--systems a,b] [--sweeps 100,200,...] [--repetitions 1024] [--out file.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import annealing_sign_problem_b200 as asp  # noqa: E402
from annealing_sign_problem_b200 import common  # noqa: E402


def exact_ground_state(operator):
    import scipy.sparse
    import scipy.sparse.linalg

    dev = torch.device("cuda", 0)
    basis = operator.basis
    d_states = basis.states_device()
    n = int(d_states.shape[0])
    other, coeffs, counts = operator.batched_apply_device(d_states)
    cols = basis.batched_index_device(other)
    rows = torch.repeat_interleave(torch.arange(n, device=dev), counts)
    h = scipy.sparse.coo_matrix((coeffs.cpu().numpy(), (rows.cpu().numpy(), cols.cpu().numpy())), shape=(n, n)).tocsr()
    w, v = scipy.sparse.linalg.eigsh(h, k=2, which="SA", tol=1e-13, v0=np.random.default_rng(0).standard_normal(n))
    k = int(np.argmin(w))
    return float(w[k]), np.ascontiguousarray(v[:, k])


def ladder(system, sweeps_list, repetitions, seed=0):
    operator = asp.load_hamiltonian(asp.ls.system_path(system))
    e0, psi = exact_ground_state(operator)
    with np.errstate(divide="ignore"):
        model = asp.make_ising_model(operator.basis.states, operator, log_psi=np.log(psi.astype(np.complex128)))
    ham = model.ising_hamiltonian
    weights = psi ** 2 / np.sum(psi ** 2)
    rows = {}
    for sweeps in sweeps_list:
        t0 = time.perf_counter()
        xs, es = asp.sa.anneal(ham, seed=seed + sweeps, number_sweeps=sweeps, repetitions=repetitions, only_best=False)
        dt = time.perf_counter() - t0
        acc, ov = common.accuracy_and_overlap_batched(xs, model.initial_signs, weights, model.size)
        err = np.abs((es - e0) / e0)
        rows[int(sweeps)] = {"acc_prob": float(np.mean(acc > 0.995)), "overlap_prob": float(np.mean(ov > 0.995)),
                             "residual_prob": float(np.mean(err <= 1e-12)), "best_energy_minus_E0": float(es.min() - e0),
                             "seconds": dt, "proposals_per_sec": repetitions * sweeps * model.size / dt}
    return {"states": int(model.size), "E0": e0, "repetitions": repetitions, "ladder": rows}


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--systems", default="heisenberg_kagome_16,j1j2_square_4x4,sk_16_3,heisenberg_kagome_18")
    p.add_argument("--sweeps", default="100,200,400,800,1600,3200,6400")
    p.add_argument("--repetitions", type=int, default=1024)
    p.add_argument("--out", default="")
    args = p.parse_args()
    published = json.load(open(os.path.join(ROOT, "tests", "golden", "published_sa_statistics.json")))["systems"]
    result = {}
    for system in args.systems.split(","):
        r = ladder(system, [int(s) for s in args.sweeps.split(",")], args.repetitions)
        result[system] = r
        for sweeps, row in r["ladder"].items():
            pub = published.get(system, {}).get(str(sweeps))
            print("%-22s %6d sweeps: acc %.3f overlap %.3f residual %.3f (best E - E0 %.1e, %.2e proposals/s)%s" % (
                system, sweeps, row["acc_prob"], row["overlap_prob"], row["residual_prob"], row["best_energy_minus_E0"], row["proposals_per_sec"],
                "   published acc %.3f [%.3f, %.3f]" % (pub["acc_prob_mean"], pub["acc_prob_min"], pub["acc_prob_max"]) if pub else ""), flush=True)
    if args.out:
        json.dump(result, open(args.out, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
