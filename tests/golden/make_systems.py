"""Restate the reference's physical-system definitions as compact JSON input files.

Run in the build container only (needs /root/reference):

    python tests/golden/make_systems.py

/root/reference does not exist on the GPU box, and the bond lists / symmetry generators
of the named shapes are needed there by ``bench.py`` and the ``-m gpu``
tests.  This script parses physical_systems/*.yaml (reference schema, read by
annealing_sign_problem/common.py:782-787) and writes the same numbers, same schema, as
JSON under ``annealing-sign-problem_b200/systems/``.  Data only -- no reference source.
"""
import json
import os
import sys

import yaml

REF = "/root/reference/physical_systems"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "..", "annealing-sign-problem_b200", "systems")

NAMES = [
    "j1j2_square_4x4",
    "heisenberg_kagome_16",
    "heisenberg_kagome_18",
    "heisenberg_kagome_36",
    "heisenberg_pyrochlore_2x2x2",
    "sk_16_1",
    "sk_16_2",
    "sk_16_3",
    "sk_32_1",
]


def main():
    os.makedirs(OUT, exist_ok=True)
    for name in NAMES:
        with open(os.path.join(REF, name + ".yaml")) as f:
            cfg = yaml.load(f, Loader=yaml.SafeLoader)
        basis = cfg["basis"]
        slim = {
            "source": "physical_systems/%s.yaml" % name,
            "basis": {
                "number_spins": int(basis["number_spins"]),
                "hamming_weight": basis.get("hamming_weight"),
                "spin_inversion": basis.get("spin_inversion"),
                "symmetries": [
                    {"permutation": [int(v) for v in s["permutation"]], "sector": int(s["sector"])}
                    for s in (basis.get("symmetries") or [])
                ],
            },
            "hamiltonian": {
                "name": cfg["hamiltonian"].get("name", name),
                "terms": [
                    {
                        "matrix": [[float(v) for v in row] for row in t["matrix"]],
                        "sites": [[int(a), int(b)] for a, b in t["sites"]],
                    }
                    for t in cfg["hamiltonian"]["terms"]
                ],
            },
        }
        path = os.path.join(OUT, name + ".json")
        with open(path, "w") as f:
            json.dump(slim, f, separators=(",", ":"))
        nb = sum(len(t["sites"]) for t in slim["hamiltonian"]["terms"])
        print("%-32s spins=%d bonds=%d generators=%d -> %d bytes" % (
            name, slim["basis"]["number_spins"], nb, len(slim["basis"]["symmetries"]), os.path.getsize(path)))


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference (build container only)")
    main()
