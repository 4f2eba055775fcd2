"""Sequential restatement of the greedy solver the reference PRESERVES as commented Python
(annealing_sign_problem/common.py:298-438, `strongest_coupling_greedy_color`) -- TEST INFRASTRUCTURE ONLY.

The live reference calls `ising_glass_annealer.greedy_solve` (third-party Haskell, absent: parity unpinned);
this file follows the preserved Python rule by rule, INCLUDING the two rules the product's parallel solver
(csrc/greedy.cu, oracle/greedy_port.c) deviates from, so that the deviation can be measured
(tests/test_oracle.py::test_greedy_deviation_from_the_preserved_algorithm_is_bounded):

  * edges in descending |J| (argsort(|data|)[::-1], s1 < s2; common.py:313-320);
  * two free spins: s1 = +1, s2 = -sign(J) (common.py:398-404);
  * ONE free spin joining a cluster: +1, flipped when its merge energy -- the sum over ALL its couplings into
    the cluster, sign_cluster * J -- is positive (common.py:374-396; `merge_energy`, common.py:333-345).
    (As written, `zip(cluster.spins, cluster.signs)` pairs a spin with a dict KEY; the intent -- the sign -- is
    what is restated here);
  * two clusters: the second one is flipped when the JOINING edge is frustrated (common.py:359-372:
    `should_flip = is_frustrated`; the merge-energy variant is commented out there);
  * descent: sweeps over the spins in the insertion order of the surviving cluster's sign dictionary, a spin
    with positive local energy e = s_i sum_j J_ij s_j is flipped at once; until a sweep flips nothing
    (common.py:417-433).
The preserved code asserts a single connected component; here every component is solved the same way.
"""
import numpy as np
import scipy.sparse


def greedy_reference(exchange: scipy.sparse.spmatrix):
    """-> (signs float64[n] of +-1, number of descent sweeps)."""
    coo = scipy.sparse.coo_matrix(exchange).copy()
    coo.setdiag(np.zeros(coo.shape[0]))
    coo.eliminate_zeros()
    csr = coo.tocsr()
    csr.sort_indices()
    coo = csr.tocoo()
    n = coo.shape[0]
    order = np.argsort(np.abs(coo.data), kind="stable")[::-1]
    cluster_of = {}   # spin -> cluster id (insertion order = the order spins were first clustered)
    members = {}      # cluster id -> dict spin -> sign (insertion ordered)
    next_id = 0

    def merge_energy_single(spin, cluster):
        energy = 0.0
        for k in range(csr.indptr[spin], csr.indptr[spin + 1]):
            other = csr.indices[k]
            if other in cluster:
                energy += cluster[other] * csr.data[k]
        return energy

    for k in order:
        s1, s2 = int(coo.row[k]), int(coo.col[k])
        if not s1 < s2:
            continue
        coupling = float(coo.data[k])
        in1, in2 = s1 in cluster_of, s2 in cluster_of
        if in1 and in2:
            c1, c2 = cluster_of[s1], cluster_of[s2]
            if c1 == c2:
                continue
            flip = members[c1][s1] * members[c2][s2] * coupling > 0
            for key in list(cluster_of.keys()):  # the reference walks its dict in insertion order
                if cluster_of[key] == c2:
                    sign = members[c2][key]
                    cluster_of[key] = c1
                    members[c1][key] = -sign if flip else sign
            del members[c2]
        elif in1 or in2:
            inside, free = (s1, s2) if in1 else (s2, s1)
            c = cluster_of[inside]
            sign = -1.0 if merge_energy_single(free, members[c]) > 0 else 1.0
            cluster_of[free] = c
            members[c][free] = sign
        else:
            members[next_id] = {s1: 1.0, s2: -float(np.sign(coupling))}
            cluster_of[s1] = cluster_of[s2] = next_id
            next_id += 1
    for i in range(n):  # isolated spins
        if i not in cluster_of:
            members[next_id] = {i: 1.0}
            cluster_of[i] = next_id
            next_id += 1
    signs = {}
    for c in members.values():
        signs.update(c)
    sweeps = 0
    while True:
        changed = False
        sweeps += 1
        for s1 in signs.keys():
            e = 0.0
            for k in range(csr.indptr[s1], csr.indptr[s1 + 1]):
                e += signs[csr.indices[k]] * csr.data[k]
            if e * signs[s1] > 0:
                changed = True
                signs[s1] = -signs[s1]
        if not changed:
            break
    return np.array([signs[i] for i in range(n)]), sweeps
