"""B200-native drop-in for the hot path of twesterhout/annealing-sign-problem:
Ising-model extraction (``make_ising_model``) and replica simulated annealing
(``solve_ising_model``), behind the reference's own Python entry points
(annealing_sign_problem/common.py:131-261)."""
__version__ = "0.1.0"

from .common import (  # noqa: F401
    IsingModel,
    SamplingResult,
    add_noise_to_amplitudes,
    amplitude_overlap,
    create_small_cluster_around_point,
    determine_exact_solution,
    ground_state_to_log_coeff_fn,
    monte_carlo_sampling,
    binary_search,
    compute_accuracy_and_overlap,
    get_strongest_off_diag,
    load_hamiltonian,
    make_hamiltonian_extension,
    make_ising_model,
    solve_ising_model,
    sparsify_using_global_cutoff,
)
from . import annealer as sa  # noqa: F401  (plays the role of `import ising_glass_annealer as sa`)
from . import symmetries as ls  # noqa: F401  (plays the role of `import lattice_symmetries as ls`)
