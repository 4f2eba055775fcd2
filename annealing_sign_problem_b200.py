"""Import shim: ``import annealing_sign_problem_b200`` loads the package that lives in the
directory ``annealing-sign-problem_b200/`` (a hyphen is not importable as a module name)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "annealing-sign-problem_b200")
_spec = importlib.util.spec_from_file_location(
    "annealing_sign_problem_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["annealing_sign_problem_b200"] = _mod
_spec.loader.exec_module(_mod)
