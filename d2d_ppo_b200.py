"""Importable alias for the package directory ``d2d-ppo_b200/`` (a hyphen is not a valid module name).

``import d2d_ppo_b200`` loads ``d2d-ppo_b200/__init__.py`` as the package ``d2d_ppo_b200`` so that
``from d2d_ppo_b200.envs import CombinatorialEnv`` works from the repo root.
"""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "d2d-ppo_b200")
_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
