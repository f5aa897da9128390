"""Import alias: the package directory is named ``ppo-bihyb_b200`` (not a valid Python identifier), so
``import ppo_bihyb_b200`` resolves here and loads that directory as the package of the same (underscored) name."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ppo-bihyb_b200")
_spec = importlib.util.spec_from_file_location(
    "ppo_bihyb_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_pkg = importlib.util.module_from_spec(_spec)
sys.modules["ppo_bihyb_b200"] = _pkg
_spec.loader.exec_module(_pkg)
