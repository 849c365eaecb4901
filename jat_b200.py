"""Import alias: ``import jat_b200`` loads the package directory
``jatsr-just-audio-transformer-super-solution_b200/`` (hyphens are not importable as a name)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "jatsr-just-audio-transformer-super-solution_b200")
_spec = importlib.util.spec_from_file_location(
    "jat_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["jat_b200"] = _mod
_spec.loader.exec_module(_mod)
