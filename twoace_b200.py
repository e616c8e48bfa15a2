"""Import shim: the product package lives in ``2ace-mmwave-channel-estimation_b200/`` (a directory
name that is not a Python identifier); ``import twoace_b200`` loads it under this name."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "2ace-mmwave-channel-estimation_b200")
_spec = importlib.util.spec_from_file_location(
    "twoace_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["twoace_b200"] = _mod
_spec.loader.exec_module(_mod)
