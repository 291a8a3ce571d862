"""Import shim: the product package directory is named ``homogenization.jl_b200`` (with a dot),
which Python cannot import by name.  ``import hmgb200`` loads it under the module name
``homogenization_jl_b200`` and re-exports it."""
import importlib.util
import os
import sys

_NAME = "homogenization_jl_b200"
if _NAME not in sys.modules:
    _dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "homogenization.jl_b200")
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_dir, "__init__.py"),
                                                   submodule_search_locations=[_dir])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)
_mod = sys.modules[_NAME]
globals().update({k: v for k, v in vars(_mod).items() if not k.startswith("__")})
