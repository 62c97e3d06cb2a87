"""Import shim: ``import pstl_b200`` loads the package that lives in ``pstl-diffusion-policy_b200/``.

The package directory keeps the name the build contract asks for; a hyphen cannot be
imported, so this one-file loader registers that directory as the package ``pstl_b200``.
"""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pstl-diffusion-policy_b200")
_spec = importlib.util.spec_from_file_location(
    "pstl_b200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules["pstl_b200"] = _mod
_spec.loader.exec_module(_mod)
