"""Importable alias of the product package.

The package directory is named ``soft-intro-vae-for-3d-mri_b200`` (hyphens are not valid in a Python
module name), so ``import sivae_b200`` loads that directory as the package ``sivae_b200``.
"""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "soft-intro-vae-for-3d-mri_b200")
_spec = importlib.util.spec_from_file_location(
    "sivae_b200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["sivae_b200"] = _mod
_spec.loader.exec_module(_mod)
