"""Tiny helper importable WITHOUT importing the package (which needs the .so)."""
import importlib.util
import os

_ROOT = os.path.dirname(os.path.abspath(__file__))


def _load_build_module():
    spec = importlib.util.spec_from_file_location(
        "_b200_build", os.path.join(_ROOT, "librecommender_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def ensure_built(force: bool = False, verbose: bool = False) -> str:
    return _load_build_module().build(force=force, verbose=verbose)
