"""Loader for oracle/_ref/*.so -- the reference's own CUDA kernels compiled from
/root/reference by oracle/build_ref.py.  TEST INFRASTRUCTURE ONLY; needs a GPU to run.
Returns None when a module was not built (tests then skip the reference comparison and rely
on the committed golden vectors those modules produced)."""
import importlib.util
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_cache = {}


def load(name):
    if name in _cache:
        return _cache[name]
    path = os.path.join(_HERE, "_ref", name + ".so")
    mod = None
    if os.path.exists(path):
        import torch  # noqa: F401  (the .so links against libtorch)
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    _cache[name] = mod
    return mod


def nerfacc_cuda():
    return load("nerfacc_cuda")


def renderutils_plugin():
    return load("renderutils_plugin")
