"""Loader for the reference's own CUDA extension compiled by oracle/build_ref_cuda.py.

Test infrastructure only (tests/, tools/ref_gpu_bench.py); the product never imports this.
The module exposes the reference's four functions unchanged
(inf/utils/inv_conv_cuda/inv_conv_with_bp_general.cpp:115-120):
    inverse(input, kernel, output), forward(input, kernel, output),
    dy(grad_out, kernel, M, output), dw(input, kernel, grad_out, M, output)
each returning [output] (caller-allocated, pre-zeroed).
"""
import importlib.util
import os

from . import build_ref_cuda

_mod = None


def available():
    return os.path.exists(build_ref_cuda.so_path())


def load():
    """the compiled reference extension, or None when oracle/_ref/ does not hold it"""
    global _mod
    if _mod is None and available():
        import torch  # noqa: F401  (libtorch must be loaded before the extension)
        spec = importlib.util.spec_from_file_location(build_ref_cuda.NAME, build_ref_cuda.so_path())
        _mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_mod)
    return _mod
