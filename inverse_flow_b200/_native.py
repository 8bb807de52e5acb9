"""ctypes binding of the C ABI (include/ifk.h) -> inverse_flow_b200/lib/libifk_b200.so.

torch is used here only for device memory and the current stream; every compute call goes
through the C ABI with raw device pointers.  ctypes releases the GIL for the duration of
each foreign call (the reference's pybind module holds it).

There is NO fallback: if the library is missing or fails to load this raises, and so does
every op on a machine without a CUDA device.
"""
import ctypes
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
# IFK_LIBRARY: another build of the same ABI (development: tools/variant_build.py A/B-tests compile-time variants)
LIB_PATH = os.environ.get("IFK_LIBRARY") or os.path.join(_PKG, "lib", "libifk_b200.so")

EXPORTS = (
    "ifk_version", "ifk_status_string", "ifk_prepared_floats", "ifk_prepare_f32", "ifk_prepare_many_f32",
    "ifk_inverse_f32", "ifk_conv_f32", "ifk_bwd_input_f32", "ifk_bwd_weight_workspace_bytes",
    "ifk_bwd_weight_f32", "ifk_bwd_weight_partial_f32", "ifk_bwd_weight_reduce_many_f32",
    "ifk_backward_f32", "ifk_describe_solve", "ifk_inverse_once_f32", "ifk_inverse_chain_f32",
    "ifk_inverse_probe_f32", "ifk_debug_reload_env", "ifk_debug_fp32_peak", "ifk_debug_latencies",
    "ifk_allreduce_flag_bytes", "ifk_allreduce_peer_f32", "ifk_inverse_fused_f32", "ifk_bwd_input_fused_f32",
)

FLAG_STABLE_PREPARED = 1        # enum ifk_flags
ERR_BAD_LAYOUT = -7              # enum ifk_status
ERR_UNSUPPORTED = -4


class Problem(ctypes.Structure):
    """struct ifk_problem"""
    _fields_ = [(n, ctypes.c_int) for n in ("B", "C", "H", "W", "KH", "KW", "Cw", "groups", "orient", "flags")]


class Fused(ctypes.Structure):
    """struct ifk_fused: the ActNorm affine / Squeeze re-indexing fused into a solve's load and store"""
    _fields_ = [("in_scale", ctypes.c_void_p), ("in_bias", ctypes.c_void_p), ("out_scale", ctypes.c_void_p),
                ("squeeze", ctypes.c_int)]


class IfkError(RuntimeError):
    pass


_lib = None


def load():
    """Load the shared library (once).  Raises IfkError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise IfkError(
            "native library %s not found: run `python -m inverse_flow_b200.build` "
            "(there is no CPU or PyTorch fallback for this path)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    P, vp, sz, ci = ctypes.POINTER(Problem), ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
    lib.ifk_version.restype = ci
    lib.ifk_status_string.restype = ctypes.c_char_p
    lib.ifk_status_string.argtypes = [ci]
    lib.ifk_prepared_floats.restype = sz
    lib.ifk_prepared_floats.argtypes = [P]
    lib.ifk_bwd_weight_workspace_bytes.restype = sz
    lib.ifk_bwd_weight_workspace_bytes.argtypes = [P]
    for name, nptr in (("ifk_prepare_f32", 2), ("ifk_inverse_f32", 3), ("ifk_conv_f32", 3),
                       ("ifk_bwd_input_f32", 3), ("ifk_bwd_weight_f32", 4), ("ifk_backward_f32", 6)):
        fn = getattr(lib, name)
        fn.restype = ci
        fn.argtypes = [P] + [vp] * nptr + [vp]          # ..., stream
    lib.ifk_prepare_many_f32.restype = ci
    lib.ifk_prepare_many_f32.argtypes = [P, ci, vp, sz, vp, sz, vp]
    lib.ifk_bwd_weight_partial_f32.restype = ci
    lib.ifk_bwd_weight_partial_f32.argtypes = [P, vp, vp, vp, vp]
    lib.ifk_bwd_weight_reduce_many_f32.restype = ci
    lib.ifk_bwd_weight_reduce_many_f32.argtypes = [P, ci, vp, sz, vp, sz, vp]
    lib.ifk_describe_solve.restype = ci
    lib.ifk_describe_solve.argtypes = [P, ctypes.c_char_p, sz]
    lib.ifk_inverse_once_f32.restype = ci
    lib.ifk_inverse_once_f32.argtypes = [P, vp, vp, vp, vp, vp]
    lib.ifk_inverse_chain_f32.restype = ci
    lib.ifk_inverse_chain_f32.argtypes = [P, ci, ctypes.POINTER(ci), ctypes.POINTER(vp), vp, ctypes.POINTER(vp), vp]
    lib.ifk_inverse_probe_f32.restype = ci
    lib.ifk_inverse_probe_f32.argtypes = [P, vp, vp, vp, vp, vp]
    lib.ifk_inverse_fused_f32.restype = ci
    lib.ifk_inverse_fused_f32.argtypes = [P, ctypes.POINTER(Fused), vp, vp, vp, vp]
    lib.ifk_bwd_input_fused_f32.restype = ci
    lib.ifk_bwd_input_fused_f32.argtypes = [P, ctypes.POINTER(Fused), vp, vp, vp, vp, vp]
    lib.ifk_debug_reload_env.restype = None
    lib.ifk_debug_reload_env.argtypes = []
    lib.ifk_allreduce_flag_bytes.restype = sz
    lib.ifk_allreduce_flag_bytes.argtypes = []
    lib.ifk_allreduce_peer_f32.restype = ci
    lib.ifk_allreduce_peer_f32.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(vp), ci, ci, vp, sz, vp]
    lib.ifk_debug_fp32_peak.restype = ci
    lib.ifk_debug_fp32_peak.argtypes = [vp, ctypes.POINTER(ctypes.c_double)]
    lib.ifk_debug_latencies.restype = ci
    lib.ifk_debug_latencies.argtypes = [vp, ctypes.POINTER(ctypes.c_double)]
    _lib = lib
    return lib


def check(status):
    if status != 0:
        msg = load().ifk_status_string(status).decode()
        if status < 0:
            raise ValueError("ifk: %s (status %d)" % (msg, status))
        raise IfkError("ifk: CUDA error %d: %s" % (status, msg))


ORIENTS = {"TL": 0, "TR": 1, "BL": 2, "BR": 3}      # enum ifk_orient: bit 0 reflects W, bit 1 reflects H


def orient_code(orient):
    """'TL' / 'TR' / 'BL' / 'BR' (the reference layers' `order`) or the IFK_ORIENT_* integer"""
    if isinstance(orient, str):
        if orient not in ORIENTS:
            raise ValueError("unknown order: %r" % (orient,))
        return ORIENTS[orient]
    return int(orient)


def problem(B, C, H, W, KH, KW, Cw, groups, orient=0, flags=0):
    return Problem(int(B), int(C), int(H), int(W), int(KH), int(KW), int(Cw), int(groups), orient_code(orient),
                   int(flags))


def with_flags(p, flags):
    """a copy of the problem with other IFK_FLAG_* bits"""
    return Problem(p.B, p.C, p.H, p.W, p.KH, p.KW, p.Cw, p.groups, p.orient, int(flags))


def reload_env():
    """re-read the IFK_* knobs (they are cached by the library; tests pin kernels through them)"""
    load().ifk_debug_reload_env()


def hw_microbench(device=None):
    """measured roofline denominators of this GPU: FP32 FMA rate (TFLOP/s) and the latencies (cycles) of the
    instructions a wavefront step chains.  Synchronises; a measuring aid for bench.py / tools."""
    lib = load()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    with torch.cuda.device(device):
        buf = torch.zeros(16, dtype=torch.int64, device=device)
        tf = (ctypes.c_double * 2)()
        lat = (ctypes.c_double * 8)()
        torch.cuda.synchronize()
        check(lib.ifk_debug_fp32_peak(ctypes.c_void_p(buf.data_ptr()), tf))
        check(lib.ifk_debug_latencies(ctypes.c_void_p(buf.data_ptr()), lat))
    names = ["ffma", "ffma2", "shfl", "lds", "sts_syncwarp_lds", "sts_barsync8_lds", "barsync8", "fadd"]
    return {"fp32_tflops_ffma": tf[0], "fp32_tflops_ffma2": tf[1],
            "latency_cycles": {n: lat[i] for i, n in enumerate(names)}}


def current_stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def describe_solve(p):
    buf = ctypes.create_string_buffer(256)
    check(load().ifk_describe_solve(ctypes.byref(p), buf, 256))
    return buf.value.decode()
