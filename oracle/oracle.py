"""numpy front-end of the CPU oracle (oracle/ifk_oracle.c).

TEST INFRASTRUCTURE ONLY -- the checker for the CUDA path, never the product.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  Each function restates a reference routine; the
citations live next to the C code (oracle/ifk_oracle.c) and in DESIGN.md.

Also holds two independent numpy/float64 cross-checks used to validate the C
restatement itself: a dense block-Toeplitz matrix `dense_L` (the recipe of the
reference's inf/utils/toeplitz.py:9-44, built directly) and `masked_weight`.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libifk_oracle.so")
_lib = None


def build(force=False):
    """Compile oracle/ifk_oracle.c with the committed Makefile (gcc, a few seconds)."""
    src = os.path.join(_HERE, "ifk_oracle.c")
    if (force or not os.path.exists(_LIB_PATH)
            or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.ifk_oracle_max_threads.restype = ctypes.c_int
    return _lib


def max_threads():
    return int(lib().ifk_oracle_max_threads())


def _suffix(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "f32"
    if dtype == np.float64:
        return "f64"
    raise TypeError("oracle supports float32/float64, got %s" % dtype)


def _prep(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _geometry(x, w, groups):
    B, C, H, W = x.shape
    Cw, KH, KW = w.shape[1], w.shape[2], w.shape[3]
    if w.shape[0] != C:
        raise ValueError("weight rows %d != channels %d" % (w.shape[0], C))
    if C % groups or Cw < C // groups:
        raise ValueError("bad groups=%d for C=%d, weight cols %d" % (groups, C, Cw))
    return [ctypes.c_int(v) for v in (B, C, H, W, KH, KW, Cw, groups)]


def _call(name, first, w, groups, threads, second=None, out_like=None):
    dtype = first.dtype
    fn = getattr(lib(), "ifk_oracle_%s_%s" % (name, _suffix(dtype)))
    fn.restype = None
    w = _prep(w, dtype)
    geom = _geometry(first, w, groups)
    out = np.empty_like(first if out_like is None else out_like)
    if second is None:
        fn(_ptr(first), _ptr(w), _ptr(out), *geom, ctypes.c_int(threads))
    else:
        fn(_ptr(first), _ptr(second), _ptr(out), *geom, ctypes.c_int(threads))
    return out


_ORIENT_AXES = {0: (), 1: (3,), 2: (2,), 3: (2, 3), "TL": (), "TR": (3,), "BL": (2,), "BR": (2, 3)}


def _reflect(a, orient):
    """F a: the image reflected along the axes of an orientation (a copy; the reference layers do
    exactly this with torch.flip around their top-left kernels, inf/layers/inv_conv.py:198-214)."""
    axes = _ORIENT_AXES[orient]
    return np.ascontiguousarray(np.flip(a, axes)) if axes else a


def inverse(x, w, groups=1, threads=1, wavefront=False, orient=0):
    """y = F L^-1 F x  (solve_mc.py:88-114; wavefront=True follows solve_mc.py:8-50)."""
    x = _prep(_reflect(x, orient), x.dtype)
    return _reflect(_call("inverse_wavefront" if wavefront else "inverse", x, w, groups, threads), orient)


def conv(y, w, groups=1, threads=1, orient=0):
    """x = F L F y, the masked convolution (sampling direction)."""
    y = _prep(_reflect(y, orient), y.dtype)
    return _reflect(_call("conv", y, w, groups, threads), orient)


def bwd_input(g, w, groups=1, threads=1, orient=0):
    """dX = F L^-T F g."""
    g = _prep(_reflect(g, orient), g.dtype)
    return _reflect(_call("bwd_input", g, w, groups, threads), orient)


def bwd_weight(dx, y, w_shape, groups=1, threads=1, orient=0):
    """dW = -corr(F dX, F y), shaped like the weight (C, Cw, KH, KW)."""
    dx = _prep(_reflect(dx, orient), dx.dtype)
    y = _prep(_reflect(y, orient), dx.dtype)
    w_like = np.empty(w_shape, dtype=dx.dtype)
    fn = getattr(lib(), "ifk_oracle_bwd_weight_%s" % _suffix(dx.dtype))
    fn.restype = None
    geom = _geometry(dx, w_like, groups)
    fn(_ptr(dx), _ptr(y), _ptr(w_like), *geom, ctypes.c_int(threads))
    return w_like


def backward(g, y, w, groups=1, threads=1, orient=0):
    """(dX, dW) for upstream gradient g at saved output y."""
    dx = bwd_input(g, w, groups, threads, orient)
    return dx, bwd_weight(dx, y, w.shape, groups, threads, orient)


def literal_inverse(x, w):
    """Bug-for-bug inv_conv_with_bp.inverse (inv_conv_with_bp_kernel_general.cu:52-65)."""
    x = _prep(x, x.dtype)
    w = _prep(w, x.dtype)
    fn = getattr(lib(), "ifk_oracle_literal_inverse_%s" % _suffix(x.dtype))
    fn.restype = None
    out = np.empty_like(x)
    fn(_ptr(x), _ptr(w), _ptr(out), *_geometry(x, w, 1)[:7])
    return out


def literal_dy(g, w):
    """Bug-for-bug inv_conv_with_bp.dy (= L^-1 g; .cu:307-327, 371-383)."""
    g = _prep(g, g.dtype)
    w = _prep(w, g.dtype)
    fn = getattr(lib(), "ifk_oracle_literal_dy_%s" % _suffix(g.dtype))
    fn.restype = None
    out = np.empty_like(g)
    fn(_ptr(g), _ptr(w), _ptr(out), *_geometry(g, w, 1)[:7])
    return out


# ----------------------------------------------------------------------------------
# Independent float64 cross-checks (numpy only; small shapes).
# ----------------------------------------------------------------------------------
def masked_weight(w, groups=1):
    """Effective conv weight (C, Cg, KH, KW): unit diagonal at the centre tap, centre-tap
    upper triangle zeroed -- what F.conv2d(F.pad(y, (KW-1,0,KH-1,0)), ., groups) needs to
    equal L y."""
    w = np.asarray(w, dtype=np.float64)
    C, _, KH, KW = w.shape
    Cg = C // groups
    m = w[:, :Cg].copy()
    for c in range(C):
        cl = c % Cg
        m[c, cl:, KH - 1, KW - 1] = 0.0
        m[c, cl, KH - 1, KW - 1] = 1.0
    return m


def dense_L(w, H, W, groups=1):
    """Dense (C*H*W)^2 matrix of the masked causal convolution for ONE image."""
    m = masked_weight(w, groups)
    C, Cg, KH, KW = m.shape
    n = C * H * W
    L = np.zeros((n, n), dtype=np.float64)
    for c in range(C):
        base = (c // Cg) * Cg
        for h in range(H):
            for x in range(W):
                row = (c * H + h) * W + x
                for kc in range(Cg):
                    for qh in range(min(KH, h + 1)):
                        for qw in range(min(KW, x + 1)):
                            col = ((base + kc) * H + h - qh) * W + x - qw
                            L[row, col] = m[c, kc, KH - 1 - qh, KW - 1 - qw]
    return L


def max_rel_err(a, ref):
    """max |a - ref| / max |ref|: the parity metric used throughout tests/ and bench.py.  A MAX-NORM relative
    error (one denominator for the whole tensor), not an elementwise one -- see max_elem_rel_err."""
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    den = float(np.max(np.abs(ref))) if ref.size else 0.0
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - ref))) / (den if den > 0 else 1.0)


def max_elem_rel_err(a, ref, floor=1e-3):
    """elementwise relative error max |a_i - ref_i| / |ref_i| over the entries with |ref_i| >= floor * max|ref|
    (entries below the floor are covered by the max-norm metric only: their relative error is dominated by
    the float32 rounding of terms far larger than they are)."""
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    if a.size == 0:
        return 0.0
    den = float(np.max(np.abs(ref)))
    if den == 0.0:
        return float(np.max(np.abs(a)))
    big = np.abs(ref) >= floor * den
    return float(np.max(np.abs(a[big] - ref[big]) / np.abs(ref[big])))
