"""Tensor-level entry points over the C ABI: validation, buffer allocation, launch.

These mirror the reference's pybind functions `inverse / forward / dy / dw`
(inf/utils/inv_conv_cuda/inv_conv_with_bp_general.cpp:19-28, 44-53, 70-81, 99-112) with the
math-contract semantics of SURVEY.md section 8a and an explicit `groups`.
"""
import ctypes

import torch

from . import _native


def default_groups(C):
    """The reference kernels hard-code 4 channel groups (inv_conv_with_bp_kernel_general.cu:
    94-98) and write nothing when C < 4; this is their evident intent (cinc_kernel_level2.cu)
    where it is defined and full coupling otherwise."""
    return 4 if C % 4 == 0 else 1


def _check_activation(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:                      # reference: CHECK_CUDA -> RuntimeError (.cpp:15)
        raise RuntimeError("%s must be a CUDA tensor" % name)
    if t.dtype != torch.float32:
        raise RuntimeError("%s must be float32, got %s" % (name, t.dtype))
    if t.dim() != 4:
        raise ValueError("%s must be 4-D (B, C, H, W), got shape %s" % (name, tuple(t.shape)))
    if not t.is_contiguous():              # reference: CHECK_CONTIGUOUS -> RuntimeError (.cpp:16)
        raise RuntimeError("%s must be contiguous" % name)


def _problem(x, weight, groups, orient=0):
    _check_activation(x, "input")
    _check_activation(weight, "kernel")
    if weight.device != x.device:
        raise RuntimeError("input and kernel live on different devices")
    B, C, H, W = x.shape
    if weight.shape[0] != C:
        raise ValueError("kernel has %d output rows, input has %d channels" % (weight.shape[0], C))
    if groups is None:
        groups = default_groups(C)
    if min(C, H, W, weight.shape[1], weight.shape[2], weight.shape[3]) == 0:
        raise ValueError("zero-sized channel / spatial / kernel dimension")
    return _native.problem(B, C, H, W, weight.shape[2], weight.shape[3], weight.shape[1], groups, orient)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


class Prepared:
    """Prepared (T-folded) weights of one weight tensor: serves inverse and backward of any
    batch / image size with that weight."""

    def __init__(self, weight, groups=None):
        lib = _native.load()
        _check_activation(weight, "kernel")
        C, Cw, KH, KW = weight.shape
        if min(C, Cw, KH, KW) == 0:
            raise ValueError("zero-sized kernel dimension")
        self.groups = default_groups(C) if groups is None else int(groups)
        self.weight_shape = (C, Cw, KH, KW)
        self.device = weight.device
        p = _native.problem(0, C, 1, 1, KH, KW, Cw, self.groups)
        n = lib.ifk_prepared_floats(ctypes.byref(p))
        self.buffer = torch.empty(max(n, 1), dtype=torch.float32, device=weight.device)
        with torch.cuda.device(weight.device):
            _native.check(lib.ifk_prepare_f32(ctypes.byref(p), _ptr(weight), _ptr(self.buffer),
                                              _native.current_stream(weight.device)))

    def for_batch(self, x, orient=0):
        C, Cw, KH, KW = self.weight_shape
        if x.shape[1] != C:
            raise ValueError("kernel has %d output rows, input has %d channels" % (C, x.shape[1]))
        if x.device != self.device:
            raise RuntimeError("input and kernel live on different devices")
        if x.shape[2] == 0 or x.shape[3] == 0:
            raise ValueError("zero-sized spatial dimension")
        return _native.problem(x.shape[0], C, x.shape[2], x.shape[3], KH, KW, Cw, self.groups, orient)


def prepare(weight, groups=None):
    return Prepared(weight, groups)


def inverse(x, weight, groups=None, out=None, prepared=None, orient=0):
    """y = L^-1 x (training direction).  `out` may be supplied (reference call style).
    `orient` ('TL'/'TR'/'BL'/'BR' or IFK_ORIENT_*): the corner the causal support grows from; every
    function of this module then computes F (.) F with F the reflection, inside the kernel."""
    lib = _native.load()
    if prepared is None:
        prepared = Prepared(weight, groups)
    _check_activation(x, "input")
    p = prepared.for_batch(x, orient)
    if out is None:
        out = torch.empty_like(x)
    else:
        _check_activation(out, "output")
        if out.shape != x.shape or out.device != x.device:
            raise ValueError("output must match input shape/device")
        if out.data_ptr() == x.data_ptr() and x.numel():
            raise ValueError("input and output must not alias")
    with torch.cuda.device(x.device):
        _native.check(lib.ifk_inverse_f32(ctypes.byref(p), _ptr(x), _ptr(prepared.buffer), _ptr(out),
                                          _native.current_stream(x.device)))
    return out


def conv(y, weight, groups=None, out=None, orient=0):
    """x = L y (sampling direction); takes the raw weight."""
    lib = _native.load()
    p = _problem(y, weight, groups, orient)
    if out is None:
        out = torch.empty_like(y)
    else:
        _check_activation(out, "output")
        if out.shape != y.shape or out.device != y.device:
            raise ValueError("output must match input shape/device")
        if out.data_ptr() == y.data_ptr() and y.numel():
            raise ValueError("input and output must not alias")
    with torch.cuda.device(y.device):
        _native.check(lib.ifk_conv_f32(ctypes.byref(p), _ptr(y), _ptr(weight), _ptr(out),
                                       _native.current_stream(y.device)))
    return out


def bwd_input(grad, weight, groups=None, out=None, prepared=None, orient=0):
    """dX = L^-T grad."""
    lib = _native.load()
    if prepared is None:
        prepared = Prepared(weight, groups)
    _check_activation(grad, "grad_output")
    p = prepared.for_batch(grad, orient)
    if out is None:
        out = torch.empty_like(grad)
    with torch.cuda.device(grad.device):
        _native.check(lib.ifk_bwd_input_f32(ctypes.byref(p), _ptr(grad), _ptr(prepared.buffer), _ptr(out),
                                            _native.current_stream(grad.device)))
    return out


def bwd_weight(dx, y, weight, groups=None, out=None, orient=0):
    """dW = -corr(dX, y), shaped like `weight`."""
    lib = _native.load()
    p = _problem(dx, weight, groups, orient)
    _check_activation(y, "saved output")
    if y.shape != dx.shape:
        raise ValueError("dx and y shapes differ")
    if out is None:
        out = torch.empty_like(weight)
    nbytes = lib.ifk_bwd_weight_workspace_bytes(ctypes.byref(p))
    ws = torch.empty((nbytes + 3) // 4, dtype=torch.float32, device=dx.device)
    with torch.cuda.device(dx.device):
        _native.check(lib.ifk_bwd_weight_f32(ctypes.byref(p), _ptr(dx), _ptr(y), _ptr(out), _ptr(ws),
                                             _native.current_stream(dx.device)))
    return out


def backward(grad, y, weight, groups=None, prepared=None, orient=0):
    """(dX, dW) for upstream `grad` at the saved output `y` -- one C call."""
    lib = _native.load()
    if prepared is None:
        prepared = Prepared(weight, groups)
    _check_activation(grad, "grad_output")
    _check_activation(y, "saved output")
    if y.shape != grad.shape:
        raise ValueError("grad_output and saved output shapes differ")
    p = prepared.for_batch(grad, orient)
    dx = torch.empty_like(grad)
    dw = torch.empty_like(weight)
    nbytes = lib.ifk_bwd_weight_workspace_bytes(ctypes.byref(p))
    ws = torch.empty((nbytes + 3) // 4, dtype=torch.float32, device=grad.device)
    with torch.cuda.device(grad.device):
        _native.check(lib.ifk_backward_f32(ctypes.byref(p), _ptr(grad), _ptr(y), _ptr(prepared.buffer),
                                           _ptr(dx), _ptr(dw), _ptr(ws),
                                           _native.current_stream(grad.device)))
    return dx, dw
