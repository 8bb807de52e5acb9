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
        if t.is_contiguous(memory_format=torch.channels_last):
            # IFK_ERR_BAD_LAYOUT (ifk.h): the kernels read NCHW; a channels_last tensor handed over as a raw
            # pointer would be read as garbage, so it is refused here instead
            raise RuntimeError("%s is channels_last (NHWC); the inverse-conv kernels take NCHW-contiguous tensors: "
                               "call .contiguous() first (ifk status %d)" % (name, _native.ERR_BAD_LAYOUT))
        raise RuntimeError("%s must be contiguous" % name)


def _check_out(out, like, name):
    """a caller-allocated output (the reference's call style, inv_conv_with_bp_general.cpp:19-28): same shape,
    dtype, device as `like`, contiguous, not aliasing it"""
    _check_activation(out, name)
    if out.shape != like.shape or out.device != like.device:
        raise ValueError("%s must match the shape and device of its input: %s on %s vs %s on %s" % (
            name, tuple(out.shape), out.device, tuple(like.shape), like.device))
    if out.data_ptr() == like.data_ptr() and like.numel():
        raise ValueError("%s must not alias its input" % name)


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

    def check_weight(self, weight):
        """the prepared buffer belongs to ONE weight tensor's shape"""
        if weight is not None and tuple(weight.shape) != self.weight_shape:
            raise ValueError("prepared weights are for a kernel of shape %s, got %s" % (
                self.weight_shape, tuple(weight.shape)))

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
    else:
        prepared.check_weight(weight)
    _check_activation(x, "input")
    p = prepared.for_batch(x, orient)
    if out is None:
        out = torch.empty_like(x)
    else:
        _check_out(out, x, "output")
    with torch.cuda.device(x.device):
        _native.check(lib.ifk_inverse_f32(ctypes.byref(p), _ptr(x), _ptr(prepared.buffer), _ptr(out),
                                          _native.current_stream(x.device)))
    return out


def inverse_chain(x, prepared_list, orients, outs=None):
    """Consecutive inverse-conv layers that feed each other -- y_0 = L_0^-1 x, y_i = L_i^-1 y_{i-1} -- as ONE launch
    (ifk_inverse_chain_f32): the image stays in shared memory from layer to layer; every y_i is still written
    (the backward needs it).  `prepared_list`: one Prepared per layer (same kernel shape and groups), `orients`:
    the layers' orders.  Returns the list of outputs, bit-identical to calling `inverse` per layer.  Geometries
    without a resident chain kernel are served by that per-layer loop."""
    lib = _native.load()
    n = len(prepared_list)
    if n == 0 or len(orients) != n:
        raise ValueError("need one orientation per layer")
    _check_activation(x, "input")
    first = prepared_list[0]
    if any(pr.weight_shape != first.weight_shape or pr.groups != first.groups for pr in prepared_list):
        raise ValueError("the layers of a chain share one kernel shape and grouping")
    if outs is None:
        outs = [torch.empty_like(x) for _ in range(n)]
    else:
        if len(outs) != n:
            raise ValueError("need one output per layer")
        for o in outs:
            _check_out(o, x, "output")
    p = first.for_batch(x, 0)
    codes = (ctypes.c_int * n)(*[_native.orient_code(o) for o in orients])
    preps = (ctypes.c_void_p * n)(*[pr.buffer.data_ptr() for pr in prepared_list])
    ys = (ctypes.c_void_p * n)(*[o.data_ptr() for o in outs])
    with torch.cuda.device(x.device):
        status = lib.ifk_inverse_chain_f32(ctypes.byref(p), n, codes, preps, _ptr(x), ys, _native.current_stream(x.device))
        if status == _native.ERR_UNSUPPORTED:              # no resident kernel for this geometry: layer by layer
            cur = x
            for pr, o, out in zip(prepared_list, orients, outs):
                q = pr.for_batch(cur, o)
                _native.check(lib.ifk_inverse_f32(ctypes.byref(q), _ptr(cur), _ptr(pr.buffer), _ptr(out),
                                                  _native.current_stream(x.device)))
                cur = out
        else:
            _native.check(status)
    return outs


def conv(y, weight, groups=None, out=None, orient=0):
    """x = L y (sampling direction); takes the raw weight."""
    lib = _native.load()
    p = _problem(y, weight, groups, orient)
    if out is None:
        out = torch.empty_like(y)
    else:
        _check_out(out, y, "output")
    with torch.cuda.device(y.device):
        _native.check(lib.ifk_conv_f32(ctypes.byref(p), _ptr(y), _ptr(weight), _ptr(out),
                                       _native.current_stream(y.device)))
    return out


def bwd_input(grad, weight, groups=None, out=None, prepared=None, orient=0):
    """dX = L^-T grad."""
    lib = _native.load()
    if prepared is None:
        prepared = Prepared(weight, groups)
    else:
        prepared.check_weight(weight)
    _check_activation(grad, "grad_output")
    p = prepared.for_batch(grad, orient)
    if out is None:
        out = torch.empty_like(grad)
    else:
        _check_out(out, grad, "output")
    with torch.cuda.device(grad.device):
        _native.check(lib.ifk_bwd_input_f32(ctypes.byref(p), _ptr(grad), _ptr(prepared.buffer), _ptr(out),
                                            _native.current_stream(grad.device)))
    return out


def bwd_weight(dx, y, weight, groups=None, out=None, orient=0):
    """dW = -corr(dX, y), shaped like `weight`."""
    lib = _native.load()
    p = _problem(dx, weight, groups, orient)
    _check_activation(y, "saved output")
    if y.shape != dx.shape or y.device != dx.device:
        raise ValueError("dx and y shapes / devices differ")
    if out is None:
        out = torch.empty_like(weight)
    else:
        _check_activation(out, "output")
        if out.shape != weight.shape or out.device != weight.device:
            raise ValueError("the weight gradient must have the kernel's shape %s on %s, got %s on %s" % (
                tuple(weight.shape), weight.device, tuple(out.shape), out.device))
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
    else:
        prepared.check_weight(weight)
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


# ---- elementwise neighbours fused into the solve (ifk.h: ifk_fused) ------------------
def _check_vec(v, C, device, name):
    if v is None:
        return None
    if not (isinstance(v, torch.Tensor) and v.is_cuda and v.dtype == torch.float32 and v.dim() == 1 and
            v.numel() == C and v.is_contiguous() and v.device == device):
        raise ValueError("%s must be a contiguous float32 CUDA vector of %d entries on %s" % (name, C, device))
    return v


def space_to_depth(x):
    """reference inf/layers/squeeze.py:5-13 (used by the unfused path and the tests)"""
    B, C, H, W = x.shape
    return x.view(B, C, H // 2, 2, W // 2, 2).permute(0, 1, 3, 5, 2, 4).contiguous().view(B, C * 4, H // 2, W // 2)


def depth_to_space(x):
    """reference inf/layers/squeeze.py:16-24"""
    B, C, H, W = x.shape
    return x.view(B, C // 4, 2, 2, H, W).permute(0, 1, 4, 2, 5, 3).contiguous().view(B, C // 4, H * 2, W * 2)


def inverse_fused(x, weight, in_scale=None, in_bias=None, squeeze=False, groups=None, out=None, prepared=None, orient=0):
    """y = L^-1( in_scale (.) S(x) + in_bias ): the ActNorm affine (reference inf/layers/actnorm.py:36) and, with
    `squeeze`, the Squeeze re-indexing S = space_to_depth (inf/layers/squeeze.py:5-13; x is then the
    (B, C/4, 2H, 2W) tensor) applied while the image enters shared memory -- one launch (ifk_inverse_fused_f32).
    Geometries the pipelined wavefront kernel does not serve run the same composition unfused."""
    lib = _native.load()
    if prepared is None:
        prepared = Prepared(weight, groups)
    else:
        prepared.check_weight(weight)
    _check_activation(x, "input")
    C = prepared.weight_shape[0]
    B = x.shape[0]
    if squeeze:
        if x.shape[1] * 4 != C or x.shape[2] % 2 or x.shape[3] % 2:
            raise ValueError("squeeze: input %s does not squeeze to %d channels" % (tuple(x.shape), C))
        H, W = x.shape[2] // 2, x.shape[3] // 2
    else:
        H, W = x.shape[2], x.shape[3]
    in_scale = _check_vec(in_scale, C, x.device, "in_scale")
    in_bias = _check_vec(in_bias, C, x.device, "in_bias")
    KH, KW = prepared.weight_shape[2:]
    p = _native.problem(B, C, H, W, KH, KW, prepared.weight_shape[1], prepared.groups, orient)
    if out is None:
        out = torch.empty(B, C, H, W, dtype=torch.float32, device=x.device)
    f = _native.Fused(in_scale.data_ptr() if in_scale is not None else None,
                      in_bias.data_ptr() if in_bias is not None else None, None, 1 if squeeze else 0)
    with torch.cuda.device(x.device):
        status = lib.ifk_inverse_fused_f32(ctypes.byref(p), ctypes.byref(f), _ptr(x), _ptr(prepared.buffer), _ptr(out),
                                           _native.current_stream(x.device))
        if status == _native.ERR_UNSUPPORTED or (squeeze and status == -2):
            xs = space_to_depth(x) if squeeze else x                  # the unfused sequence, same kernels
            if in_scale is not None:
                xs = xs * in_scale.view(1, C, 1, 1)
            if in_bias is not None:
                xs = xs + in_bias.view(1, C, 1, 1)
            return inverse(xs.contiguous(), weight, out=out, prepared=prepared, orient=orient)
        _native.check(status)
    return out


def bwd_input_fused(grad, weight, out_scale=None, squeeze=False, groups=None, prepared=None, orient=0, want_dx=True):
    """(dx, dz): dx = L^-T grad (raw: what bwd_weight reads) and dz = S^T( out_scale (.) dx ), the gradient handed to
    the layer in front of the fused ActNorm / Squeeze -- one launch (ifk_bwd_input_fused_f32)."""
    lib = _native.load()
    if prepared is None:
        prepared = Prepared(weight, groups)
    else:
        prepared.check_weight(weight)
    _check_activation(grad, "grad_output")
    B, C, H, W = grad.shape
    out_scale = _check_vec(out_scale, C, grad.device, "out_scale")
    p = prepared.for_batch(grad, orient)
    dx = torch.empty_like(grad) if want_dx else None
    dz = torch.empty((B, C // 4, 2 * H, 2 * W) if squeeze else (B, C, H, W), dtype=torch.float32, device=grad.device)
    f = _native.Fused(None, None, out_scale.data_ptr() if out_scale is not None else None, 1 if squeeze else 0)
    with torch.cuda.device(grad.device):
        status = lib.ifk_bwd_input_fused_f32(ctypes.byref(p), ctypes.byref(f), _ptr(grad), _ptr(prepared.buffer),
                                             _ptr(dx) if dx is not None else None, _ptr(dz),
                                             _native.current_stream(grad.device))
        if status == _native.ERR_UNSUPPORTED or (squeeze and status == -2):
            dx = bwd_input(grad, weight, prepared=prepared, orient=orient)
            t = dx * out_scale.view(1, C, 1, 1) if out_scale is not None else dx
            return dx, (depth_to_space(t) if squeeze else t.contiguous())
        _native.check(status)
    return dx, dz
