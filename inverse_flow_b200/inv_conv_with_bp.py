"""Drop-in for the reference's `inv_conv_with_bp` extension module.

Same four names as the pybind module exported at
inf/utils/inv_conv_cuda/inv_conv_with_bp_general.cpp:115-120:

    import inverse_flow_b200.inv_conv_with_bp as inv_conv_with_bp

`inverse`, `forward` and `dy` take the reference's positional arguments unchanged; each function writes
into the caller-allocated `output` and returns `[output]`, and the scratch tensors `M` the reference
needs are accepted and ignored.  Work is enqueued on the current stream without device synchronisation.

NOT a drop-in for the reference's `inv_conv_.backward` as written (inf/layers/inv_conv.py:62-81): that code
calls `dw(input_x, kernel, output_grad, M, out)`, i.e. it feeds the layer INPUT and the UPSTREAM gradient,
which is not what the weight gradient is a function of.  The true gradient is
dW = -corr(dX, y) with y the saved OUTPUT of `inverse` and dX the result of `dy`, so `dw` here takes those
two tensors and takes them BY KEYWORD ONLY: a reference-style positional call raises a TypeError instead
of silently computing something else.  `layers/inv_conv.py` holds the backward that goes with it.
`dy` returns the true input gradient L^-T g (the reference's dy computes L^-1 g).
"""
from . import functional as F

__all__ = ["inverse", "forward", "dy", "dw", "backward"]


def inverse(input, kernel, output=None, groups=None):
    """reference: inverse(input, kernel, output) -> [output]   (.cpp:19-28)"""
    return [F.inverse(input, kernel, groups=groups, out=output)]


def forward(input, kernel, output=None, groups=None):
    """reference: forward(input, kernel, output) -> [output]   (.cpp:44-53)"""
    return [F.conv(input, kernel, groups=groups, out=output)]


def dy(grad_output, kernel, M=None, output=None, groups=None):
    """reference: dy(grad_output, kernel, M, output) -> [output]   (.cpp:70-81)"""
    return [F.bwd_input(grad_output, kernel, groups=groups, out=output)]


def dw(*, saved_output, kernel, grad_input, M=None, output=None, groups=None):
    """reference: dw(input, kernel, grad_output, M, output) -> [output]   (.cpp:99-112).

    Keyword-only on purpose (see the module docstring): `saved_output` = y returned by `inverse`,
    `grad_input` = dX returned by `dy`."""
    return [F.bwd_weight(grad_input, saved_output, kernel, groups=groups, out=output)]


def backward(grad_output, saved_output, kernel, groups=None):
    """fused (dX, dW); no reference counterpart (it calls dy then dw, inv_conv.py:74-79)."""
    return list(F.backward(grad_output, saved_output, kernel, groups=groups))
