"""Drop-in for the reference's `inv_conv_with_bp` extension module.

Same four names and argument order as the pybind module exported at
inf/utils/inv_conv_cuda/inv_conv_with_bp_general.cpp:115-120, so the reference's layer file
(inf/layers/inv_conv.py:21,52,74,79,266,459) runs unchanged with

    import inverse_flow_b200.inv_conv_with_bp as inv_conv_with_bp

Like the reference, each function writes into the caller-allocated `output` and returns
`[output]`; the scratch tensors `M` the reference needs are accepted and ignored.
Differences by design: `dy` returns the true input gradient L^-T g and
`dw` the true weight gradient (the reference's `dw` is fed the layer INPUT; here the third
positional tensor must be the saved OUTPUT y of `inverse`, see `inv_conv_` in layers/);
work is enqueued on the current stream without device synchronisation.
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


def dw(saved_output, kernel, grad_input, M=None, output=None, groups=None):
    """reference: dw(input, kernel, grad_output, M, output) -> [output]   (.cpp:99-112).

    Here: `saved_output` = y of `inverse`, `grad_input` = dX returned by `dy`."""
    return [F.bwd_weight(grad_input, saved_output, kernel, groups=groups, out=output)]


def backward(grad_output, saved_output, kernel, groups=None):
    """fused (dX, dW); no reference counterpart (it calls dy then dw, inv_conv.py:74-79)."""
    return list(F.backward(grad_output, saved_output, kernel, groups=groups))
