"""ActNorm (+ Squeeze) fused into the inverse-convolution layer that follows it.

In the if_* Glow models every flow step is  ActNorm -> inv_flow_with_pad -> ... -> Coupling  and every block opens
with a Squeeze (reference inf/experiments/if_glow_mnist.py:62-124).  The ActNorm affine
`(x - translation) * exp(-log_scale)` (inf/layers/actnorm.py:36) and the space_to_depth re-indexing
(inf/layers/squeeze.py:5-13) are applied by the solve kernel while the image moves between global and shared
memory (ifk_inverse_fused_f32 / ifk_bwd_input_fused_f32): two elementwise kernels and their HBM round trips per
layer disappear from the forward pass, two from the backward.  Parameters, initialisation and log-det are the
reference ActNorm's; the inverse-conv weight is an `inv_flow_with_pad`'s."""
import torch
import torch.autograd as autograd
import torch.nn as nn

from .. import functional as IF
from .flowlayer import FlowLayer
from .inv_conv import inv_flow_with_pad


class actnorm_inv_conv_(autograd.Function):
    """y = L^-1( (S(x) - t) * exp(-ls) ) in one launch; backward: one fused adjoint solve + dW + the ActNorm sums"""

    @staticmethod
    def forward(ctx, x, translation, log_scale, W, groups, orient, squeeze):
        x = x.contiguous()
        Wc = W.contiguous()
        s = torch.exp(-log_scale).contiguous()
        b = (-translation * s).contiguous()
        prepared = IF.Prepared(Wc, groups)
        y = IF.inverse_fused(x, Wc, in_scale=s, in_bias=b, squeeze=squeeze, prepared=prepared, orient=orient)
        ctx.save_for_backward(x, y, Wc, s, translation)
        ctx.prepared, ctx.orient, ctx.squeeze, ctx.groups = prepared, orient, squeeze, prepared.groups
        return y

    @staticmethod
    def backward(ctx, grad):
        x, y, W, s, t = ctx.saved_tensors
        dx, dz = IF.bwd_input_fused(grad.contiguous(), W, out_scale=s, squeeze=ctx.squeeze, prepared=ctx.prepared,
                                    orient=ctx.orient)
        dw = IF.bwd_weight(dx, y, W, groups=ctx.groups, orient=ctx.orient)
        # ActNorm parameters (out = (u - t) s with u = S(x)): d t = -s sum(dx), d ls = -s sum(dx (u - t))
        u = IF.space_to_depth(x) if ctx.squeeze else x
        sum_dx = dx.sum(dim=(0, 2, 3))
        dt = -s * sum_dx
        dls = -s * ((dx * u).sum(dim=(0, 2, 3)) - t * sum_dx)
        return dz, dt, dls, dw, None, None, None


class ActNormInvFlow(FlowLayer):
    """ActNorm(n_dims) followed by inv_flow_with_pad(n_dims, n_dims, kernel_size, order) as ONE layer, optionally
    absorbing the Squeeze in front of it (`squeeze=True`: the input is the un-squeezed (B, C/4, 2H, 2W) tensor).
    forward -> (out, logdet) with the ActNorm log-det (the inverse conv's is exactly 0)."""

    def __init__(self, n_dims, kernel_size=(3, 3), order="TL", squeeze=False, groups=None):
        super().__init__()
        self.n_dims, self.squeeze = n_dims, bool(squeeze)
        self.translation = nn.Parameter(torch.zeros(n_dims))
        self.log_scale = nn.Parameter(torch.zeros(n_dims))
        self.register_buffer("initialized", torch.tensor(0))
        self._init_checked = False      # host-side memo of `initialized`: no device read per call (graph capture)
        self.conv = inv_flow_with_pad(n_dims, n_dims, kernel_size, order=order, groups=groups)

    def _initialize(self, input):
        """data-dependent initialisation, as the reference ActNorm (actnorm.py:21-27), on the squeezed view"""
        with torch.no_grad():
            u = IF.space_to_depth(input) if self.squeeze else input
            self.translation.data.copy_(u.mean(dim=(0, 2, 3)))
            self.log_scale.data.copy_(torch.log(u.std(dim=(0, 2, 3)) + 1e-8))
            self.initialized.fill_(1)

    def forward(self, input, context=None, compute_expensive=False):
        if not self._init_checked:
            if not bool(self.initialized):
                self._initialize(input)
            self._init_checked = True
        out = actnorm_inv_conv_.apply(input, self.translation, self.log_scale, self.conv.weight_fwd, self.conv.groups,
                                      self.conv.order, self.squeeze)
        H, W = out.shape[2:]
        ldj = -self.log_scale.sum().expand(input.size(0)) * H * W
        return out, ldj

    def reverse(self, input, context=None, compute_expensive=False):
        assert self._init_checked or bool(self.initialized)
        u = self.conv.reverse(input)
        u = u * torch.exp(self.log_scale).view(1, -1, 1, 1) + self.translation.view(1, -1, 1, 1)
        return IF.depth_to_space(u) if self.squeeze else u

    def logdet(self, input, context=None, compute_expensive=False):
        H, W = input.shape[2:]
        if self.squeeze:
            H, W = H // 2, W // 2
        return -self.log_scale.sum().expand(input.size(0)) * H * W
