"""Inv_FlowUnit: the four-orientation composite TL -> TR -> BL -> BR of the reference
(inf/layers/inv_flow.py:13-53).

The four layers feed each other directly, so the forward pass is ONE launch (functional.inverse_chain ->
ifk_inverse_chain_f32): the image stays in shared memory across the four solves and only the four outputs the
backward needs go to memory.  The backward walks the layers in reverse with the fused dX / dW call each."""
import torch.autograd as autograd
import torch.nn as nn

from .. import functional as IF
from .inv_conv import inv_flow_with_pad

ORDERS = ("TL", "TR", "BL", "BR")


class inv_flow_unit_(autograd.Function):
    """y = L_BR^-1 L_BL^-1 L_TR^-1 L_TL^-1 x as one chained launch, with the per-layer parallel backward"""

    @staticmethod
    def forward(ctx, x, groups, *weights):
        x = x.contiguous()
        ws = [w.contiguous() for w in weights]
        prepared = [IF.Prepared(w, groups) for w in ws]
        ys = IF.inverse_chain(x, prepared, ORDERS[:len(ws)])
        ctx.save_for_backward(*ys, *ws)
        ctx.prepared = prepared
        return ys[-1]

    @staticmethod
    def backward(ctx, grad):
        n = len(ctx.prepared)
        ys, ws = ctx.saved_tensors[:n], ctx.saved_tensors[n:]
        g = grad.contiguous()
        dws = [None] * n
        for i in reversed(range(n)):
            g, dws[i] = IF.backward(g, ys[i], ws[i], prepared=ctx.prepared[i], orient=ORDERS[i])
        return (g, None, *dws)


class Inv_FlowUnit(nn.Module):

    def __init__(self, in_channels, out_channels, kernel_size, groups=None):
        super().__init__()
        if isinstance(kernel_size, int) or len(kernel_size) == 1:
            kernel_size = (kernel_size, kernel_size) if isinstance(kernel_size, int) else tuple(kernel_size) * 2
        assert in_channels % 4 == 0, "Input channels have to be a multiple of 4"
        self.conv_tl = inv_flow_with_pad(out_channels, out_channels, kernel_size, order='TL', groups=groups)
        self.conv_tr = inv_flow_with_pad(out_channels, out_channels, kernel_size, order='TR', groups=groups)
        self.conv_bl = inv_flow_with_pad(out_channels, out_channels, kernel_size, order='BL', groups=groups)
        self.conv_br = inv_flow_with_pad(out_channels, out_channels, kernel_size, order='BR', groups=groups)

    def _convs(self):
        return (self.conv_tl, self.conv_tr, self.conv_bl, self.conv_br)

    def forward(self, x, context=None):
        convs = self._convs()
        y = inv_flow_unit_.apply(x, convs[0].groups, *[c.weight_fwd for c in convs])
        for c in convs:
            if c.training:
                c.logabsdet_dirty = True
        return y, 0.0               # four unit-diagonal operators: log|det| = 0 (reference: sum of four 0.0)

    def reverse(self, x, context=None):
        for conv in (self.conv_br, self.conv_bl, self.conv_tr, self.conv_tl):
            x = conv.reverse(x)
        return x

    def logdet(self, input, context=None):
        return 0.0
