"""FInC-Flow layers whose reverse pass runs on the same inverse kernels.

Mirrors inf/layers/conv.py:22-222 (`PaddedConv2d`) and inf/layers/fincflow.py:14-101
(`Finc_FlowUnit`): same constructor signatures, the same `conv.weight` parameter (state_dict
key), the same initialisation / mask, forward = zero-padded nn.Conv2d in the given corner order.
What changes is `reverse`: the reference ships the GPU tensor to the host, runs the Cython solver
in float64 and copies back (conv.py:110-166), or JIT-loads a second CUDA extension with one
launch + device sync per diagonal (cinc_kernel_level1/2.cu); here it is one asynchronous call of
`ifk_inverse_f32` -- groups = 1 for a single PaddedConv2d ("level 1"), groups = 4 for the packed
four-orientation unit ("level 2", fincflow.py:79-101).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import functional as IF
from .flowlayer import FlowLayer

_FLIP = {"TL": None, "TR": [3], "BL": [2], "BR": [2, 3]}


def _flip(t, order):
    dims = _FLIP[order]
    return t if dims is None else torch.flip(t, dims)


class PaddedConv2d(FlowLayer):
    """Conv2d zero-padded towards one corner; `order` in {TL, TR, BL, BR} (conv.py:22-61)."""

    def __init__(self, in_channels, out_channels, kernel_size, bias=False, order='TL'):
        super().__init__()
        assert len(kernel_size) == 2
        assert order in _FLIP, 'unknown order: {}'.format(order)
        self.kernel_size = kernel_size
        self.order = order
        K_H, K_W = kernel_size
        self.pad = {'TL': (K_W - 1, 0, K_H - 1, 0), 'TR': (0, K_W - 1, K_H - 1, 0),
                    'BL': (K_W - 1, 0, 0, K_H - 1), 'BR': (0, K_W - 1, 0, K_H - 1)}[order]
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, bias=bias)
        self.reset_parameters()

    def reset_parameters(self):
        """N(0, 0.05) taps, unit diagonal, zero centre-tap upper triangle, then the stored weight
        is flipped into the layer's orientation (conv.py:63-79)."""
        w = self.conv.weight.data
        nn.init.normal_(w, mean=0.0, std=0.05)
        if self.conv.bias is not None:
            nn.init.constant_(self.conv.bias, 0)
        for c_out in range(w.shape[0]):
            w[c_out, c_out, -1, -1] = 1.0
            w[c_out, c_out + 1:, -1, -1] = 0.0
        self.conv.weight.data = _flip(w, self.order).contiguous()
        self.mask = self.get_mask()

    def get_mask(self):
        mask = torch.ones_like(self.conv.weight.data)
        for c_out in range(mask.shape[0]):
            mask[c_out, c_out:, -1, -1] = 0.0
        return _flip(mask, self.order)

    def reset_gradients(self):
        if self.conv.weight.grad is not None:
            self.conv.weight.grad = self.conv.weight.grad * self.mask.to(self.conv.weight.grad.device)

    def tl_weight(self):
        """the kernel in top-left orientation, the form every solver takes (conv.py:118-160)."""
        return _flip(self.conv.weight.data, self.order).contiguous()

    def forward(self, x, context=None, compute_expensive=None):
        return self.conv(F.pad(x, self.pad)), 0.0

    def reverse(self, x, context=None, compute_expensive=None):
        """-> (y, 0.0) like the reference's reverse_cython / reverse_cuda (conv.py:166, 219)."""
        if self.conv.bias is not None:
            x = x - self.conv.bias.reshape(1, -1, 1, 1)
        # the image is not flipped: the kernel walks it from the layer's corner (ifk.h IFK_ORIENT_*)
        return IF.inverse(x.contiguous(), self.tl_weight(), groups=1, orient=self.order), 0.0

    def logdet(self, x, context=None):
        return 0.0


class Finc_FlowUnit(nn.Module):
    """four PaddedConv2d, one per corner order, each on a quarter of the channels
    (fincflow.py:14-50); reverse packs them into ONE 4-group solve (fincflow.py:79-101)."""

    def __init__(self, in_channels, out_channels, kernel_size):
        super().__init__()
        if isinstance(kernel_size, int) or len(kernel_size) == 1:
            kernel_size = (kernel_size, kernel_size) if isinstance(kernel_size, int) else tuple(kernel_size) * 2
        assert in_channels % 4 == 0, "Input channels have to be a multiple of 4"
        q = in_channels // 4
        self.conv_tl = PaddedConv2d(q, q, kernel_size, order='TL')
        self.conv_tr = PaddedConv2d(q, q, kernel_size, order='TR')
        self.conv_bl = PaddedConv2d(q, q, kernel_size, order='BL')
        self.conv_br = PaddedConv2d(q, q, kernel_size, order='BR')

    def _convs(self):
        return (self.conv_tl, self.conv_tr, self.conv_bl, self.conv_br)

    def forward(self, x, context=None):
        outs, logdet = [], 0.0
        for conv, xi in zip(self._convs(), torch.chunk(x, 4, dim=1)):
            o, ld = conv.forward(xi)
            outs.append(o)
            logdet += ld
        return torch.cat(outs, dim=1), logdet

    def reverse(self, x, context=None):
        return self.reverse_level2(x)

    def reverse_level1(self, x):
        outs = [conv.reverse(xi)[0] for conv, xi in zip(self._convs(), torch.chunk(x, 4, dim=1))]
        return torch.cat(outs, dim=1)

    def reverse_level2(self, x):
        kernel = torch.cat([c.tl_weight() for c in self._convs()], dim=0)          # (C, C/4, k, k)
        chunks = [_flip(xi, c.order) for c, xi in zip(self._convs(), torch.chunk(x, 4, dim=1))]
        y = IF.inverse(torch.cat(chunks, dim=1).contiguous(), kernel.contiguous(), groups=4)
        outs = [_flip(yi, c.order) for c, yi in zip(self._convs(), torch.chunk(y, 4, dim=1))]
        return torch.cat(outs, dim=1)
