"""Inv_FlowUnit: the four-orientation composite TL -> TR -> BL -> BR of the reference
(inf/layers/inv_flow.py:13-53)."""
import torch.nn as nn

from .inv_conv import inv_flow_with_pad


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

    def forward(self, x, context=None):
        logdet_accum = 0.0
        for conv in (self.conv_tl, self.conv_tr, self.conv_bl, self.conv_br):
            x, logdet = conv.forward(x)
            logdet_accum += logdet
        return x, logdet_accum

    def reverse(self, x, context=None):
        for conv in (self.conv_br, self.conv_bl, self.conv_tr, self.conv_tl):
            x = conv.reverse(x)
        return x

    def logdet(self, input, context=None):
        return 0.0
