from .flowlayer import FlowLayer
from .inv_conv import inv_conv_, inv_conv_4d, inv_flow_no_pad, inv_flow_with_pad
from .inv_flow import Inv_FlowUnit
from .fincflow import Finc_FlowUnit, PaddedConv2d
from .fused import ActNormInvFlow

__all__ = ["FlowLayer", "inv_conv_", "inv_conv_4d", "inv_flow_no_pad", "inv_flow_with_pad", "Inv_FlowUnit", "Finc_FlowUnit", "PaddedConv2d", "ActNormInvFlow"]
