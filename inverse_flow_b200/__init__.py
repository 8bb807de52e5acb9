"""inverse_flow_b200: B200-native (sm_100a) kernels for Inverse-Flow's inverse-convolution
hot path, behind a C ABI (include/ifk.h), with the reference's torch.autograd.Function /
FlowLayer surface on top.  See DESIGN.md."""
from . import functional
from .functional import default_groups

__all__ = ["functional", "default_groups"]
__version__ = "0.2.0"
