"""Inverse-convolution autograd binding and FlowLayers, mirroring inf/layers/inv_conv.py.

Same class / function names, constructor signatures, parameter name (`weight_fwd`), return
conventions ((out, 0.0) from forward, tensor from reverse) and helper methods as the
reference (inv_conv.py:43-91 `inv_conv_`, `inv_conv_4d`; :94-364 `inv_flow_with_pad`;
:365-513 `inv_flow_no_pad`), so `create_model` of the if_* experiments works unchanged and
reference checkpoints load (state_dict key `weight_fwd`).  Loaded weights mean the same operator
for `inv_flow_no_pad` and order 'TL' -- the only variants the if_* models use; a reference checkpoint of
a TR/BL/BR layer stores the weight in whatever flip state its last forward left it (the reference
toggles `weight_fwd.data` on every call, inv_conv.py:198-214), while here `weight_fwd` is always the
top-left kernel of the reflected frame.  The compute goes through the C ABI
(inverse_flow_b200.functional); there is no PyTorch fallback.

Deliberate differences, all documented in DESIGN.md:
  * backward returns the true gradients (dX = L^-T g, dW = -corr(dX, y)), SURVEY.md 0.4;
  * the saved tensors are (y, W), not (x, W, y); no CPU-side scratch allocations
    (reference inv_conv.py:70-77 builds a (B,C,k,k,H,W) tensor on the host every call);
  * orders TR/BL/BR act as F L^-1 F with F the reflection of the order's axes -- what flipping
    the data around a top-left kernel computes -- but as index reflections inside the kernels
    (`orient`, ifk.h), without flip copies; `weight_fwd` is left untouched and `reverse` uses the
    same orientation, so every order round-trips (the reference re-flips the stored weight on
    every call and never flips in reverse, inv_conv.py:198-214, 249-267);
  * `groups`: None -> the reference's 4 channel groups when C % 4 == 0, else 1.
"""
import numpy as np
import torch
import torch.autograd as autograd
import torch.nn as nn
from torch.nn.modules.utils import _pair

from .. import functional as IF
from .flowlayer import FlowLayer, mark_expensive

class inv_conv_(autograd.Function):
    """y = L^-1 x with its parallel backward (reference inv_conv.py:43-86)."""

    @staticmethod
    def forward(ctx, x, W, groups=None, orient=0):
        x = x.contiguous()
        Wc = W.contiguous()
        prepared = IF.Prepared(Wc, groups)
        y = IF.inverse(x, Wc, prepared=prepared, orient=orient)
        ctx.save_for_backward(y, Wc)
        ctx.prepared = prepared
        ctx.orient = orient
        return y

    @staticmethod
    def backward(ctx, output_grad):
        y, W = ctx.saved_tensors
        dx, dw = IF.backward(output_grad.contiguous(), y, W, prepared=ctx.prepared, orient=ctx.orient)
        return dx, dw, None, None


def inv_conv_4d(x, W, groups=None, orient=0):
    return inv_conv_.apply(x, W, groups, orient)


class _InvFlowBase(FlowLayer):

    def _init_common(self, in_channels, out_channels, kernel_size, sym_recon_grad, only_R_recon,
                     recon_loss_weight, recon_loss_lr, recon_alpha, groups):
        assert len(kernel_size) == 2
        if in_channels != out_channels:
            raise ValueError("an invertible convolution needs in_channels == out_channels")
        self.kernel_size = _pair(kernel_size)
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.groups = groups
        self.sym_recon_grad = sym_recon_grad
        self.only_R_recon = only_R_recon
        self.recon_loss_weight = recon_loss_weight
        self.recon_loss_lr = recon_loss_lr
        self.recon_loss_ema = None
        self.alpha = recon_alpha

    def reset_parameters(self):
        """dirac + xavier_normal(gain=0.01), then W[c_out, -1, -1, -1] = 1
        (reference inv_conv.py:153-170, 399-416)."""
        self.logabsdet_dirty = True
        w_shape = (self.out_channels, self.in_channels, *self.kernel_size)
        w_eye = nn.init.dirac_(torch.empty(w_shape))
        w_noise = nn.init.xavier_normal_(torch.empty(w_shape), gain=0.01)
        if self.kernel_size[0] == 1 and self.kernel_size[1] == 1:
            w_np = np.random.randn(self.out_channels, self.in_channels)
            w_init = torch.tensor(np.linalg.qr(w_np)[0]).to(torch.float).view(w_shape)
        else:
            w_init = w_eye + w_noise
        self.weight_fwd = nn.Parameter(w_init)
        with torch.no_grad():
            self.weight_fwd[:, -1, -1, -1] = 1.0
        self.mask = self.get_mask()

    def get_mask(self):
        """1 where a weight entry is trainable: the centre tap's diagonal and upper triangle are
        fixed (reference inv_conv.py:233-248)."""
        mask = torch.ones_like(self.weight_fwd.data)
        for c_out in range(mask.shape[0]):
            mask[c_out, c_out:, -1, -1] = 0.0
        return mask

    def reset_gradients(self):
        """zero the gradient of the masked entries (reference inv_conv.py:223-231)."""
        if self.weight_fwd.grad is not None:
            self.mask = self.get_mask()
            self.weight_fwd.grad = self.weight_fwd.grad * self.mask.to(self.weight_fwd.grad.device)

    def forward(self, input, context=None, compute_expensive=False):
        if self.training:
            self.logabsdet_dirty = True
        self.input = input
        self.output = inv_conv_4d(input, self.weight_fwd, self.groups, getattr(self, "order", "TL"))
        return self.output, 0.0

    def reverse(self, input, context=None, compute_expensive=False):
        return IF.conv(input.contiguous(), self.weight_fwd.detach().contiguous(), groups=self.groups,
                       orient=getattr(self, "order", "TL"))

    @mark_expensive
    def logdet(self, input, context=None, compute_expensive=False):
        """The operator has a unit diagonal: log|det| = 0 exactly (reference returns 0.0 on the
        path FlowSequential uses, inv_conv.py:331-332, 480-481)."""
        if compute_expensive:
            return torch.zeros(len(input), device=input.device)
        return 0.0


class inv_flow_with_pad(_InvFlowBase):
    def __init__(self, in_channels, out_channels, kernel_size, order='TL', sym_recon_grad=False,
                 only_R_recon=False, recon_loss_weight=1.0, recon_loss_lr=0.0, recon_alpha=0.9,
                 groups=None):
        super().__init__()
        assert order in {'TL', 'TR', 'BL', 'BR'}, 'unknown order: {}'.format(order)
        self.order = order
        self._init_common(in_channels, out_channels, kernel_size, sym_recon_grad, only_R_recon,
                          recon_loss_weight, recon_loss_lr, recon_alpha, groups)
        K_H, K_W = self.kernel_size
        self.pad = {'TL': (K_W - 1, 0, K_H - 1, 0), 'TR': (0, K_W - 1, K_H - 1, 0),
                    'BL': (K_W - 1, 0, 0, K_H - 1), 'BR': (0, K_W - 1, 0, K_H - 1)}[order]
        self.reset_parameters()


class inv_flow_no_pad(_InvFlowBase):
    def __init__(self, in_channels, out_channels, kernel_size, sym_recon_grad=False,
                 only_R_recon=False, recon_loss_weight=1.0, recon_loss_lr=0.0, recon_alpha=0.9,
                 groups=None):
        super().__init__()
        self._init_common(in_channels, out_channels, kernel_size, sym_recon_grad, only_R_recon,
                          recon_loss_weight, recon_loss_lr, recon_alpha, groups)
        self.reset_parameters()
