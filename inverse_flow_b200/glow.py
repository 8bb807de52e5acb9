"""A compact if_glow model around the inverse-convolution layers, for end-to-end training runs.

SURVEY.md section 8f rank 1: the reference trains Glow-style flows whose 1x1 convolutions are
replaced by `inv_flow_*` layers (inf/experiments/if_glow_mnist.py:33-132,
inf/if_multiGPU_imagenet32.py:174-250): per block  Squeeze -> K x [ActNorm -> inv_flow ->
Coupling]  with a split prior between blocks and a Gaussian base density.  Only the
inverse-conv layer is this repository's product; the layers around it are deliberately minimal
PyTorch (they are out of scope) and exist so that the hot path can
be timed inside a real forward/backward/optimizer step and under DistributedDataParallel.

Every layer follows the reference's FlowLayer contract (inf/layers/flowlayer.py:7-51):
forward(x) -> (y, log|det J| per sample or 0.0), reverse(y) -> x.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .layers import ActNormInvFlow, inv_flow_no_pad


class Squeeze(nn.Module):
    """space-to-depth by 2 (reference inf/layers/squeeze.py:5-40 computes the same map)."""

    def forward(self, x, context=None):
        B, C, H, W = x.shape
        x = x.view(B, C, H // 2, 2, W // 2, 2).permute(0, 1, 3, 5, 2, 4)
        return x.reshape(B, C * 4, H // 2, W // 2), 0.0

    def reverse(self, y, context=None):
        B, C, H, W = y.shape
        y = y.view(B, C // 4, 2, 2, H, W).permute(0, 1, 4, 2, 5, 3)
        return y.reshape(B, C // 4, H * 2, W * 2)


class ActNorm(nn.Module):
    """per-channel affine y = (x + b) * exp(s), data-dependent initialisation on request (the
    inverse convolutions amplify: with the reference's dirac-centred 3x3 init the solve is a
    running alternating sum along the diagonal, so activations must be re-normalised per step)."""

    def __init__(self, channels):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(1, channels, 1, 1))
        self.log_scale = nn.Parameter(torch.zeros(1, channels, 1, 1))
        self.init_from_next_batch = False

    def forward(self, x, context=None):
        if self.init_from_next_batch:
            with torch.no_grad():
                mean = x.mean(dim=(0, 2, 3), keepdim=True)
                std = x.std(dim=(0, 2, 3), keepdim=True) + 1e-6
                self.bias.copy_(-mean)
                self.log_scale.copy_(-torch.log(std))
            self.init_from_next_batch = False
        return (x + self.bias) * torch.exp(self.log_scale), self.log_scale.sum() * x.shape[2] * x.shape[3]

    def reverse(self, y, context=None):
        return y * torch.exp(-self.log_scale) - self.bias


class Coupling(nn.Module):
    """affine coupling: the first half of the channels parametrises scale/shift of the second."""

    def __init__(self, channels, width):
        super().__init__()
        self.c1 = channels // 2
        c2 = channels - self.c1
        self.net = nn.Sequential(
            nn.Conv2d(self.c1, width, 3, padding=1), nn.ReLU(),
            nn.Conv2d(width, width, 1), nn.ReLU(),
            nn.Conv2d(width, 2 * c2, 3, padding=1))
        nn.init.zeros_(self.net[-1].weight)
        nn.init.zeros_(self.net[-1].bias)

    def _params(self, x1):
        shift, raw = self.net(x1).chunk(2, dim=1)
        return shift, F.logsigmoid(raw + 2.0)

    def forward(self, x, context=None):
        x1, x2 = x[:, :self.c1], x[:, self.c1:]
        shift, log_s = self._params(x1)
        y2 = (x2 + shift) * torch.exp(log_s)
        return torch.cat([x1, y2], dim=1), log_s.flatten(1).sum(1)

    def reverse(self, y, context=None):
        y1, y2 = y[:, :self.c1], y[:, self.c1:]
        shift, log_s = self._params(y1)
        return torch.cat([y1, y2 * torch.exp(-log_s) - shift], dim=1)


def gaussian_logprob(z):
    return (-0.5 * (z ** 2) - 0.5 * math.log(2 * math.pi)).flatten(1).sum(1)


class IFGlow(nn.Module):
    """L blocks of K flow steps; after each block but the last, half of the channels is factored
    out under a standard normal (split prior).  forward(x) -> (list of latents, log p(x))."""

    def __init__(self, shape=(1, 28, 28), num_blocks=2, block_size=16, kernel_size=2,
                 coupling_width=128, groups=None, actnorm=True, fused=False):
        super().__init__()
        C, H, W = shape
        self.shape = tuple(shape)
        self.blocks = nn.ModuleList()
        self.inv_layers = []
        self.fused = bool(fused and actnorm)     # ActNorm (and the block's Squeeze) inside the inverse-conv kernels
        for level in range(num_blocks):
            C, H, W = C * 4, H // 2, W // 2
            steps = nn.ModuleList()
            for k_step in range(block_size):
                step = nn.ModuleList()
                if self.fused:
                    layer = ActNormInvFlow(C, (kernel_size, kernel_size), order="TL", squeeze=(k_step == 0), groups=groups)
                    self.inv_layers.append(layer.conv)
                    step.append(layer)
                    step.append(Coupling(C, coupling_width))
                    steps.append(step)
                    continue
                if actnorm:
                    step.append(ActNorm(C))
                inv = inv_flow_no_pad(C, C, (kernel_size, kernel_size), groups=groups)
                self.inv_layers.append(inv)
                step.append(inv)
                step.append(Coupling(C, coupling_width))
                steps.append(step)
            self.blocks.append(steps)
            if level < num_blocks - 1:
                C = C // 2
        self.num_blocks = num_blocks

    def forward(self, x):
        logdet = torch.zeros(x.shape[0], device=x.device)
        latents = []
        for level, steps in enumerate(self.blocks):
            if not self.fused:                   # (fused: the first step of the block reads the un-squeezed tensor)
                x, _ = Squeeze()(x)
            for step in steps:
                for layer in step:
                    x, ld = layer(x)
                    logdet = logdet + ld
            if level < self.num_blocks - 1:
                keep = x.shape[1] // 2
                latents.append(x[:, keep:])
                x = x[:, :keep].contiguous()
        latents.append(x)
        logp = logdet + sum(gaussian_logprob(z) for z in latents)
        return latents, logp

    @torch.no_grad()
    def reverse(self, latents):
        x = latents[-1]
        for level in reversed(range(self.num_blocks)):
            if level < self.num_blocks - 1:
                x = torch.cat([x, latents[level]], dim=1)
            for step in reversed(self.blocks[level]):
                for layer in reversed(step):
                    x = layer.reverse(x)
            if not self.fused:
                x = Squeeze().reverse(x)
        return x

    @torch.no_grad()
    def initialize(self, x):
        """data-dependent ActNorm initialisation from one batch (call once, before wrapping in
        DistributedDataParallel, with the SAME batch on every rank)."""
        for m in self.modules():
            if isinstance(m, ActNorm):
                m.init_from_next_batch = True
        self.forward(x)

    def loss(self, x):
        """negative log-likelihood in bits per dimension (without the dequantisation constant)."""
        _, logp = self.forward(x)
        return -logp.mean() / (math.log(2.0) * x[0].numel())


CONFIGS = {
    # name: (input shape, L, K, kernel, per-GPU batch)   -- reference experiment files cited in bench.py
    "if_glow_mnist": ((1, 28, 28), 2, 16, 2, 100),
    "if_glow_cifar": ((3, 32, 32), 2, 16, 3, 256),
    "if_glow_imagenet32": ((3, 32, 32), 3, 48, 3, 100),
}


def build(name, coupling_width=128, groups=None, fused=False):
    shape, L, K, k, batch = CONFIGS[name]
    return IFGlow(shape, L, K, k, coupling_width, groups, fused=fused), shape, batch
