"""InvConvStack: the inverse-convolution layers of an if_* Glow model as one static pipeline.

The if_glow models interleave `inv_flow_*` layers with ActNorm / Spline / Coupling layers
(reference experiments/if_glow_mnist.py:62-124).  This class holds ONLY the inverse-conv
layers of such a model -- per stage `n_layers` layers of one (C, H, W, k) shape, chained --
with every buffer preallocated, all launches going straight through the C ABI, and the whole
forward + backward captured in one CUDA graph (B200-first: streams and graphs instead of a
per-layer Python loop).  It is what bench.py times and what a training driver would replay.

Forward per layer  : prepare(W) ; y = L^-1 x                (x of layer i+1 = y of layer i)
Backward per layer : dX = L^-T g ; dW = -corr(dX, y)        (g of layer i-1 = dX of layer i)
dW of every layer is written into one flat bucket (`grad_bucket`), i.e. directly into the
buffer a data-parallel all-reduce consumes.
"""
import ctypes

import torch

from . import _native
from .functional import default_groups


def reference_init_weight(C, k, generator=None):
    """inv_flow_*.reset_parameters (reference inf/layers/inv_conv.py:153-170)."""
    w = torch.nn.init.dirac_(torch.empty(C, C, k, k))
    w = w + torch.nn.init.xavier_normal_(torch.empty(C, C, k, k), gain=0.01, generator=generator)
    w[:, -1, -1, -1] = 1.0
    return w


class _Stage:
    pass


class InvConvStack:
    def __init__(self, stages, batch, groups=None, device="cuda", seed=0):
        """stages: iterable of (C, H, W, k, n_layers)."""
        self.lib = _native.load()
        self.device = torch.device(device)
        self.batch = int(batch)
        gen = torch.Generator().manual_seed(seed)
        self.stages = []
        n_w = sum(C * C * k * k * n for (C, H, W, k, n) in stages)
        self.weights = torch.empty(n_w, dtype=torch.float32, device=self.device)
        self.grad_bucket = torch.zeros(n_w, dtype=torch.float32, device=self.device)
        off = 0
        for (C, H, W, k, n) in stages:
            st = _Stage()
            st.C, st.H, st.W, st.k, st.n = C, H, W, k, n
            st.groups = default_groups(C) if groups is None else groups
            st.problem = _native.problem(self.batch, C, H, W, k, k, C, st.groups)
            pf = self.lib.ifk_prepared_floats(ctypes.byref(st.problem))
            ws = self.lib.ifk_bwd_weight_workspace_bytes(ctypes.byref(st.problem))
            if pf == 0:
                raise ValueError("unsupported stage %s" % ((C, H, W, k, n),))
            st.w, st.dw, st.prepared = [], [], []
            for _ in range(n):
                sz = C * C * k * k
                st.w.append(self.weights[off:off + sz].view(C, C, k, k))
                st.dw.append(self.grad_bucket[off:off + sz].view(C, C, k, k))
                st.w[-1].copy_(reference_init_weight(C, k, gen))
                st.prepared.append(torch.empty(pf, dtype=torch.float32, device=self.device))
                off += sz
            shape = (self.batch, C, H, W)
            st.act = [torch.zeros(shape, device=self.device) for _ in range(n + 1)]   # act[0] = x
            st.grad = [torch.zeros(shape, device=self.device) for _ in range(2)]     # ping-pong
            st.grad_in = torch.zeros(shape, device=self.device)                       # upstream g
            st.workspace = torch.empty((ws + 3) // 4, dtype=torch.float32, device=self.device)
            self.stages.append(st)
        self.graph = None
        self.launches_per_step = sum(st.n * 5 for st in self.stages)

    # -- raw launches ------------------------------------------------------------------
    def _stream(self):
        return _native.current_stream(self.device)

    def forward(self):
        lib, s = self.lib, self._stream()
        for st in self.stages:
            p = ctypes.byref(st.problem)
            for i in range(st.n):
                _native.check(lib.ifk_prepare_f32(p, st.w[i].data_ptr(), st.prepared[i].data_ptr(), s))
                _native.check(lib.ifk_inverse_f32(p, st.act[i].data_ptr(), st.prepared[i].data_ptr(),
                                                  st.act[i + 1].data_ptr(), s))

    def backward(self):
        lib, s = self.lib, self._stream()
        for st in self.stages:
            p = ctypes.byref(st.problem)
            g = st.grad_in
            for i in reversed(range(st.n)):
                dx = st.grad[i & 1]
                _native.check(lib.ifk_backward_f32(p, g.data_ptr(), st.act[i + 1].data_ptr(),
                                                   st.prepared[i].data_ptr(), dx.data_ptr(),
                                                   st.dw[i].data_ptr(), st.workspace.data_ptr(), s))
                g = dx
            st.dx = g

    def forward_backward(self):
        self.forward()
        self.backward()

    # -- graph -------------------------------------------------------------------------
    def capture(self):
        with torch.cuda.device(self.device):
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self.forward_backward()           # warm-up: sets function attributes, loads modules
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.forward_backward()
        return self

    def step(self):
        """one forward+backward over the resident batch (device buffers)."""
        if self.graph is None:
            self.capture()
        self.graph.replay()

    # -- host-buffer path (the end-to-end call) -----------------------------------------
    def make_host_buffers(self):
        pin = dict(pin_memory=True)
        hb = {"x": [], "g": [], "y": [], "dx": []}
        for st in self.stages:
            shape = (self.batch, st.C, st.H, st.W)
            hb["x"].append(torch.randn(shape).pin_memory())
            hb["g"].append(torch.randn(shape).pin_memory())
            hb["y"].append(torch.empty(shape, **pin))
            hb["dx"].append(torch.empty(shape, **pin))
        hb["dw"] = torch.empty(self.grad_bucket.shape, **pin)
        return hb

    def step_host(self, hb):
        """host x, g -> device -> forward+backward -> host y, dX, dW.  Returns (h2d, d2h) bytes."""
        h2d = d2h = 0
        for st, x, g in zip(self.stages, hb["x"], hb["g"]):
            st.act[0].copy_(x, non_blocking=True)
            st.grad_in.copy_(g, non_blocking=True)
            h2d += x.numel() * 4 + g.numel() * 4
        self.step()
        for st, y, dx in zip(self.stages, hb["y"], hb["dx"]):
            y.copy_(st.act[st.n], non_blocking=True)
            dx.copy_(st.dx, non_blocking=True)
            d2h += y.numel() * 4 + dx.numel() * 4
        hb["dw"].copy_(self.grad_bucket, non_blocking=True)
        d2h += hb["dw"].numel() * 4
        torch.cuda.current_stream(self.device).synchronize()
        return h2d, d2h

    # -- accounting ----------------------------------------------
    def algorithmic_bytes_per_step(self):
        tot = 0
        for st in self.stages:
            N = self.batch * st.C * st.H * st.W
            tot += st.n * (20 * N + 12 * st.C * (st.C // st.groups) * st.k * st.k)
        return tot

    def algorithmic_flops_per_step(self):
        tot = 0
        for st in self.stages:
            N = self.batch * st.C * st.H * st.W
            tot += st.n * 6 * N * ((st.C // st.groups) * st.k * st.k - 1)
        return tot

    def wavefront_steps_per_step(self):
        return sum(st.n * 2 * (st.H + st.W - 1) for st in self.stages)
