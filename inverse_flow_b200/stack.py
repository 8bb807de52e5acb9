"""InvConvStack: the inverse-convolution layers of an if_* Glow model as one static pipeline.

The if_glow models interleave `inv_flow_*` layers with ActNorm / Spline / Coupling layers
(reference experiments/if_glow_mnist.py:62-124).  This class holds ONLY the inverse-conv
layers of such a model -- per stage `n_layers` layers of one (C, H, W, k) shape, chained --
with every buffer preallocated, all launches going straight through the C ABI, and the whole
forward + backward captured in one CUDA graph (B200-first: streams and graphs instead of a
per-layer Python loop).  It is what bench.py times and what a training driver would replay.

Per step            : one batched prepare per stage (weights do not depend on activations)
Forward per layer   : y = L^-1 x                            (x of layer i+1 = y of layer i)
Backward per layer  : dX = L^-T g on the main stream        (g of layer i-1 = dX of layer i);
                      stage 1 of dW = -corr(dX, y) on a SIDE stream, overlapping the next
                      layer's solve (a fork/join in the captured graph);
End of backward     : one batched stage-2 reduction per stage writes dW of every layer into
                      one flat bucket (`grad_bucket`), i.e. directly into the buffer a
                      data-parallel all-reduce consumes.
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
            sz = C * C * k * k
            st.w_base, st.dw_base, st.w_stride = self.weights[off:], self.grad_bucket[off:], sz
            st.w = [self.weights[off + i * sz: off + (i + 1) * sz].view(C, C, k, k) for i in range(n)]
            st.dw = [self.grad_bucket[off + i * sz: off + (i + 1) * sz].view(C, C, k, k) for i in range(n)]
            for wt in st.w:
                wt.copy_(reference_init_weight(C, k, gen))
            off += n * sz
            st.prepared_all = torch.empty(n * pf, dtype=torch.float32, device=self.device)
            st.prepared = [st.prepared_all[i * pf:(i + 1) * pf] for i in range(n)]
            st.pf = pf
            shape = (self.batch, C, H, W)
            st.act = [torch.zeros(shape, device=self.device) for _ in range(n + 1)]   # act[0] = x
            st.dxs = [torch.zeros(shape, device=self.device) for _ in range(n)]       # dX of every layer
            st.grad_in = torch.zeros(shape, device=self.device)                       # upstream g
            st.ws_floats = (ws + 15) // 16 * 4                                        # 16-byte multiple
            st.workspace = torch.empty(n * st.ws_floats, dtype=torch.float32, device=self.device)
            self.stages.append(st)
        self.side = torch.cuda.Stream(device=self.device)
        self.graph = None
        # per stage: 1 prepare + n inverse + n dX + n dW stage 1 + 1 dW stage 2
        self.launches_per_step = sum(3 * st.n + 2 for st in self.stages)

    # -- raw launches ------------------------------------------------------------------
    def _stream(self):
        return _native.current_stream(self.device)

    def forward(self):
        lib, s = self.lib, self._stream()
        for st in self.stages:
            p = ctypes.byref(st.problem)
            _native.check(lib.ifk_prepare_many_f32(p, st.n, st.w_base.data_ptr(), st.w_stride,
                                                   st.prepared_all.data_ptr(), st.pf, s))
            for i in range(st.n):
                _native.check(lib.ifk_inverse_f32(p, st.act[i].data_ptr(), st.prepared[i].data_ptr(),
                                                  st.act[i + 1].data_ptr(), s))

    def backward(self):
        lib = self.lib
        main = torch.cuda.current_stream(self.device)
        s = ctypes.c_void_p(main.cuda_stream)
        side_s = ctypes.c_void_p(self.side.cuda_stream)
        for st in self.stages:
            p = ctypes.byref(st.problem)
            g = st.grad_in
            for i in reversed(range(st.n)):
                dx = st.dxs[i]
                _native.check(lib.ifk_bwd_input_f32(p, g.data_ptr(), st.prepared[i].data_ptr(), dx.data_ptr(), s))
                self.side.wait_stream(main)                       # fork: dW stage 1 needs this dX
                ws = st.workspace[i * st.ws_floats:]
                _native.check(lib.ifk_bwd_weight_partial_f32(p, dx.data_ptr(), st.act[i + 1].data_ptr(),
                                                             ws.data_ptr(), side_s))
                g = dx
            st.dx = g
        main.wait_stream(self.side)                               # join
        for st in self.stages:
            _native.check(lib.ifk_bwd_weight_reduce_many_f32(
                ctypes.byref(st.problem), st.n, st.workspace.data_ptr(), st.ws_floats * 4,
                st.dw_base.data_ptr(), st.w_stride, s))

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
