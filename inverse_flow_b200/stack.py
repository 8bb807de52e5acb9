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
import os

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
            # every solve but the one right behind the stage's prepare may fetch its weights ahead of the
            # programmatic-launch wait (ifk.h: IFK_FLAG_STABLE_PREPARED)
            st.problem_stable = _native.with_flags(st.problem, _native.FLAG_STABLE_PREPARED)
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
        # dW stage 1 of the layers are independent of each other: several side streams let them overlap
        # (one stream would serialise them and become the critical path once the solves are fast)
        # (lowest priority: the dependent chain of solves on the capturing stream gets freed SMs first)
        self.sides = [torch.cuda.Stream(device=self.device, priority=0)
                      for _ in range(int(os.environ.get("IFK_STACK_SIDE_STREAMS", "8")))]
        self.reduce_stream = torch.cuda.Stream(device=self.device, priority=0)
        self.main = torch.cuda.Stream(device=self.device, priority=-1)      # the stream graphs are captured on
        self._side_rr = 0
        self._sides_forked = []           # side streams forked since the last join (a graph capture must
                                          # only join streams that are part of it)
        self.graph = None
        # per stage: prepare (T kernel, tap products for k > 1, + the wave kernel's weight pack) + n inverse + n dX
        # + n dW stage 1 + 1 dW stage 2
        self.launches_per_step = sum(3 * st.n + 2 + (1 if st.k > 1 else 0) +
                                     (1 if "wave<" in _native.describe_solve(st.problem) else 0) for st in self.stages)

    # -- raw launches ------------------------------------------------------------------
    def _stream(self):
        return _native.current_stream(self.device)

    def prepare_stage(self, st, stream=None):
        """T-fold (and pack) the weights of all the stage's layers: one batched launch pair"""
        s = self._stream() if stream is None else ctypes.c_void_p(stream.cuda_stream)
        _native.check(self.lib.ifk_prepare_many_f32(ctypes.byref(st.problem), st.n, st.w_base.data_ptr(), st.w_stride,
                                                    st.prepared_all.data_ptr(), st.pf, s))

    def prepare_all_forked(self):
        """the prepares of every stage depend on the weights only and are latency-bound (one small CTA per layer and
        tap): they run CONCURRENTLY, one stream per stage, and the first solve waits for all of them -- the critical
        path is one prepare, and no prepare CTA competes with the latency-bound solve chain afterwards"""
        main = torch.cuda.current_stream(self.device)
        self._prep_events = {}
        for k, st in enumerate(self.stages[1:]):
            side = self.sides[-1 - (k % len(self.sides))]
            side.wait_stream(main)
            self.prepare_stage(st, side)
            ev = torch.cuda.Event()
            ev.record(side)
            self._prep_events[id(st)] = ev
        self.prepare_stage(self.stages[0])
        for st in self.stages[1:]:               # join now: the solves start only when every prepare has finished
            main.wait_event(self._prep_events.pop(id(st)))

    def forward_stage(self, st, prepared=False):
        lib, s = self.lib, self._stream()
        p = ctypes.byref(st.problem)
        if not prepared:
            self.prepare_stage(st)
        elif id(st) in getattr(self, "_prep_events", {}):
            torch.cuda.current_stream(self.device).wait_event(self._prep_events.pop(id(st)))
        ps = ctypes.byref(st.problem_stable)
        for i in range(st.n):
            # the first solve follows the prepare (or the join with it): it fetches its weights after the
            # dependency wait; the others vouch that their predecessor did not write them
            _native.check(lib.ifk_inverse_f32(p if i == 0 else ps, st.act[i].data_ptr(), st.prepared[i].data_ptr(),
                                              st.act[i + 1].data_ptr(), s))

    def backward_stage(self, st):
        """dX chain on the current stream, dW stage 1 of every layer forked onto a side stream."""
        lib = self.lib
        main = torch.cuda.current_stream(self.device)
        s = ctypes.c_void_p(main.cuda_stream)
        p = ctypes.byref(st.problem)
        ps = ctypes.byref(st.problem_stable)      # the prepares ran in the forward pass, long before
        g = st.grad_in
        for i in reversed(range(st.n)):
            dx = st.dxs[i]
            _native.check(lib.ifk_bwd_input_f32(ps, g.data_ptr(), st.prepared[i].data_ptr(), dx.data_ptr(), s))
            side = self.sides[self._side_rr % len(self.sides)]
            self._side_rr += 1
            side.wait_stream(main)                            # fork: dW stage 1 needs this dX
            if side not in self._sides_forked:
                self._sides_forked.append(side)
            ws = st.workspace[i * st.ws_floats:]
            _native.check(lib.ifk_bwd_weight_partial_f32(p, dx.data_ptr(), st.act[i + 1].data_ptr(),
                                                         ws.data_ptr(), ctypes.c_void_p(side.cuda_stream)))
            g = dx
        st.dx = g

    def finish_weight_gradients(self, stages=None, local=False):
        """join the side streams, then one batched dW stage 2 per stage into the flat bucket (`local`: into the
        communicator's peer-mapped bucket, see attach_comm)."""
        main = torch.cuda.current_stream(self.device)
        for side in self._sides_forked:
            main.wait_stream(side)
        self._sides_forked = []
        s = ctypes.c_void_p(main.cuda_stream)
        for st in (self.stages if stages is None else stages):
            _native.check(self.lib.ifk_bwd_weight_reduce_many_f32(
                ctypes.byref(st.problem), st.n, st.workspace.data_ptr(), st.ws_floats * 4,
                (st.dw_local_base if local else st.dw_base).data_ptr(), st.w_stride, s))

    def forward(self):
        self.prepare_all_forked()
        for st in self.stages:
            self.forward_stage(st, prepared=True)

    def finish_stage_async(self, st, local=False):
        """stage 2 of dW for ONE stage on the reduce stream, behind the stage's dW kernels: it overlaps the next
        stage's dX chain instead of waiting at the end of the backward pass"""
        red = self.reduce_stream
        for side in self._sides_forked:
            red.wait_stream(side)
        _native.check(self.lib.ifk_bwd_weight_reduce_many_f32(
            ctypes.byref(st.problem), st.n, st.workspace.data_ptr(), st.ws_floats * 4,
            (st.dw_local_base if local else st.dw_base).data_ptr(), st.w_stride, ctypes.c_void_p(red.cuda_stream)))

    def backward(self):
        main = torch.cuda.current_stream(self.device)
        self.reduce_stream.wait_stream(main)          # fork (a capture must see the stream join and leave)
        for st in self.stages:
            self.backward_stage(st)
            self.finish_stage_async(st)
        main.wait_stream(self.reduce_stream)          # join: transitively every dW side stream
        self._sides_forked = []

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
            with torch.cuda.graph(self.graph, stream=self.main):
                self.forward_backward()
        return self

    def capture_bucketed(self):
        """Two graphs for data-parallel runs, in the order a real model produces gradients: graph A =
        forward of every stage + backward of the LAST stage (its slice of the bucket is then final),
        graph B = backward of the remaining stages.  The caller all-reduces the last stage's bucket
        slice asynchronously between the two replays, so that exchange overlaps graph B."""
        with torch.cuda.device(self.device):
            if self.graph is None:
                self.capture()                    # warm-up and the single-graph variant
            last, rest = self.stages[-1:], self.stages[:-1]
            self.graph_a, self.graph_b = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_a, stream=self.main):
                self.forward()
                for st in last:
                    self.backward_stage(st)
                self.finish_weight_gradients(last)
            with torch.cuda.graph(self.graph_b, stream=self.main):
                for st in rest:
                    self.backward_stage(st)
                self.finish_weight_gradients(rest)
            n_last = sum(st.n * st.w_stride for st in last)
            n_all = self.grad_bucket.numel()
            self.bucket_last = self.grad_bucket[n_all - n_last:]
            self.bucket_rest = self.grad_bucket[:n_all - n_last]
        return self

    def attach_comm(self, comm):
        """Data-parallel mode: the stage-2 reductions write this rank's dW into `comm.bucket` (peer-mapped
        memory, parallel.PeerAllReduce) and the fused peer all-reduce leaves the sum over ranks in
        `grad_bucket` -- so the exchange needs no copy in between."""
        if comm.numel != self.grad_bucket.numel():
            raise ValueError("communicator bucket size differs from the gradient bucket")
        self.comm = comm
        off = 0
        for st in self.stages:
            st.dw_local_base = comm.bucket[off:]
            off += st.n * st.w_stride
        return self

    def capture_parallel(self):
        """ONE graph per step for data-parallel runs: forward, backward, and the gradient exchange as two
        launches of the fused peer all-reduce kernel -- the last stage's slice of the bucket (final first, as
        in a real model) on a side stream while the remaining backward runs, the rest at the end."""
        comm = self.comm
        with torch.cuda.device(self.device):
            if self.graph is None:
                self.capture()                    # warm-up (sets function attributes) and the local variant
            last, rest = self.stages[-1:], self.stages[:-1]
            n_last = sum(st.n * st.w_stride for st in last)
            n_all = self.grad_bucket.numel()
            split = n_all - n_last
            # one exchange at the end when the slices would not start at multiples of 4 floats, or when the bucket
            # is so small that a second hand-shake with the peers costs more than overlapping it could hide
            if split % 4 or n_all * 4 < 256 * 1024:
                last, rest, split = self.stages, [], 0

            def run():
                main = torch.cuda.current_stream(self.device)
                red = self.reduce_stream
                self.forward()
                red.wait_stream(main)                            # fork the reduce stream
                for st in last:
                    self.backward_stage(st)
                    self.finish_stage_async(st, local=True)      # stage 2 behind the stage's dW kernels, off the main stream
                side = self.sides[-1]
                if rest:
                    side.wait_stream(red)                        # fork: the last stage's slice travels now
                    with torch.cuda.stream(side):
                        comm.allreduce(self.grad_bucket, offset=split, numel=n_all - split)
                    for st in rest:
                        self.backward_stage(st)
                        self.finish_stage_async(st, local=True)
                    main.wait_stream(red)
                    comm.allreduce(self.grad_bucket, offset=0, numel=split)
                    main.wait_stream(side)                       # join
                else:
                    main.wait_stream(red)
                    comm.allreduce(self.grad_bucket)
                self._sides_forked = []

            warm = torch.cuda.Stream()
            warm.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(warm):
                run()
            torch.cuda.current_stream().wait_stream(warm)
            torch.cuda.synchronize()
            self.graph_parallel = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_parallel, stream=self.main):
                run()
        return self

    def step_parallel(self):
        self.graph_parallel.replay()

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

    def _host_pipeline(self, hb, parallel=False):
        """x, g from pinned host memory -> forward + backward -> y, dX, dW to pinned host memory, with
        the copies on their own stream: stage k+1's upload and stage k's download overlap compute."""
        main = torch.cuda.current_stream(self.device)
        cp = self.copy
        cp.wait_stream(main)                                         # fork
        ev_x, ev_g = [], []
        with torch.cuda.stream(cp):
            for st, x in zip(self.stages, hb["x"]):
                st.act[0].copy_(x, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(cp)
                ev_x.append(ev)
            for st, g in zip(self.stages, hb["g"]):
                st.grad_in.copy_(g, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(cp)
                ev_g.append(ev)
        self.prepare_all_forked()                  # weights only: runs while the first images are still in flight
        for k, st in enumerate(self.stages):
            main.wait_event(ev_x[k])
            self.forward_stage(st, prepared=True)
            done = torch.cuda.Event()
            done.record(main)
            with torch.cuda.stream(cp):
                cp.wait_event(done)
                hb["y"][k].copy_(st.act[st.n], non_blocking=True)
        for k, st in enumerate(self.stages):
            main.wait_event(ev_g[k])
            self.backward_stage(st)
            done = torch.cuda.Event()
            done.record(main)
            with torch.cuda.stream(cp):
                cp.wait_event(done)
                hb["dx"][k].copy_(st.dx, non_blocking=True)
        if parallel:                                                 # data-parallel: this rank's dW into the peer-mapped
            self.finish_weight_gradients(local=True)                 # bucket, then the fused all-reduce over the ranks
            self.comm.allreduce(self.grad_bucket)
        else:
            self.finish_weight_gradients()
        hb["dw"].copy_(self.grad_bucket, non_blocking=True)
        main.wait_stream(cp)                                         # join

    def capture_host(self, hb, parallel=False):
        """the whole host-to-host step as ONE CUDA graph (memcpy nodes included; `parallel`: with the gradient
        exchange over the ranks -- every rank must capture and replay it the same number of times)."""
        with torch.cuda.device(self.device):
            self.copy = torch.cuda.Stream(device=self.device)
            warm = torch.cuda.Stream()
            warm.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(warm):
                self._host_pipeline(hb, parallel)
            torch.cuda.current_stream().wait_stream(warm)
            torch.cuda.synchronize()
            self.graph_host = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_host, stream=self.main):
                self._host_pipeline(hb, parallel)
        self._host_buffers = hb
        return self

    def step_host(self, hb, parallel=False):
        """host x, g -> device -> forward+backward -> host y, dX, dW.  Returns (h2d, d2h) bytes."""
        if getattr(self, "graph_host", None) is None or self._host_buffers is not hb:
            self.capture_host(hb, parallel)
        self.graph_host.replay()
        torch.cuda.current_stream(self.device).synchronize()
        n_act = sum(x.numel() for x in hb["x"])
        return 2 * n_act * 4, 2 * n_act * 4 + hb["dw"].numel() * 4

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
