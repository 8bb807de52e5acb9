"""Where a step's time goes: the stack's forward, backward-dX chain and dW timed separately and together
(each as its own CUDA graph, L2 flushed, CUDA events).

    python tools/step_breakdown.py [--workload glow_imagenet32] [--groups 1]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS  # noqa: E402
from inverse_flow_b200.stack import InvConvStack  # noqa: E402


def graph_of(fn, stream=None):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=stream):
        fn()
    return g


def timed(g, flush, reps=20):
    for _ in range(3):
        g.replay()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        g.replay()
        e.record()
        e.synchronize()
        tot += s.elapsed_time(e)
    return tot / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="glow_imagenet32", choices=sorted(WORKLOADS))
    ap.add_argument("--groups", type=int, default=1)
    args = ap.parse_args()
    stages, batch, desc = WORKLOADS[args.workload]
    stack = InvConvStack(stages, batch, groups=args.groups or None)
    for st in stack.stages:
        st.act[0].normal_()
        st.grad_in.normal_()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def bwd_dx_only():
        import ctypes
        from inverse_flow_b200 import _native
        for st in stack.stages:
            ps = ctypes.byref(st.problem_stable)
            g = st.grad_in
            for i in reversed(range(st.n)):
                _native.check(stack.lib.ifk_bwd_input_f32(ps, g.data_ptr(), st.prepared[i].data_ptr(),
                                                          st.dxs[i].data_ptr(), stack._stream()))
                g = st.dxs[i]

    def bwd_dx_with_event_records():
        """the dX chain with an (unused) event recorded after every solve: does a record between two kernels of the
        capturing stream cost the programmatic dependent launch overlap?"""
        import ctypes
        from inverse_flow_b200 import _native
        main = torch.cuda.current_stream()
        for st in stack.stages:
            ps = ctypes.byref(st.problem_stable)
            g = st.grad_in
            for i in reversed(range(st.n)):
                _native.check(stack.lib.ifk_bwd_input_f32(ps, g.data_ptr(), st.prepared[i].data_ptr(),
                                                          st.dxs[i].data_ptr(), stack._stream()))
                ev = torch.cuda.Event()
                ev.record(main)
                g = st.dxs[i]

    def bwd_dx_with_forked_noops():
        """the dX chain with a fork after every solve whose side stream runs a trivial kernel"""
        import ctypes
        from inverse_flow_b200 import _native
        main = torch.cuda.current_stream()
        sides = stack.sides
        k = 0
        for st in stack.stages:
            ps = ctypes.byref(st.problem_stable)
            g = st.grad_in
            for i in reversed(range(st.n)):
                _native.check(stack.lib.ifk_bwd_input_f32(ps, g.data_ptr(), st.prepared[i].data_ptr(),
                                                          st.dxs[i].data_ptr(), stack._stream()))
                side = sides[k % len(sides)]
                k += 1
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    tiny.add_(1.0)
                g = st.dxs[i]
        for side in sides:
            main.wait_stream(side)

    tiny = torch.zeros(32, device="cuda")

    def dw_only():
        import ctypes
        from inverse_flow_b200 import _native
        for st in stack.stages:
            p = ctypes.byref(st.problem)
            for i in range(st.n):
                ws = st.workspace[i * st.ws_floats:]
                _native.check(stack.lib.ifk_bwd_weight_partial_f32(p, st.dxs[i].data_ptr(), st.act[i + 1].data_ptr(),
                                                                   ws.data_ptr(), stack._stream()))
        stack.finish_weight_gradients()

    def fwd_solves_only():
        for st in stack.stages:
            stack.forward_stage(st, prepared=True)

    def prepares_only():
        for st in stack.stages:
            stack.prepare_stage(st)

    stack.forward_backward()
    torch.cuda.synchronize()
    res = {"workload": args.workload, "layers": sum(st.n for st in stack.stages)}
    res["forward_ms"] = timed(graph_of(stack.forward), flush)
    res["forward_main_stream_graph_ms"] = timed(graph_of(stack.forward, stack.main), flush)
    res["forward_solves_only_ms"] = timed(graph_of(fwd_solves_only), flush)
    res["prepares_only_ms"] = timed(graph_of(prepares_only), flush)
    res["backward_dx_chain_ms"] = timed(graph_of(bwd_dx_only), flush)
    res["backward_dx_chain_event_records_ms"] = timed(graph_of(bwd_dx_with_event_records, stack.main), flush)
    res["backward_dx_chain_forked_noops_ms"] = timed(graph_of(bwd_dx_with_forked_noops, stack.main), flush)
    res["dw_serial_one_stream_ms"] = timed(graph_of(dw_only), flush)
    res["backward_full_ms"] = timed(graph_of(stack.backward, stack.main), flush)
    res["step_ms"] = timed(graph_of(stack.forward_backward, stack.main), flush)
    res["images_per_s"] = batch / (res["step_ms"] * 1e-3)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
