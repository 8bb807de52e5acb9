"""Launch time of dW stage 1 (ifk_bwd_weight_partial_f32) at one geometry: 32 launches back to back in a graph.
    python tools/probe_dw.py B C H W k groups
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inverse_flow_b200 import _native  # noqa: E402


def main():
    B, C, H, W, k, g = (int(v) for v in sys.argv[1:7])
    lib = _native.load()
    p = _native.problem(B, C, H, W, k, k, C, g)
    dx = torch.randn(B, C, H, W, device="cuda")
    y = torch.randn(B, C, H, W, device="cuda")
    ws = torch.empty(lib.ifk_bwd_weight_workspace_bytes(ctypes.byref(p)) // 4 + 4, device="cuda")
    s = _native.current_stream(dx.device)
    _native.check(lib.ifk_bwd_weight_partial_f32(ctypes.byref(p), dx.data_ptr(), y.data_ptr(), ws.data_ptr(), s))
    torch.cuda.synchronize()
    reps = 32
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        s2 = _native.current_stream(dx.device)
        for _ in range(reps):
            _native.check(lib.ifk_bwd_weight_partial_f32(ctypes.byref(p), dx.data_ptr(), y.data_ptr(), ws.data_ptr(), s2))
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps * 1e3)
    flops = 2.0 * B * C * H * W * (C // g) * k * k
    print("%s dW stage 1: %.2f us per launch, %.1f TFLOP/s, workspace %d KB (IFK_DW_QUAD=%s)" % (
        (B, C, H, W, k, g), best, flops / best / 1e6, ws.numel() * 4 // 1024, os.environ.get("IFK_DW_QUAD", "")))


if __name__ == "__main__":
    main()
