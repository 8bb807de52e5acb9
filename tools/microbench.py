"""Per-kernel timings over a list of shapes (CUDA events, current stream).

    python tools/microbench.py [--shapes model|sweep|all] [--iters N] [--out file.jsonl]

For each shape: prepare, inverse, bwd_input, bwd_weight, conv; reports microseconds,
images/s, achieved algorithmic GB/s and GFLOP/s.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inverse_flow_b200 import _native, functional as IF  # noqa: E402

MODEL = [
    (64, 1, 28, 28, 3, 1), (100, 4, 14, 14, 2, 1), (100, 8, 7, 7, 2, 1), (100, 4, 14, 14, 2, 4),
    (100, 8, 7, 7, 2, 4), (100, 12, 16, 16, 3, 1), (100, 24, 8, 8, 3, 1), (100, 48, 4, 4, 3, 1),
    (100, 12, 16, 16, 3, 4), (100, 24, 8, 8, 3, 4), (100, 48, 4, 4, 3, 4),
    (256, 12, 16, 16, 3, 1), (256, 24, 8, 8, 3, 1),
]
SWEEP = [
    (512, 3, 16, 16, 3, 1), (512, 3, 32, 32, 3, 1), (512, 3, 64, 64, 3, 1), (512, 3, 64, 64, 5, 1),
    (512, 12, 16, 16, 3, 1), (512, 12, 32, 32, 3, 1), (64, 12, 64, 64, 3, 1), (512, 12, 16, 16, 5, 1),
    (512, 48, 16, 16, 3, 1), (64, 48, 32, 32, 3, 1), (64, 96, 16, 16, 3, 1), (8, 96, 32, 32, 3, 1),
]


LARGE = [   # sweep cells whose images do not fit in shared memory (window / stream / fallback kernels)
    (64, 12, 64, 64, 3, 1), (512, 12, 64, 64, 3, 1), (64, 12, 64, 64, 5, 1), (64, 12, 64, 64, 7, 1),
    (64, 48, 32, 32, 3, 1), (512, 48, 32, 32, 3, 1), (64, 48, 64, 64, 3, 1), (64, 48, 16, 16, 5, 1),
    (64, 48, 32, 32, 5, 1), (64, 48, 16, 16, 7, 1), (64, 96, 16, 16, 3, 1), (8, 96, 32, 32, 3, 1),
    (64, 96, 32, 32, 3, 1), (64, 96, 16, 16, 5, 1), (64, 3, 128, 128, 3, 1),
]


def make_weight(C, k, gen):
    w = torch.zeros(C, C, k, k)
    torch.nn.init.dirac_(w)
    w += torch.nn.init.xavier_normal_(torch.empty(C, C, k, k), gain=0.01, generator=gen)
    w[:, -1, -1, -1] = 1.0
    return w


def Cg_of(C, g):
    return C // g


def time_op(fn, iters, flush=None, reps=20):
    """microseconds per call: `reps` calls captured in one CUDA graph (no Python / launch
    overhead between them), replayed `iters` times, CUDA events around each replay."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(reps):
            fn()
    graph.replay()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        graph.replay()
        e.record()
        torch.cuda.synchronize()
        tot += s.elapsed_time(e) * 1e3
    return tot / iters / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="model")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--out", default=None)
    ap.add_argument("--flush", action="store_true", help="flush L2 between iterations (cold)")
    ap.add_argument("--solve-only", action="store_true", help="time the forward solve only")
    ap.add_argument("--shape", default=None, help="one shape B,C,H,W,k,g instead of a list")
    args = ap.parse_args()
    shapes = {"model": MODEL, "sweep": SWEEP, "large": LARGE, "all": MODEL + SWEEP}[args.shapes]
    if args.shape:
        shapes = [tuple(int(v) for v in args.shape.split(","))]
    gen = torch.Generator().manual_seed(0)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda") if args.flush else None
    lib = _native.load()
    rows = []
    print("%-28s %8s %8s %8s %8s %8s | %10s %8s %8s  %s" % ("shape(B,C,H,W,k,g)", "prep", "inv", "dX", "dW", "conv",
                                                          "fb img/s", "fbGB/s", "GFLOP/s", "variant"))
    for (B, C, H, W, k, g) in shapes:
        x = torch.randn(B, C, H, W, device="cuda")
        grad = torch.randn(B, C, H, W, device="cuda")
        w = make_weight(C, k, gen).cuda()
        prep = IF.Prepared(w, g)
        y = IF.inverse(x, w, prepared=prep)
        dx = torch.empty_like(x)
        out = torch.empty_like(x)
        dw = torch.empty_like(w)
        reps = 20 if B * C * H * W * Cg_of(C, g) * k * k < 2e10 else 2
        if args.solve_only:
            t_inv = time_op(lambda: IF.inverse(x, w, out=out, prepared=prep), args.iters, flush, reps)
            N = B * C * H * W
            flops = 2 * N * (C // g * k * k - 1)
            desc = _native.describe_solve(_native.problem(B, C, H, W, k, k, C, g))
            rows.append(dict(shape=[B, C, H, W, k, g], us=dict(inverse=t_inv), inverse_GFLOPs=flops / t_inv / 1e3,
                             inverse_images_per_s=B / (t_inv * 1e-6), variant=desc))
            print("%-28s inverse %10.1f us %9.1f GFLOP/s  %s" % (str((B, C, H, W, k, g)), t_inv, flops / t_inv / 1e3, desc),
                  flush=True)
            continue
        t_prep = time_op(lambda: IF.Prepared(w, g), args.iters, flush)
        t_inv = time_op(lambda: IF.inverse(x, w, out=out, prepared=prep), args.iters, flush)
        t_dx = time_op(lambda: IF.bwd_input(grad, w, out=dx, prepared=prep), args.iters, flush)
        t_dw = time_op(lambda: IF.bwd_weight(dx, y, w, groups=g, out=dw), args.iters, flush)
        t_conv = time_op(lambda: IF.conv(y, w, groups=g, out=out), args.iters, flush)
        N = B * C * H * W
        Cg = C // g
        K = k * k
        t_fb = t_prep + t_inv + t_dx + t_dw
        bytes_fb = 20 * N + 12 * C * Cg * K
        flops_fb = 6 * N * (Cg * K - 1)
        desc = _native.describe_solve(_native.problem(B, C, H, W, k, k, C, g))
        row = dict(shape=[B, C, H, W, k, g], us=dict(prepare=t_prep, inverse=t_inv, bwd_input=t_dx, bwd_weight=t_dw,
                                                      conv=t_conv), fwd_bwd_images_per_s=B / (t_fb * 1e-6),
                   fwd_bwd_GBps=bytes_fb / (t_fb * 1e-6) / 1e9, fwd_bwd_GFLOPs=flops_fb / (t_fb * 1e-6) / 1e9,
                   inverse_GBps=4 * (2 * N + C * Cg * K) / (t_inv * 1e-6) / 1e9, variant=desc)
        rows.append(row)
        print("%-28s %8.1f %8.1f %8.1f %8.1f %8.1f | %10.0f %8.1f %8.1f  %s" % (
            str((B, C, H, W, k, g)), t_prep, t_inv, t_dx, t_dw, t_conv, row["fwd_bwd_images_per_s"],
            row["fwd_bwd_GBps"], row["fwd_bwd_GFLOPs"], desc))
    if args.out:
        with open(args.out, "w") as f:
            for r in rows:
                f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
