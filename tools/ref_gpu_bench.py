"""Time the reference's own CUDA extension beside our kernels on the same GPU.

    python tools/ref_gpu_bench.py [--out file.jsonl] [--iters N] [--skip-dw]

The reference extension (oracle/_ref/, built by oracle/build_ref_cuda.py from the reference's
sources) blocks the host inside every call ((H+W-1)*C/4 launches, each followed by
cudaDeviceSynchronize: inv_conv_with_bp_kernel_general.cu:97-127), so it is timed by wall
clock around synchronised calls; ours is timed with CUDA events around a stream of calls.
Shapes: the inverse-conv layers of the reference's models with C % 4 == 0 (the literal kernels
launch nothing for C < 4) and B >= C (its dw indexes samples as weight rows).
One JSON line per shape: microseconds for inverse / dy / dw of the reference and for
inverse / bwd_input / bwd_weight of this library (groups=4, the literal grouping), and the
fwd+bwd images/s of both.  The reference's dw runs in a child process first: it is known to
index out of bounds for some shapes, and a fault must not take the parent's context with it.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_cuda  # noqa: E402
from inverse_flow_b200 import functional as IF  # noqa: E402

SHAPES = [
    # B, C, H, k      where in the reference
    (100, 4, 14, 2),   # if_glow_mnist level 1 (experiments/if_glow_mnist.py:58-60)
    (100, 8, 7, 2),    # if_glow_mnist level 2
    (100, 12, 16, 3),  # if_multiGPU_imagenet32 level 1 (inf/if_multiGPU_imagenet32.py:196-199)
    (100, 24, 8, 3),   # level 2
    (100, 48, 4, 3),   # level 3
    (256, 12, 16, 3),  # if_glow_cifar level 1, BASELINE batch
]


def make_weight(C, k, gen):
    w = torch.zeros(C, C, k, k)
    torch.nn.init.dirac_(w)
    w += torch.nn.init.xavier_normal_(torch.empty(C, C, k, k), gain=0.01, generator=gen)
    w[:, -1, -1, -1] = 1.0
    return w


def wall_us(fn, iters):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters * 1e6


def event_us(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3


def ref_dw_child(B, C, H, k, iters):
    ref = ref_cuda.load()
    gen = torch.Generator().manual_seed(0)
    w = make_weight(C, k, gen).cuda()
    torch.manual_seed(0)
    x = torch.randn(B, C, H, H, device="cuda")
    g = torch.randn(B, C, H, H, device="cuda")

    def call():
        # the allocations the reference's backward performs per call (layers/inv_conv.py:70-79)
        M = torch.zeros((B, C, k, k, H, H)).to("cuda")
        out = torch.zeros_like(w)
        ref.dw(x, w, g, M, out)

    print(json.dumps({"ref_dw_us": wall_us(call, iters)}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--skip-dw", action="store_true")
    ap.add_argument("--dw-child", nargs=4, type=int, default=None)
    ap.add_argument("--shapes", default=None, help='"B,C,H,k;B,C,H,k;..." instead of the built-in list')
    ap.add_argument("--ref-only", action="store_true", help="time the reference extension only")
    args = ap.parse_args()
    if args.dw_child:
        ref_dw_child(*args.dw_child, args.iters)
        return 0
    ref = ref_cuda.load()
    if ref is None:
        print(json.dumps({"unavailable": "oracle/_ref/ has no compiled reference extension"}))
        return 0
    gen = torch.Generator().manual_seed(0)
    lines = []
    shapes = SHAPES if not args.shapes else [tuple(int(v) for v in sh.split(",")) for sh in args.shapes.split(";")]
    for (B, C, H, k) in shapes:
        w = make_weight(C, k, gen).cuda()
        torch.manual_seed(0)
        x = torch.randn(B, C, H, H, device="cuda")
        g = torch.randn(B, C, H, H, device="cuda")
        rec = {"shape": [B, C, H, H], "k": k, "groups": 4}

        rec["ref_inverse_us"] = wall_us(lambda: ref.inverse(x, w, x * 0.0), args.iters)
        rec["ref_dy_us"] = wall_us(lambda: ref.dy(g, w, torch.zeros_like(g), torch.zeros_like(g)), args.iters)
        rec["ref_dw_us"] = None
        if not args.skip_dw:
            try:
                out = subprocess.run([sys.executable, os.path.abspath(__file__), "--iters", str(max(1, args.iters // 2)),
                                      "--dw-child", str(B), str(C), str(H), str(k)],
                                     capture_output=True, text=True, timeout=300)
                for ln in out.stdout.splitlines():
                    if ln.startswith("{"):
                        rec["ref_dw_us"] = json.loads(ln)["ref_dw_us"]
                if rec["ref_dw_us"] is None:
                    rec["ref_dw_error"] = (out.stderr or "no output")[-300:]
            except subprocess.TimeoutExpired:
                rec["ref_dw_error"] = "timeout"

        if args.ref_only:
            print(json.dumps(rec), flush=True)
            lines.append(rec)
            continue
        prep = IF.prepare(w, groups=4)
        y = IF.inverse(x, w, groups=4, prepared=prep)
        dx = IF.bwd_input(g, w, groups=4, prepared=prep)
        out_y, out_dx, out_dw = torch.empty_like(x), torch.empty_like(x), torch.empty_like(w)
        rec["ours_prepare_us"] = event_us(lambda: IF.prepare(w, groups=4), 50)
        rec["ours_inverse_us"] = event_us(lambda: IF.inverse(x, w, groups=4, out=out_y, prepared=prep), 50)
        rec["ours_bwd_input_us"] = event_us(lambda: IF.bwd_input(g, w, groups=4, out=out_dx, prepared=prep), 50)
        rec["ours_bwd_weight_us"] = event_us(lambda: IF.bwd_weight(dx, y, w, groups=4, out=out_dw), 50)
        ours = rec["ours_prepare_us"] + rec["ours_inverse_us"] + rec["ours_bwd_input_us"] + rec["ours_bwd_weight_us"]
        rec["ours_fwd_bwd_images_per_s"] = B / (ours * 1e-6)
        if rec["ref_dw_us"] is not None:
            reft = rec["ref_inverse_us"] + rec["ref_dy_us"] + rec["ref_dw_us"]
            rec["ref_fwd_bwd_images_per_s"] = B / (reft * 1e-6)
        rec["ref_inverse_dy_images_per_s"] = B / ((rec["ref_inverse_us"] + rec["ref_dy_us"]) * 1e-6)
        rec["ours_inverse_dx_images_per_s"] = B / ((rec["ours_inverse_us"] + rec["ours_bwd_input_us"]) * 1e-6)
        # same answer where the literal op is the math contract (C == 4)
        if C == 4:
            y_ref = ref.inverse(x, w, x * 0.0)[0]
            rec["inverse_max_abs_diff_vs_ref"] = float((y_ref - y).abs().max())
        print(json.dumps(rec), flush=True)
        lines.append(rec)
    if args.out:
        with open(args.out, "w") as f:
            for r in lines:
                f.write(json.dumps(r) + "\n")
    return 0


if __name__ == "__main__":
    sys.exit(main())
