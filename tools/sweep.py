"""The microbenchmark sweep of BASELINE.json configs[4] with the CPU path beside every cell.

    python tools/sweep.py [--out profiles/r02_sweep.jsonl] [--batches 1,8,64,256,512] [--quick]

Grid: C in {3,12,48,96} x H=W in {16,32,64} x k in {3,5,7} x batch x groups in {1,4} (groups=4 where 4 | C:
the reference kernels' grouping; harness analogue: reference inf/experiments/if_timescaling.py:98-151).
Per cell, one JSON line:
  * GPU: microseconds of prepare / inverse / dX / dW through the public functional API (CUDA events around a
    CUDA graph of `reps` calls, inputs resident), fwd+bwd images/s, the solve kernel that served the cell;
  * the three roofline terms of the inverse solve (algorithmic bytes / measured HBM peak, flops / measured FP32
    peak, (H+W-1) x the dependent chain of a step / SM clock) and the binding one;
  * CPU: the float32 oracle port (OpenMP, all host cores of this process) on a bounded sample of the same cell
    (fwd + bwd of n_cpu <= B images, about a second of work), scaled to images/s;
  * parity: conv(inverse(x)) - x max relative error on the GPU, and inverse vs the float64 oracle on one image.
"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from inverse_flow_b200 import _native, functional as IF  # noqa: E402
from inverse_flow_b200.stack import reference_init_weight  # noqa: E402
from oracle import oracle  # noqa: E402


def time_graph(fn, reps, iters):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(reps):
            fn()
    graph.replay()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        graph.replay()
        e.record()
        e.synchronize()
        best = min(best, s.elapsed_time(e) * 1e3 / reps)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="profiles/r02_sweep.jsonl")
    ap.add_argument("--batches", default="1,8,64,256,512")
    ap.add_argument("--quick", action="store_true", help="C in {3,12}, H=16 only (smoke run of the tool)")
    ap.add_argument("--cpu-seconds", type=float, default=0.6, help="CPU work per cell, about")
    args = ap.parse_args()
    batches = [int(b) for b in args.batches.split(",")]
    Cs, Hs, ks = (3, 12, 48, 96), (16, 32, 64), (3, 5, 7)
    if args.quick:
        Cs, Hs = (3, 12), (16,)
    threads = len(os.sched_getaffinity(0))
    hw = _native.hw_microbench(torch.device("cuda"))
    fp32_peak = max(hw["fp32_tflops_ffma"], hw["fp32_tflops_ffma2"])
    hbm_peak, _ = bench.measured_peak_gbs()
    sm_mhz = 1965.0
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    f = open(args.out, "w")
    f.write(json.dumps({"what": "inverse-conv microbenchmark sweep, GPU (B200) beside the CPU oracle port", "cpu_threads": threads,
                        "fp32_peak_tflops": fp32_peak, "hbm_peak_gbs": hbm_peak, "sm_mhz": sm_mhz,
                        "latency_cycles": hw["latency_cycles"]}) + "\n")
    gen = torch.Generator().manual_seed(0)
    cpu_rate = {}                                      # (C, H, k, g) -> seconds per image (fwd + bwd), measured once
    for C in Cs:
        for H in Hs:
            for k in ks:
                for g in (1, 4):
                    if C % g:
                        continue
                    w = reference_init_weight(C, k, gen)
                    wd = w.cuda()
                    Cg, K, N1 = C // g, k * k, C * H * H
                    # ---- CPU leg, once per (C, H, k, g): a bounded sample, scaled
                    flops_img = 6.0 * N1 * (Cg * K - 1)
                    n_cpu = int(max(1, min(64, args.cpu_seconds * 2e9 * min(threads, 8) / flops_img)))
                    xs = np.random.default_rng(1).standard_normal((n_cpu, C, H, H)).astype(np.float32)
                    gs = np.random.default_rng(2).standard_normal((n_cpu, C, H, H)).astype(np.float32)
                    w32 = w.numpy().astype(np.float32)
                    t0 = time.perf_counter()
                    y_cpu = oracle.inverse(xs, w32, g, threads=min(threads, n_cpu))
                    oracle.backward(gs, y_cpu, w32, g, threads=min(threads, n_cpu))
                    cpu_s = time.perf_counter() - t0
                    # throughput of the whole host: images are independent, so `threads` images run at once
                    cpu_img_s = n_cpu / cpu_s * (threads / min(threads, n_cpu))
                    y64 = oracle.inverse(xs[:1].astype(np.float64), w32.astype(np.float64), g)
                    for B in batches:
                        if B * N1 * 4 * 6 > 24e9 or B * flops_img / 3.0 > 3e11:
                            continue                                      # (cells beyond ~10 s per solve on the fallback kernel)
                        p = _native.problem(B, C, H, H, k, k, C, g)
                        variant = _native.describe_solve(p)
                        x = torch.randn(B, C, H, H, device="cuda")
                        x[:1] = torch.from_numpy(xs[:1]).cuda()
                        grad = torch.randn(B, C, H, H, device="cuda")
                        prep = IF.Prepared(wd, g)
                        y = IF.inverse(x, wd, prepared=prep)
                        rec = IF.conv(y, wd, groups=g)
                        torch.cuda.synchronize()
                        rt_err = float((rec - x).abs().max() / x.abs().max())
                        inv_err = oracle.max_rel_err(y[:1].cpu().numpy(), y64)
                        dx, out, dw = torch.empty_like(x), torch.empty_like(x), torch.empty_like(wd)
                        work = B * flops_img / 3.0                        # flops of one solve
                        reps = 16 if work < 5e9 else (4 if work < 1e11 else 1)
                        iters = 3 if work < 1e11 else 1
                        t_prep = time_graph(lambda: IF.Prepared(wd, g), min(reps, 4), iters)
                        t_inv = time_graph(lambda: IF.inverse(x, wd, out=out, prepared=prep), reps, iters)
                        t_dx = time_graph(lambda: IF.bwd_input(grad, wd, out=dx, prepared=prep), reps, iters)
                        t_dw = time_graph(lambda: IF.bwd_weight(dx, y, wd, groups=g, out=dw), reps, iters)
                        t_fb = t_prep + t_inv + t_dx + t_dw
                        bytes_inv = 4.0 * (2 * B * N1 + C * Cg * K)
                        flops_inv = 2.0 * B * N1 * (Cg * K - 1)
                        chain, _ = bench.chain_cycles(variant, hw["latency_cycles"], Cg, k)
                        roof = bench.solve_roofline(t_inv, 2 * H - 1, bytes_inv, flops_inv, chain, "", hbm_peak, fp32_peak, sm_mhz)
                        row = {
                            "shape_B_C_H_W_k_groups": [B, C, H, H, k, g], "variant": variant.split(" ")[0],
                            "us": {"prepare": round(t_prep, 2), "inverse": round(t_inv, 2), "bwd_input": round(t_dx, 2),
                                   "bwd_weight": round(t_dw, 2)},
                            "gpu_fwd_bwd_images_per_s": B / (t_fb * 1e-6),
                            "cpu_fwd_bwd_images_per_s": cpu_img_s, "cpu_sample_images": n_cpu,
                            "cpu_note": "float32 oracle port, %d images on %d threads, scaled to the %d host cores" % (
                                n_cpu, min(threads, n_cpu), threads),
                            "gpu_over_cpu": B / (t_fb * 1e-6) / cpu_img_s,
                            "inverse_roofline": {"bound": roof["bound"], "frac": roof["frac"],
                                                 "terms_us": {k_: round(v_, 3) for k_, v_ in roof["terms_us"].items()},
                                                 "tflops": flops_inv / t_inv / 1e6, "gbs": bytes_inv / t_inv / 1e3},
                            "round_trip_max_rel_err": rt_err, "inverse_vs_oracle_max_rel_err": inv_err,
                        }
                        f.write(json.dumps(row) + "\n")
                        f.flush()
                        print("%-26s %-7s inv %9.1f us dX %9.1f dW %9.1f | gpu %10.0f img/s cpu %8.1f (x%.0f) | %s %.3f | err %.1e %.1e" % (
                            str((B, C, H, H, k, g)), row["variant"][:7], t_inv, t_dx, t_dw, row["gpu_fwd_bwd_images_per_s"],
                            cpu_img_s, row["gpu_over_cpu"], roof["bound"][:4], roof["frac"], rt_err, inv_err), flush=True)
                        del x, grad, y, rec, dx, out
                        torch.cuda.empty_cache()
    f.close()


if __name__ == "__main__":
    main()
