"""bench.py -- inverse-conv fwd+bwd throughput on B200 (the BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload glow_mnist|glow_cifar|glow_imagenet32|cnn_mnist] [--groups G]

A "step" is one pass of the hot path over one batch: for every inverse-conv layer of the
named if_* model (reference experiments/if_glow_mnist.py:62-124 etc.), forward
(prepare + y = L^-1 x) then backward (dX = L^-T g, dW), batch-sharded over the GPUs (weak
scaling: the per-GPU batch is the model's batch size), with ONE all-reduce of the flat dW
bucket per step when N > 1 -- the data-parallel gradient exchange of the path.

The default workload is glow_imagenet32 (the largest single-GPU configuration, the one the 1 -> 8 GPU target is
quoted on); the other workloads are reported in the same line under "workloads" (N == 1).

`value`  : images/s, inputs resident in HBM, whole forward+backward in one CUDA graph.
`e2e`    : the same step driven from pinned HOST buffers: H2D of x and g, the graph, D2H of
           y, dX and the dW bucket, synchronised, every step.
L2 is flushed (256 MiB memset) before every timed step; each step is timed with its own
CUDA event pair on the launching stream and the per-step times are summed.

`roofline`: the dominant kernel (the stage-1 wavefront solve) against all THREE terms of its bound --
HBM bytes, FP32 flops, and the (H+W-1)-step dependency chain -- with the denominators MEASURED on the box in
the same run (HBM: MEASURED_PEAKS.json; FP32 FMA rate and instruction latencies: ifk_debug_fp32_peak /
ifk_debug_latencies); `bound` / `frac` name the binding term.
`reference_cuda`: the reference's own CUDA extension (oracle/_ref, compiled from its sources) timed per
layer call at the workload's stage shapes, outside the timed region (N == 1).

--impl reference times the reference's CPU implementation of the same steps on the host
cores: the oracle port with OpenMP (the reference has no CPU backward) and,
for the forward solve alone, the reference's own compiled Cython solver (oracle/_ref).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# name -> (stages [(C, H, W, k, n_layers)], per-GPU batch, description)
WORKLOADS = {
    # experiments/if_glow_mnist.py:146-190: L=2 blocks x K=16, inv_flow_no_pad (2,2), split prior
    "glow_mnist": ([(4, 14, 14, 2, 16), (8, 7, 7, 2, 16)], 100,
                   "if_glow_mnist L=2 K=16 inverse-conv layers, batch 100/GPU"),
    # experiments/if_glow_cifar.py:48-52,82-96: 3x32x32, k=3
    "glow_cifar": ([(12, 16, 16, 3, 16), (24, 8, 8, 3, 16)], 256,
                   "if_glow_cifar 3x32x32 inverse-conv layers, batch 256/GPU"),
    # inf/if_multiGPU_imagenet32.py:196-199,290-294: 3 levels x 48 layers, k=3, batch 100/GPU
    "glow_imagenet32": ([(12, 16, 16, 3, 48), (24, 8, 8, 3, 48), (48, 4, 4, 3, 48)], 100,
                        "if_multiGPU_imagenet32 3x48 inverse-conv layers, batch 100/GPU"),
    # experiments/if_cnn_mnist.py:42-67, BASELINE configs[0]
    "cnn_mnist": ([(1, 28, 28, 3, 1), (4, 14, 14, 3, 1)], 64,
                  "if_cnn_mnist inverse-conv layers, batch 64/GPU"),
}

# DRAM bytes (read + write) of ONE launch of the dominant kernel, from a committed ncu --set full
# capture of that kernel at the workload's stage-1 shape: workload -> (bytes, where it is recorded)
PROFILED_DRAM_TRAFFIC = {
    "glow_mnist": (327168, "profiles/r01_ncu_shfl_100x4x14.txt (dram__bytes_read 327168 + dram__bytes_write 0)"),
    "glow_imagenet32": (1262592, "profiles/r02_ncu_wave_100x12x16.txt (dram__bytes_read 1262592 + dram__bytes_write 0: the "
                                 "output stays in L2 for the next layer)"),
}

def parallelism_text(n_gpus, comm_kind=None):
    return "dp%d batch-sharded, one process per GPU; dW bucket summed over ranks once per step%s" % (
        n_gpus, " (%s)" % comm_kind if comm_kind else "")


def solve_kernel_name(variant):
    """__global__ function behind an ifk_describe_solve string"""
    return {"shfl": "solve_shfl_kernel", "wave": "solve_wave_kernel", "smem": "solve_smem_kernel", "window": "solve_window_kernel",
            "stream": "solve_stream_kernel"}.get(variant.split("<")[0], "solve_global_kernel")


def chain_cycles(variant, lat, Cg, k):
    """Dependent chain of ONE wavefront step of a solve kernel, in cycles, from the instruction latencies
    measured on this box (ifk_debug_latencies): what a step cannot be shorter than however the rest of the
    work is hidden.  Returns (cycles, description)."""
    import math
    import re
    alu = lat["fadd"]
    if variant.startswith("wave"):
        m = re.search(r"cc=(\d+),ns=(\d+),vec=(\d+)", variant)
        cc, ns, vec = (int(v) for v in m.groups())
        nft = 2 if k > 1 else 0                       # taps on the previous diagonal: (0,1), (1,0)
        pf = math.ceil(nft * (Cg // vec) / ns) * vec // 2
        levels = int(math.log2(ns))
        chain = lat["sts_barsync8_lds"] + pf * lat["ffma2"] + alu + levels * (alu + lat["shfl"] + alu)
        what = ("st.shared -> bar.sync -> ld.shared %.0f + %d dependent FFMA2 x %.1f + add %.1f + %d reduce levels x "
                "(select %.1f + shuffle %.0f + add %.1f)" % (lat["sts_barsync8_lds"], pf, lat["ffma2"], alu, levels, alu,
                                                            lat["shfl"], alu))
    elif variant.startswith("shfl"):
        chain = lat["shfl"] + alu + Cg * lat["ffma"] + alu + alu
        what = "shuffle %.0f + select + %d dependent FMAs x %.1f + add + select" % (lat["shfl"], Cg, lat["ffma"])
    else:
        m = re.search(r"cc=(\d+),nv=(\d+),vec=(\d+)> ns=(\d+)", variant)
        cc, nv, vec, ns = (int(v) for v in m.groups()) if m else (1, 1, 1, 1)
        nacc = 2 if vec >= 2 else (1 if cc >= 4 else (2 if cc >= 2 else 4))
        levels = max(ns.bit_length() - 1, 0)
        multi = "threads=128(32)" not in variant
        hand = lat["sts_barsync8_lds"] if multi else lat["sts_syncwarp_lds"]
        chain = hand + lat["ffma"] * (nv * vec // nacc) + levels * (alu + lat["shfl"] + alu) + 2 * alu
        what = "store->sync->load %.0f + %d-deep FMA chain + %d shuffle levels + adds" % (hand, nv * vec // nacc, levels)
    return chain, what


def solve_roofline(kernel_us, steps, bytes_alg, flops_alg, chain, chain_what, hbm_gbs, fp32_tflops, sm_mhz):
    """three-term bound T = max(bytes / BW_hbm, flops / P_fp32, steps * chain / f_SM); all denominators measured"""
    t_hbm = bytes_alg / (hbm_gbs * 1e9) * 1e6
    t_fp32 = flops_alg / (fp32_tflops * 1e12) * 1e6
    t_wave = steps * chain / (sm_mhz * 1e6) * 1e6
    terms = {"hbm": t_hbm, "fp32": t_fp32, "wavefront-latency": t_wave}
    bound = max(terms, key=terms.get)
    out = {
        "bound": bound, "frac": terms[bound] / kernel_us,
        "terms_us": {"hbm": t_hbm, "fp32": t_fp32, "wavefront": t_wave, "measured_kernel": kernel_us},
        "hbm": {"achieved": bytes_alg / kernel_us / 1e3, "peak": hbm_gbs, "unit": "GB/s",
                "frac": bytes_alg / kernel_us / 1e3 / hbm_gbs},
        "fp32": {"achieved": flops_alg / kernel_us / 1e6, "peak": fp32_tflops, "unit": "TFLOP/s",
                 "frac": flops_alg / kernel_us / 1e6 / fp32_tflops},
        "wavefront": {"achieved": steps / kernel_us, "peak": steps / t_wave, "unit": "diagonals/us",
                      "frac": t_wave / kernel_us, "steps": steps, "chain_cycles_per_step": chain,
                      "chain": chain_what, "sm_mhz": sm_mhz},
    }
    if bound == "hbm":
        out.update(achieved=out["hbm"]["achieved"], peak=hbm_gbs, unit="GB/s")
    elif bound == "fp32":
        out.update(achieved=out["fp32"]["achieved"], peak=fp32_tflops, unit="TFLOP/s")
    else:
        out.update(achieved=out["wavefront"]["achieved"], peak=out["wavefront"]["peak"], unit="diagonals/us")
    return out


CLOCK_QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
               "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
               "clocks_event_reasons.sw_power_cap")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=" + CLOCK_QUERY, "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst)"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


# ------------------------------------------------------------------------------------------
# CPU leg: the oracle port over the same layer stack (checker + baseline; never the product)
# ------------------------------------------------------------------------------------------
def cpu_stack_step(weights, xs, gs, groups_of, threads):
    """forward+backward of the whole stack on the host.  Returns per-stage (y, dx, [dw])."""
    from oracle import oracle
    out = []
    for (ws, x, g, groups) in zip(weights, xs, gs, groups_of):
        acts = [x]
        for w in ws:
            acts.append(oracle.inverse(acts[-1], w, groups, threads))
        grad, dws = g, []
        for i in reversed(range(len(ws))):
            grad, dw = oracle.backward(grad, acts[i + 1], ws[i], groups, threads)
            dws.append(dw)
        out.append((acts[-1], grad, dws[::-1]))
    return out


def ref_cython_inverse_rate(weights, xs, budget_s=4.0, module="solve_parallel_mc"):
    """images/s of the reference's own compiled solver (oracle/_ref, float64) on the forward solves of
    the stack: `solve_parallel_mc` as the reference builds it (no -fopenmp: its prange runs on 1 core)
    or `solve_parallel_mc_omp` (same source with -fopenmp; num_threads=30 is hard-coded at .pyx:104 and
    capped by the host).  None if _ref is absent or the stack is grouped (the solver is full-C only)."""
    try:
        import importlib
        ref_dir = os.path.join(ROOT, "oracle", "_ref")
        if ref_dir not in sys.path:
            sys.path.insert(0, ref_dir)
        cy = importlib.import_module(module)
    except ImportError:
        return None
    t0 = time.perf_counter()
    n_img = 0
    # (the -fopenmp build spawns its 30-thread team once per anti-diagonal: ~2 images/s, so one image per pass)
    sample = 1 if module.endswith("_omp") else max(1, min(8, xs[0].shape[0]))
    while time.perf_counter() - t0 < budget_s:
        for ws, x in zip(weights, xs):
            k = ws[0].shape[2]
            cur = np.ascontiguousarray(x[:sample], dtype=np.float64)
            for w in ws:
                cur = np.asarray(cy.solve_parallel(cur.copy(), np.ascontiguousarray(w, dtype=np.float64), (k, k)))
        n_img += sample
    return n_img / (time.perf_counter() - t0)


def host_data(stack_desc, batch, groups, seed):
    """weights (reference init), inputs and upstream grads as numpy float32."""
    import torch
    from inverse_flow_b200.stack import reference_init_weight
    from inverse_flow_b200.functional import default_groups
    gen = torch.Generator().manual_seed(seed)
    weights, xs, gs, groups_of = [], [], [], []
    for (C, H, W, k, n) in stack_desc:
        weights.append([reference_init_weight(C, k, gen).numpy() for _ in range(n)])
        groups_of.append(default_groups(C) if groups is None else groups)
    rng = np.random.default_rng(seed)
    for (C, H, W, k, n) in stack_desc:
        xs.append(rng.standard_normal((batch, C, H, W)).astype(np.float32))
        gs.append(rng.standard_normal((batch, C, H, W)).astype(np.float32))
    return weights, xs, gs, groups_of


def time_cpu(weights, xs, gs, groups_of, threads, budget_s):
    """best-of-N wall time of one CPU step (the first pass doubles as warm-up)"""
    t0 = time.perf_counter()
    res = cpu_stack_step(weights, xs, gs, groups_of, threads)
    one = time.perf_counter() - t0
    reps = int(max(2, min(50, budget_s / max(one, 1e-6))))
    best = one
    for _ in range(reps):
        t0 = time.perf_counter()
        cpu_stack_step(weights, xs, gs, groups_of, threads)
        best = min(best, time.perf_counter() - t0)
    return best, reps, res


# ------------------------------------------------------------------------------------------
def base_line(args, stages, batch, desc, n_gpus):
    return {
        "metric": "inv-conv fwd+bwd images/s", "unit": "images/s", "n_gpus": n_gpus, "steps": args.steps,
        "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "name": args.workload, "stages_C_H_W_k_layers": stages,
                   "batch_per_gpu": batch, "global_batch": batch * n_gpus,
                   "groups": "reference(4 if C%4==0 else 1)" if args.groups is None else args.groups,
                   "weights": "inv_flow reset_parameters init (dirac + xavier_normal gain 0.01)",
                   "l2": "flushed before every timed step (256 MiB memset)"},
    }


def host_threads():
    """the cores this process may run on -- NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1, which
    would silently turn the CPU arm into a single-thread run at N > 1 (the oracle's loops take an explicit
    num_threads)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    stages, batch, desc = WORKLOADS[args.workload]
    n_gpus = args.gpus
    threads = host_threads()
    weights, xs, gs, groups_of = host_data(stages, batch, args.groups, seed=0)
    for _ in range(max(1, min(args.warmup, 2))):
        cpu_stack_step(weights, xs, gs, groups_of, threads)
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        cpu_stack_step(weights, xs, gs, groups_of, threads)
        times.append(time.perf_counter() - t0)
        if sum(times) > 180:
            break
    per = float(np.mean(times))
    value = batch / per
    line = base_line(args, stages, batch, desc, n_gpus)
    line.update({
        "impl": "reference", "value": value, "ms_per_step": per * 1e3, "steps": len(times),
        "best_ms_per_step": float(np.min(times)) * 1e3,
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": threads, "kind": "port",
                         "sample": "full step (batch %d, all layers), float32 oracle port, %d OpenMP threads = the cores "
                                   "of this process (sched_getaffinity; OMP_NUM_THREADS is ignored); the reference has "
                                   "no CPU backward" % (batch, threads),
                         "reference_cython_inverse_only_images_per_s":
                             ref_cython_inverse_rate(weights, xs) if all(g == 1 for g in groups_of) else None},
        "gpu_launches": 0,
        "parallelism": "host cores of rank 0 only (%d OpenMP threads), the same workload as one GPU's shard" % threads,
    })
    print(json.dumps(line))
    return 0


def reference_cuda_block(stages, batch, timeout_s=240):
    """The reference's own CUDA extension (oracle/_ref, compiled from its sources by oracle/build_ref_cuda.py) timed
    per layer call at the workload's stage shapes in a child process (its dw indexes out of bounds for some shapes; a
    fault must not take this process's context along).  Per-step estimate = sum over stages of layers x (inverse +
    dy + dw) -- the launch loops of inv_conv_with_bp_kernel_general.cu:72-129, 388-483, 634-735."""
    shapes = ";".join("%d,%d,%d,%d" % (batch, C, H, k) for (C, H, W, k, n) in stages if C % 4 == 0 and H == W)
    if not shapes:
        return {"unavailable": "the literal reference kernels launch nothing for C < 4"}
    try:
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ref_gpu_bench.py"), "--iters", "2",
                              "--shapes", shapes, "--ref-only"], capture_output=True, text=True, timeout=timeout_s)
    except subprocess.TimeoutExpired:
        return {"unavailable": "timed out after %d s" % timeout_s}
    recs = [json.loads(ln) for ln in out.stdout.splitlines() if ln.startswith("{")]
    if not recs or "unavailable" in recs[0]:
        return {"unavailable": (recs[0]["unavailable"] if recs else (out.stderr or "no output")[-200:])}
    by_shape = {(r["shape"][1], r["shape"][2], r["k"]): r for r in recs}
    total_us, complete, per_stage = 0.0, True, []
    for (C, H, W, k, n) in stages:
        r = by_shape.get((C, H, k))
        if r is None or r.get("ref_dw_us") is None:
            complete = False
            per_stage.append({"stage": [C, H, W, k, n], "inverse_us": r and r["ref_inverse_us"], "dy_us": r and r["ref_dy_us"],
                              "dw_us": None, "dw_error": r and r.get("ref_dw_error")})
            continue
        t = r["ref_inverse_us"] + r["ref_dy_us"] + r["ref_dw_us"]
        total_us += n * t
        per_stage.append({"stage": [C, H, W, k, n], "inverse_us": r["ref_inverse_us"], "dy_us": r["ref_dy_us"],
                          "dw_us": r["ref_dw_us"], "layer_fwd_bwd_us": t})
    return {
        "what": "reference CUDA extension inv_conv_with_bp (oracle/_ref) on this GPU, one layer call per stage shape, "
                "wall clock around synchronised calls (every launch of it is followed by cudaDeviceSynchronize), groups = "
                "its hard-coded 4, incl. the per-call host allocations of inv_conv_.backward (inv_conv.py:70-77)",
        "per_stage": per_stage,
        "step_ms_estimate": total_us * 1e-3 if complete else None,
        "images_per_s_estimate": batch / (total_us * 1e-6) if complete and total_us > 0 else None,
    }


def run_ours(args):
    import torch
    import torch.distributed as dist
    from inverse_flow_b200 import _native
    from inverse_flow_b200.stack import InvConvStack
    from oracle import oracle

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        # NCCL carries the rendezvous and the timing reductions only; its log (NCCL_DEBUG, if the launcher set it)
        # goes to stderr so that stdout stays ONE JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=device)
    n_gpus = world

    stages, batch, desc = WORKLOADS[args.workload]
    stack = InvConvStack(stages, batch, groups=args.groups, device=device, seed=0)
    weights, _, _, groups_of = host_data(stages, 1, args.groups, seed=0)      # same on every rank
    _, xs, gs, _ = host_data(stages, batch, args.groups, seed=1000 + rank)    # this rank's shard
    for st, ws in zip(stack.stages, weights):
        for wt, w in zip(st.w, ws):
            wt.copy_(torch.from_numpy(w))
    for st, x, g in zip(stack.stages, xs, gs):
        st.act[0].copy_(torch.from_numpy(x))
        st.grad_in.copy_(torch.from_numpy(g))
    stack.capture()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)

    comm_kind = None
    if world > 1:
        # the gradient exchange: the fused peer-memory all-reduce kernel inside the step's ONE graph
        # (inverse_flow_b200/parallel.py, csrc/ifk_comm.cu); --comm nccl keeps NCCL (two graphs + host-issued
        # all-reduces, the round-1 scheme) for comparison
        if args.comm == "peer":
            from inverse_flow_b200.parallel import PeerAllReduce
            stack.attach_comm(PeerAllReduce(stack.grad_bucket.numel(), device))
            stack.capture_parallel()
            comm_kind = "peer: ifk_allreduce_peer_f32 (one-shot over NVLink peer mappings, rank-ordered sum), in the step graph"
        else:
            stack.capture_bucketed()
            comm_kind = "nccl: two dist.all_reduce per step between two graph replays"

    def one_step():
        if world == 1:
            stack.step()
        elif args.comm == "peer":
            stack.step_parallel()
        else:
            stack.graph_a.replay()
            h = dist.all_reduce(stack.bucket_last, async_op=True)
            stack.graph_b.replay()
            if stack.bucket_rest.numel():
                dist.all_reduce(stack.bucket_rest)
            h.wait()

    def timed(fn, steps):
        tot = 0.0
        for _ in range(steps):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            e.synchronize()
            tot += s.elapsed_time(e)
        return tot            # ms

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput -------------------------------------------------------
    with ClockSampler(local) as clk:              # sampling spans warm-up + timed region (nvidia-smi is slow to start)
        t_up = time.perf_counter()
        for _ in range(max(args.warmup, 3)):
            flush.zero_()
            one_step()
        while rank == 0 and not clk.rows and time.perf_counter() - t_up < 3.0:
            flush.zero_()
            stack.step()                          # local work only (no collective): load until the first sample
        barrier()
        total_ms = timed(one_step, args.steps)
        barrier()
    total_ms = max_over_ranks(total_ms)
    ms_per_step = total_ms / args.steps
    value = batch * n_gpus / (ms_per_step * 1e-3)

    # ---- end to end from pinned host buffers ----------------------------------------------
    hb = stack.make_host_buffers()
    for x_h, g_h, x, g in zip(hb["x"], hb["g"], xs, gs):
        x_h.copy_(torch.from_numpy(x))
        g_h.copy_(torch.from_numpy(g))
    bytes_io = [0, 0]

    def one_step_host():
        if world > 1 and args.comm == "peer":
            # the host-to-host pipeline with the fused peer all-reduce inside: ONE graph per step, copies overlapped
            bytes_io[0], bytes_io[1] = stack.step_host(hb, parallel=True)
        elif world > 1:
            h2d = d2h = 0
            for st, x, g in zip(stack.stages, hb["x"], hb["g"]):
                st.act[0].copy_(x, non_blocking=True)
                st.grad_in.copy_(g, non_blocking=True)
                h2d += 2 * x.numel() * 4
            one_step()
            for st, y, dx in zip(stack.stages, hb["y"], hb["dx"]):
                y.copy_(st.act[st.n], non_blocking=True)
                dx.copy_(st.dx, non_blocking=True)
                d2h += 2 * y.numel() * 4
            hb["dw"].copy_(stack.grad_bucket, non_blocking=True)
            d2h += hb["dw"].numel() * 4
            torch.cuda.current_stream().synchronize()
            bytes_io[0], bytes_io[1] = h2d, d2h
        else:
            bytes_io[0], bytes_io[1] = stack.step_host(hb)

    for _ in range(max(args.warmup, 3)):
        one_step_host()
    barrier()
    e2e_ms = max_over_ranks(timed(one_step_host, args.steps)) / args.steps
    barrier()
    e2e_value = batch * n_gpus / (e2e_ms * 1e-3)

    # ---- dominant kernel: the wavefront solve at the first stage's shape --------------------
    st0 = stack.stages[0]
    import ctypes
    lib = stack.lib
    reps = 64
    ps0 = ctypes.byref(st0.problem_stable)
    ping, pong = st0.act[0].clone(), torch.empty_like(st0.act[0])

    def chain_of_solves():
        s0 = _native.current_stream(device)
        for i in range(reps):                 # ping-pong like consecutive layers, programmatic dependent launch between them
            src, dst = (ping, pong) if i % 2 == 0 else (pong, ping)
            _native.check(lib.ifk_inverse_f32(ps0, src.data_ptr(), st0.prepared[0].data_ptr(), dst.data_ptr(), s0))

    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        chain_of_solves()
    torch.cuda.synchronize()
    solve_graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(solve_graph):
        chain_of_solves()
    ping.copy_(st0.act[0])
    solve_us = min(timed(solve_graph.replay, 1) for _ in range(5)) / reps * 1e3
    N0 = batch * st0.C * st0.H * st0.W
    Cg0 = st0.C // st0.groups
    K0 = st0.k * st0.k
    solve_bytes = 4 * (2 * N0 + st0.C * Cg0 * K0)
    solve_flops = 2 * N0 * (Cg0 * K0 - 1)
    variant = _native.describe_solve(st0.problem)
    hbm_peak, peak_src = measured_peak_gbs()
    hw = _native.hw_microbench(device)            # FP32 FMA rate and instruction latencies of THIS GPU, measured now
    fp32_peak = max(hw["fp32_tflops_ffma"], hw["fp32_tflops_ffma2"])
    clocks = clk.summary()
    sm_mhz = clocks["sm_mhz"] or clocks["sm_max_mhz"] or 1965.0
    chain, chain_what = chain_cycles(variant, hw["latency_cycles"], Cg0, st0.k)
    steps0 = st0.H + st0.W - 1
    roofline = solve_roofline(solve_us, steps0, solve_bytes, solve_flops, chain, chain_what, hbm_peak, fp32_peak, sm_mhz)
    n_solves = sum(2 * s_.n for s_ in stack.stages)
    traffic = PROFILED_DRAM_TRAFFIC.get(args.workload, (None, None)) if args.groups == 1 else (None, None)
    roofline.update({
        "traffic": traffic[0], "traffic_source": traffic[1],
        "peak_source": {"hbm": peak_src,
                        "fp32": "measured now: ifk_debug_fp32_peak (FFMA %.1f, FFMA2 %.1f TFLOP/s, all SMs)" % (
                            hw["fp32_tflops_ffma"], hw["fp32_tflops_ffma2"]),
                        "latencies": "measured now: ifk_debug_latencies (cycles per dependent op) %s" % json.dumps(
                            {k_: round(v_, 1) for k_, v_ in hw["latency_cycles"].items()})},
        "kernel": "%s (wavefront triangular solve; inverse and bwd_input)" % solve_kernel_name(variant),
        "variant": variant, "kernel_us": solve_us,
        "algorithmic_bytes_per_launch": solve_bytes, "algorithmic_flops_per_launch": solve_flops,
        "how": "%d chained launches of ifk_inverse_f32 at stage 1 (ping-pong buffers, programmatic dependent launch) in a "
               "CUDA graph, CUDA events, L2 flushed before each replay, best of 5" % reps,
        "solve_launches_per_step": n_solves,
        "note": "T = max(bytes/BW_hbm, flops/P_fp32, (H+W-1) * chain / f_SM): at model shapes the image (%.0f KB) moves "
                "in under a microsecond and the binding term is the dependency chain of the wavefront" % (solve_bytes / 1e3),
    })

    line = base_line(args, stages, batch, desc, n_gpus)
    line.update({
        "value": value, "ms_per_step": ms_per_step,
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": bytes_io[0] * n_gpus,
                "d2h_bytes_per_step": bytes_io[1] * n_gpus, "ms_per_step": e2e_ms},
        "gpu_launches": stack.launches_per_step * args.steps,
        "gpu_launches_per_step": stack.launches_per_step,
        "clocks": clocks,
        "roofline": roofline,
        "hw_microbench": hw,
        "parallelism": parallelism_text(n_gpus, comm_kind),
        "step_algorithmic": {"GBps": stack.algorithmic_bytes_per_step() * n_gpus / (ms_per_step * 1e-3) / 1e9,
                             "GFLOPs": stack.algorithmic_flops_per_step() * n_gpus / (ms_per_step * 1e-3) / 1e9,
                             "wavefront_steps": stack.wavefront_steps_per_step()},
    })

    # ---- CPU baseline + parity (rank 0, N == 1 only) ----------------------------------------
    if world == 1 and not args.no_cpu:
        threads = host_threads()
        per, reps_cpu, res = time_cpu(weights, xs, gs, groups_of, threads, budget_s=12.0)
        stack.step()
        torch.cuda.synchronize()
        errs = []
        for st, (y, dx, dws) in zip(stack.stages, res):
            errs.append(oracle.max_rel_err(st.act[st.n].cpu().numpy(), y))
            errs.append(oracle.max_rel_err(st.dx.cpu().numpy(), dx))
            errs.append(max(oracle.max_rel_err(a.cpu().numpy(), b) for a, b in zip(st.dw, dws)))
        line["cpu_baseline"] = {
            "value": batch / per, "unit": "images/s", "cores": threads, "kind": "port",
            "sample": "best of %d full steps (batch %d, all layers) of the float32 oracle port, OpenMP over the batch on "
                      "the %d cores of this process; the reference has no CPU backward" % (reps_cpu, batch, threads),
            "ms_per_step": per * 1e3,
            "reference_cython_inverse_only_images_per_s":
                ref_cython_inverse_rate(weights, xs) if all(g == 1 for g in groups_of) else None,
        }
        line["parity"] = {"max_rel_err_vs_cpu_port_f32": float(max(errs)), "tolerance": 1e-5,
                          "metric": "max|a - ref| / max|ref| (max-norm)",
                          "what": "final y, dX and every dW of the chained stack"}
        line["reference_cuda"] = reference_cuda_block(stages, batch)
        # the other workloads, device-resident, same timing method (fewer steps)
        line["workloads"] = {args.workload: {"value": value, "ms_per_step": ms_per_step}}
        del stack
        torch.cuda.empty_cache()
        for name in sorted(WORKLOADS):
            if name == args.workload:
                continue
            st_w, b_w, d_w = WORKLOADS[name]
            other = InvConvStack(st_w, b_w, groups=args.groups, device=device, seed=0)
            for st in other.stages:
                st.act[0].normal_()
                st.grad_in.normal_()
            other.capture()
            for _ in range(3):
                other.step()
            n_w = max(5, args.steps // 2)
            ms_w = timed(other.step, n_w) / n_w
            line["workloads"][name] = {"value": b_w / (ms_w * 1e-3), "ms_per_step": ms_w, "workload": d_w,
                                       "solve_variants": [_native.describe_solve(st.problem).split(" ")[0] for st in other.stages]}
            del other
            torch.cuda.empty_cache()
    if world > 1:
        # every rank must hold bit-identical gradient sums (the peer kernel adds in rank order)
        one_step()
        torch.cuda.synchronize()
        lo, hi = stack.grad_bucket.clone(), stack.grad_bucket.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        line["comm"] = {"kind": comm_kind, "bucket_floats": stack.grad_bucket.numel(),
                        "bit_identical_across_ranks": bool(torch.equal(lo, hi)),
                        "finite": bool(torch.isfinite(stack.grad_bucket).all())}
        if args.comm == "peer":
            # ... and they must be the sum: NCCL's all-reduce of the same per-rank buckets as the checker
            ref = stack.comm.bucket.clone()
            dist.all_reduce(ref)
            den = float(ref.abs().max()) or 1.0
            line["comm"]["max_rel_err_vs_nccl_allreduce"] = float((stack.grad_bucket - ref).abs().max()) / den
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="glow_imagenet32", choices=sorted(WORKLOADS))
    ap.add_argument("--comm", default="peer", choices=["peer", "nccl"],
                    help="N > 1: the fused peer-memory all-reduce kernel inside the step graph (default) or NCCL")
    ap.add_argument("--no-cpu", action="store_true",
                    help="skip the cpu_baseline / parity leg (for runs under a profiler; the line is then incomplete)")
    ap.add_argument("--groups", type=int, default=1,
                    help="channel groups; 1 = full coupling (the reference CPU solver), 0 = reference kernels' "
                         "rule (4 when C %% 4 == 0)")
    args = ap.parse_args()
    if args.groups == 0:
        args.groups = None
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
