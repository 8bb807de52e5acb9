"""bench.py -- inverse-conv fwd+bwd throughput on B200 (the BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload glow_mnist|glow_cifar|glow_imagenet32|cnn_mnist] [--groups G]

A "step" is one pass of the hot path over one batch: for every inverse-conv layer of the
named if_* model (reference experiments/if_glow_mnist.py:62-124 etc.), forward
(prepare + y = L^-1 x) then backward (dX = L^-T g, dW), batch-sharded over the GPUs (weak
scaling: the per-GPU batch is the model's batch size), with ONE all-reduce of the flat dW
bucket per step when N > 1 -- the data-parallel gradient exchange of the path.

`value`  : images/s, inputs resident in HBM, whole forward+backward in one CUDA graph.
`e2e`    : the same step driven from pinned HOST buffers: H2D of x and g, the graph, D2H of
           y, dX and the dW bucket, synchronised, every step.
L2 is flushed (256 MiB memset) before every timed step; each step is timed with its own
CUDA event pair on the launching stream and the per-step times are summed.

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
}

def solve_kernel_name(variant):
    """__global__ function behind an ifk_describe_solve string"""
    return {"shfl": "solve_shfl_kernel", "smem": "solve_smem_kernel", "window": "solve_window_kernel",
            "stream": "solve_stream_kernel"}.get(variant.split("<")[0], "solve_global_kernel")


def wavefront_step_accounting(lib, p0, st0, device):
    """The latency roofline of the solve: cycles one anti-diagonal costs inside the kernel's loop
    (clock64 stamps of CTA 0 around the loop, one extra probed launch) against the dependent chain
    of one step built from measured instruction latencies (B300_MICROARCH.md: ld.shared 29, FMA 4,
    named barrier ~47; warp shuffle ~24)."""
    import ctypes
    import re
    import torch
    from inverse_flow_b200 import _native
    lib.ifk_debug_set_probe.argtypes = [ctypes.c_void_p]
    lib.ifk_debug_set_probe.restype = None
    probe = torch.zeros(16, dtype=torch.int64, device=device)
    lib.ifk_debug_set_probe(ctypes.c_void_p(probe.data_ptr()))
    try:
        _native.check(lib.ifk_inverse_f32(p0, st0.act[0].data_ptr(), st0.prepared[0].data_ptr(),
                                          st0.act[1].data_ptr(), _native.current_stream(device)))
        torch.cuda.synchronize()
    finally:
        lib.ifk_debug_set_probe(None)
    t = probe.cpu().tolist()
    steps = st0.H + st0.W - 1
    variant = _native.describe_solve(st0.problem)
    if t[6] <= t[5] or t[8] <= t[0]:
        return {"steps": steps, "note": "this kernel variant carries no probe stamps"}
    loop = (t[6] - t[5]) / steps
    Cg = st0.C // st0.groups
    if variant.startswith("shfl"):
        chain = 24 + 4 + 4 * Cg + 4 + 4
        what = "shuffle 24 + select 4 + %d dependent FMAs (the taps fed by the fresh values) + add 4 + select 4" % Cg
    else:
        m = re.search(r"cc=(\d+),nv=(\d+),vec=(\d+)> ns=(\d+)", variant)
        cc, nv, vec, ns = (int(v) for v in m.groups()) if m else (1, 1, 1, 1)
        nacc = 2 if vec >= 2 else (1 if cc >= 4 else (2 if cc >= 2 else 4))
        levels = max(ns.bit_length() - 1, 0)
        multi = "threads=128(32)" not in variant
        chain = 29 + 4 * (nv * vec // nacc) + 24 * levels + 8 + (47 if multi else 10)
        what = ("ld.shared 29 + %d-deep FMA chain + %d shuffle levels x 24 + adds 8 + %s" %
                (nv * vec // nacc, levels, "named barrier 47" if multi else "syncwarp 10"))
    return {"steps": steps, "loop_cycles_per_step": loop, "floor_cycles_per_step": chain, "frac": chain / loop,
            "loop_share_of_kernel": (t[6] - t[5]) / (t[8] - t[0]), "kernel_cycles": t[8] - t[0],
            "floor": what,
            "how": "clock64 stamps of CTA 0 around the diagonal loop (ifk_debug_set_probe), one probed launch"}


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
    t0 = time.perf_counter()
    res = cpu_stack_step(weights, xs, gs, groups_of, threads)
    one = time.perf_counter() - t0
    reps = int(max(1, min(50, budget_s / max(one, 1e-6))))
    t0 = time.perf_counter()
    for _ in range(reps):
        cpu_stack_step(weights, xs, gs, groups_of, threads)
    per = (time.perf_counter() - t0) / reps
    return per, reps, res


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
                   "l2": "flushed before every timed step (256 MiB memset)",
                   "parallelism": "dp%d batch-sharded; dW bucket all-reduced per step (last stage overlapped with the "
                                  "remaining backward)" % n_gpus},
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    stages, batch, desc = WORKLOADS[args.workload]
    n_gpus = args.gpus
    from oracle import oracle
    threads = oracle.max_threads()
    weights, xs, gs, groups_of = host_data(stages, batch, args.groups, seed=0)
    for _ in range(max(1, min(args.warmup, 2))):
        cpu_stack_step(weights, xs, gs, groups_of, threads)
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        cpu_stack_step(weights, xs, gs, groups_of, threads)
        times.append(time.perf_counter() - t0)
        if sum(times) > 240:
            break
    per = float(np.mean(times))
    value = batch / per
    line = base_line(args, stages, batch, desc, n_gpus)
    line.update({
        "impl": "reference", "value": value, "ms_per_step": per * 1e3, "steps": len(times),
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": threads, "kind": "port",
                         "sample": "full step (batch %d, all layers), float32 oracle port with OpenMP; the "
                                   "reference has no CPU backward" % batch,
                         "reference_cython_inverse_only_images_per_s":
                             ref_cython_inverse_rate(weights, xs) if all(g == 1 for g in groups_of) else None,
                         "reference_cython_openmp_inverse_only_images_per_s":
                             ref_cython_inverse_rate(weights, xs, module="solve_parallel_mc_omp")
                             if all(g == 1 for g in groups_of) else None},
        "gpu_launches": 0,
    })
    line["config"]["parallelism"] = "host cores of rank 0 only (%d OpenMP threads)" % threads
    print(json.dumps(line))
    return 0


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
        # keep stdout to ONE JSON line: NCCL prints its version banner to stdout at any debug level
        if "IFK_NCCL_DEBUG" in os.environ:
            os.environ["NCCL_DEBUG"] = os.environ["IFK_NCCL_DEBUG"]
        else:
            os.environ.pop("NCCL_DEBUG", None)
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

    if world > 1:
        stack.capture_bucketed()

    def one_step():
        if world == 1:
            stack.step()
            return
        # bucketed gradient exchange: the last stage's dW is all-reduced while the remaining
        # backward (graph B) runs; the rest follows.  NCCL over NVLink, nothing else is exchanged.
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
        if world > 1:
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
    solve_graph = torch.cuda.CUDAGraph()
    p0 = ctypes.byref(st0.problem)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        _native.check(lib.ifk_inverse_f32(p0, st0.act[0].data_ptr(), st0.prepared[0].data_ptr(),
                                          st0.act[1].data_ptr(), _native.current_stream(device)))
    torch.cuda.synchronize()
    with torch.cuda.graph(solve_graph):
        for i in range(reps):
            _native.check(lib.ifk_inverse_f32(p0, st0.act[0].data_ptr(), st0.prepared[0].data_ptr(),
                                              st0.act[1].data_ptr(), _native.current_stream(device)))
    solve_ms = timed(solve_graph.replay, 5) / 5 / reps
    N0 = batch * st0.C * st0.H * st0.W
    Cg0 = st0.C // st0.groups
    solve_bytes = 4 * (2 * N0 + st0.C * Cg0 * st0.k * st0.k)
    peak, peak_src = measured_peak_gbs()
    achieved = solve_bytes / (solve_ms * 1e-3) / 1e9
    n_solves = sum(2 * s.n for s in stack.stages)
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": PROFILED_DRAM_TRAFFIC.get(args.workload, (None, None))[0] if args.groups == 1 else None,
        "traffic_source": PROFILED_DRAM_TRAFFIC.get(args.workload, (None, None))[1] if args.groups == 1 else None,
        "peak_source": peak_src,
        "kernel": "%s (wavefront triangular solve; inverse and bwd_input)" % solve_kernel_name(
            _native.describe_solve(st0.problem)),
        "variant": _native.describe_solve(st0.problem),
        "kernel_us": solve_ms * 1e3, "algorithmic_bytes_per_launch": solve_bytes,
        "how": "%d back-to-back launches of ifk_inverse_f32 at stage 1 in a CUDA graph, CUDA events, L2 "
               "flushed before each replay" % reps,
        "wavefront_steps_per_launch": st0.H + st0.W - 1,
        "wavefront_ns_per_diagonal": solve_ms * 1e6 / (st0.H + st0.W - 1),
        "solve_launches_per_step": n_solves,
        "note": "latency-bound at model shapes: the image (%.0f KB) moves in well under a microsecond; the "
                "binding term is the (H+W-1)-step dependency chain (see DESIGN.md roofline)" % (solve_bytes / 1e3),
    }

    roofline["wavefront"] = wavefront_step_accounting(lib, p0, st0, device)

    line = base_line(args, stages, batch, desc, n_gpus)
    line.update({
        "value": value, "ms_per_step": ms_per_step,
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": bytes_io[0] * n_gpus,
                "d2h_bytes_per_step": bytes_io[1] * n_gpus, "ms_per_step": e2e_ms},
        "gpu_launches": stack.launches_per_step * args.steps,
        "gpu_launches_per_step": stack.launches_per_step,
        "clocks": clk.summary(),
        "roofline": roofline,
        "step_algorithmic": {"GBps": stack.algorithmic_bytes_per_step() * n_gpus / (ms_per_step * 1e-3) / 1e9,
                             "GFLOPs": stack.algorithmic_flops_per_step() * n_gpus / (ms_per_step * 1e-3) / 1e9,
                             "wavefront_steps": stack.wavefront_steps_per_step()},
    })

    # ---- CPU baseline + parity (rank 0, N == 1 only) ----------------------------------------
    if world == 1 and not args.no_cpu:
        threads = oracle.max_threads()
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
            "sample": "%d full steps (batch %d, all layers) of the float32 oracle port, OpenMP over the batch; "
                      "the reference has no CPU backward" % (reps_cpu, batch),
            "ms_per_step": per * 1e3,
            "reference_cython_inverse_only_images_per_s":
                ref_cython_inverse_rate(weights, xs) if all(g == 1 for g in groups_of) else None,
            "reference_cython_openmp_inverse_only_images_per_s":
                ref_cython_inverse_rate(weights, xs, module="solve_parallel_mc_omp")
                if all(g == 1 for g in groups_of) else None,
        }
        line["parity"] = {"max_rel_err_vs_cpu_port_f32": float(max(errs)), "tolerance": 1e-5,
                          "what": "final y, dX and every dW of the chained stack"}
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
    ap.add_argument("--workload", default="glow_mnist", choices=sorted(WORKLOADS))
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
