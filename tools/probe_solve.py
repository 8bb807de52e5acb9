"""Phase timing, launch time and parity of ONE solve geometry (development aid).

    python tools/probe_solve.py B C H W k groups [--reverse] [--flags F]

* clock64 stamps of CTA 0 through ifk_inverse_probe_f32 (one launch after warm-up);
* microseconds per launch: 64 back-to-back launches in a CUDA graph (programmatic dependent launch
  between them, IFK_FLAG_STABLE_PREPARED set as in a layer chain), CUDA events;
* max relative error of the solve against the float64 oracle on the first images.
"""
import argparse
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inverse_flow_b200 import _native, functional as IF  # noqa: E402
from inverse_flow_b200.stack import reference_init_weight  # noqa: E402

NAMES = ["start -> prologue done", "bookkeeping", "loop constants", "wait for the image (TMA)",
         "pre-pass / transpose", "diagonal loop", "write-out", "tail"]
# the wave kernel does everything that needs no global data ahead of the dependency wait
WAVE_NAMES = ["start -> pre-wait prologue done", "dependency wait + release", "image load issued", "wait for the image (TMA)",
              "transpose", "diagonal loop", "write-out", "tail"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("dims", type=int, nargs=6)
    ap.add_argument("--reverse", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--chain", type=int, default=0, help="also time N-layer chains (ifk_inverse_chain_f32, TL/TR/BL/BR)")
    args = ap.parse_args()
    B, C, H, W, k, g = args.dims
    lib = _native.load()
    torch.manual_seed(0)
    x = torch.randn(B, C, H, W, device="cuda")
    w = reference_init_weight(C, k).cuda()
    prep = IF.Prepared(w, g)
    out = torch.empty_like(x)
    p = _native.problem(B, C, H, W, k, k, C, g)
    ps = _native.with_flags(p, _native.FLAG_STABLE_PREPARED)
    stream = _native.current_stream(x.device)
    probe = torch.zeros(16, dtype=torch.int64, device="cuda")
    fn = lib.ifk_bwd_input_f32 if args.reverse else lib.ifk_inverse_f32
    for _ in range(3):
        _native.check(fn(ctypes.byref(p), x.data_ptr(), prep.buffer.data_ptr(), out.data_ptr(), stream))
    torch.cuda.synchronize()
    _native.check(lib.ifk_inverse_probe_f32(ctypes.byref(p), x.data_ptr(), prep.buffer.data_ptr(), out.data_ptr(),
                                            probe.data_ptr(), stream))
    torch.cuda.synchronize()
    t = probe.cpu().tolist()
    print((B, C, H, W, k, g), _native.describe_solve(p))
    ndiag = H + W - 1
    if t[8] > t[0]:
        for a, name in enumerate(WAVE_NAMES if _native.describe_solve(p).startswith("wave<") else NAMES):
            if t[a + 1] - t[a] > 0:
                print("  %-34s %8d cycles" % (name, t[a + 1] - t[a]))
        print("  total %d cycles; %.1f cycles per diagonal (%d diagonals)" % (t[8] - t[0], (t[6] - t[5]) / ndiag, ndiag))

    # time: 64 launches back to back in a graph (ping-pong buffers like a layer chain)
    a, b = x.clone(), torch.empty_like(x)
    side = torch.cuda.Stream()
    reps = 64
    with torch.cuda.stream(side):
        s2 = _native.current_stream(x.device)
        _native.check(fn(ctypes.byref(ps), a.data_ptr(), prep.buffer.data_ptr(), b.data_ptr(), s2))
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        s2 = _native.current_stream(x.device)
        for i in range(reps):
            src, dst = (a, b) if i % 2 == 0 else (b, a)
            _native.check(fn(ctypes.byref(ps), src.data_ptr(), prep.buffer.data_ptr(), dst.data_ptr(), s2))
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        a.copy_(x)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps * 1e3)
    print("  %.2f us per launch (%d chained launches in a graph)" % (best, reps))

    if args.chain:
        n = args.chain
        orients = [("TL", "TR", "BL", "BR")[i % 4] for i in range(n)]
        ws = [reference_init_weight(C, k).cuda() for _ in range(n)]
        preps = [IF.Prepared(wi, g) for wi in ws]
        outs = [torch.empty_like(x) for _ in range(n)]

        def single():
            cur = x
            for pr, o, out_i in zip(preps, orients, outs):
                q = _native.with_flags(pr.for_batch(cur, o), _native.FLAG_STABLE_PREPARED)
                _native.check(lib.ifk_inverse_f32(ctypes.byref(q), cur.data_ptr(), pr.buffer.data_ptr(), out_i.data_ptr(),
                                                  _native.current_stream(x.device)))
                cur = out_i

        def chained():
            IF.inverse_chain(x, preps, orients, outs=outs)

        res = {}
        for name, fn in (("single launches", single), ("one chained launch", chained)):
            with torch.cuda.stream(side):
                fn()
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                for _ in range(16):
                    fn()
            for _ in range(3):
                gr.replay()
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                gr.replay()
                e1.record()
                e1.synchronize()
                best = min(best, e0.elapsed_time(e1) / 16 * 1e3)
            res[name] = best
            print("  %d-layer unit, %s: %.2f us (%.2f us per layer)" % (n, name, best, best / n))

    if not args.no_parity:
        from oracle import oracle
        n = min(B, 8)
        x64 = x[:n].cpu().numpy().astype(np.float64)
        w64 = w.cpu().numpy().astype(np.float64)
        _native.check(fn(ctypes.byref(p), x.data_ptr(), prep.buffer.data_ptr(), out.data_ptr(), stream))
        torch.cuda.synchronize()
        if args.reverse:
            ref = oracle.bwd_input(x64, w64, g) if hasattr(oracle, "bwd_input") else None
        else:
            ref = oracle.inverse(x64, w64, g)
        if ref is not None:
            print("  max rel err vs float64 oracle: %.2e (first %d images), last image finite: %s" % (
                oracle.max_rel_err(out[:n].cpu().numpy(), ref), n, bool(torch.isfinite(out[-1]).all())))


if __name__ == "__main__":
    main()
