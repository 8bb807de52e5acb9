"""One small call through every kernel family (to be wrapped by compute-sanitizer; prints max errors vs the oracle).

    compute-sanitizer --tool memcheck python tools/sanity_run.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inverse_flow_b200 import _native, functional as IF  # noqa: E402
from oracle import oracle  # noqa: E402


def case(B, C, H, W, k, g, knobs=None, fused=False, chain=0):
    for key in ("IFK_SOLVE_SPLIT", "IFK_SOLVE_WAVE", "IFK_SOLVE_SHFL", "IFK_SOLVE_WINDOW", "IFK_SOLVE_STREAM", "IFK_SOLVE_GLOBAL"):
        os.environ.pop(key, None)
    os.environ.update(knobs or {})
    _native.reload_env()
    rng = np.random.default_rng(5)
    x = rng.standard_normal((B, C, H, W)).astype(np.float32)
    gr = rng.standard_normal((B, C, H, W)).astype(np.float32)
    w = (rng.standard_normal((C, C, k, k)) * 0.02).astype(np.float32)
    for c in range(C):
        w[c, c, -1, -1] += 1.0
    xd, gd, wd = (torch.tensor(a, device="cuda") for a in (x, gr, w))
    y = IF.inverse(xd, wd, groups=g)
    dx, dw = IF.backward(gd, y, wd, groups=g)
    rec = IF.conv(y, wd, groups=g)
    extra = ""
    if fused:
        s = torch.rand(C, device="cuda") + 0.5
        b = torch.randn(C, device="cuda")
        xin = torch.randn(B, C // 4, 2 * H, 2 * W, device="cuda")
        yf = IF.inverse_fused(xin, wd, in_scale=s, in_bias=b, squeeze=True, groups=g)
        _, dz = IF.bwd_input_fused(gd, wd, out_scale=s, squeeze=True, groups=g)
        extra = " fused ok %s %s" % (tuple(yf.shape), tuple(dz.shape))
    if chain:
        preps = [IF.Prepared(wd, g) for _ in range(chain)]
        ys = IF.inverse_chain(xd, preps, ["TL", "TR", "BL", "BR"][:chain])
        extra += " chain ok %d" % len(ys)
    torch.cuda.synchronize()
    y_ref = oracle.inverse(x.astype(np.float64), w.astype(np.float64), g)
    dx_ref, dw_ref = oracle.backward(gr.astype(np.float64), y_ref, w.astype(np.float64), g)
    print("%-22s %-40s y %.1e dx %.1e dw %.1e rt %.1e%s" % (
        (B, C, H, W, k, g), _native.describe_solve(_native.problem(B, C, H, W, k, k, C, g)).split(" ")[0],
        oracle.max_rel_err(y.cpu().numpy(), y_ref), oracle.max_rel_err(dx.cpu().numpy(), dx_ref),
        oracle.max_rel_err(dw.cpu().numpy(), dw_ref), float((rec - xd).abs().max()), extra), flush=True)


if __name__ == "__main__":
    case(5, 12, 16, 16, 3, 1, fused=True, chain=4)                       # wave kernel, fused neighbours, chain
    case(3, 24, 8, 8, 3, 1, fused=True)
    case(3, 48, 4, 4, 3, 1)
    case(5, 12, 16, 16, 3, 1, {"IFK_SOLVE_SPLIT": "1"}, chain=2)          # split kernel
    case(3, 24, 8, 8, 3, 1, {"IFK_SOLVE_SPLIT": "1"})
    case(4, 4, 14, 14, 2, 1)                                              # shuffle kernel
    case(3, 12, 16, 16, 5, 1)                                             # resident kernel
    case(2, 12, 16, 16, 3, 1, {"IFK_SOLVE_WINDOW": "1"})                  # window kernel
    case(2, 12, 16, 16, 3, 1, {"IFK_SOLVE_STREAM": "1"})                  # stream kernel
    case(2, 5, 9, 13, 3, 1, {"IFK_SOLVE_GLOBAL": "1"})                    # fallback
    print("sanity ok")
