"""ad-hoc: run one forward solve eagerly per shape and report errors (debug aid for the window kernel)"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inverse_flow_b200 import _native, functional as IF
from oracle import oracle

def run(B, C, H, W, k, scale):
    rng = np.random.default_rng(1)
    x = torch.tensor(rng.standard_normal((B, C, H, W)).astype(np.float32), device="cuda")
    w = (rng.standard_normal((C, C, k, k)) * scale).astype(np.float32)
    wd = torch.tensor(w, device="cuda")
    desc = _native.describe_solve(_native.problem(B, C, H, W, k, k, C, 1))
    try:
        y = IF.inverse(x, wd, groups=1)
        torch.cuda.synchronize()
        n = min(B, 2)
        ref = oracle.inverse(x[:n].cpu().numpy().astype(np.float64), w.astype(np.float64), 1, threads=8)
        err = oracle.max_rel_err(y[:n].cpu().numpy(), ref)
        ref2 = oracle.inverse(x[-1:].cpu().numpy().astype(np.float64), w.astype(np.float64), 1, threads=8)
        err2 = oracle.max_rel_err(y[-1:].cpu().numpy(), ref2)
        print("OK  ", (B, C, H, W, k), "err %.2e %.2e" % (err, err2), desc, flush=True)
    except Exception as e:
        print("FAIL", (B, C, H, W, k), str(e).splitlines()[0], desc, flush=True)
        raise

for shape in [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]:
    B, C, H, W, k = shape
    run(B, C, H, W, k, 0.3 / (C * k * k))
