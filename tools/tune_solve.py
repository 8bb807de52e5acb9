"""Sweep forced kernel configurations of the resident solve for one shape (tuning aid).

    python tools/tune_solve.py B C H W k groups

Uses the IFK_SOLVE_CFG=cc,nv,vec,ns,nslots override; prints microseconds per launch
(20 launches per CUDA graph) for every configuration the kernel accepts.
"""
import itertools
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inverse_flow_b200 import _native, functional as IF  # noqa: E402
from inverse_flow_b200.stack import reference_init_weight  # noqa: E402
from tools.microbench import time_op  # noqa: E402


def main():
    B, C, H, W, k, g = (int(v) for v in sys.argv[1:7])
    x = torch.randn(B, C, H, W, device="cuda")
    w = reference_init_weight(C, k).cuda()
    prep = IF.Prepared(w, g)
    out = torch.empty_like(x)
    os.environ.pop("IFK_SOLVE_CFG", None)
    ref = IF.inverse(x, w, prepared=prep).clone()
    base = time_op(lambda: IF.inverse(x, w, out=out, prepared=prep), 5)
    prob = _native.problem(B, C, H, W, k, k, C, g)
    print("default %.2f us  %s" % (base, _native.describe_solve(prob)))
    Cg = C // g
    res = []
    for vec, cc, nv, ns in itertools.product((4, 2, 1), (1, 2, 3, 4, 6, 8, 12), (1, 2, 3, 4, 5, 6, 8, 9, 10, 12, 16, 24),
                                             (1, 2, 4, 8, 16, 32)):
        if cc > Cg:
            continue
        cgv = (Cg + vec - 1) // vec
        nvt = (k * k - 1) * cgv
        if ns * nv < nvt or (ns > 1 and (ns // 2) * nv >= nvt) or (nv > 1 and ns * (nv - 1) >= nvt + ns):
            continue
        for slots in sorted({H, (H + 1) // 2, (H + 3) // 4}):
            os.environ["IFK_SOLVE_CFG"] = "%d,%d,%d,%d,%d" % (cc, nv, vec, ns, slots)
            d = _native.describe_solve(prob)
            tag = "cc=%d,nv=%d,vec=%d> ns=%d " % (cc, nv, vec, ns)
            if tag not in d or ("slots=%d " % slots) not in d:
                continue
            try:
                t = time_op(lambda: IF.inverse(x, w, out=out, prepared=prep), 3)
            except Exception as e:  # noqa: BLE001
                print("fail", d, e)
                continue
            err = float((out - ref).abs().max())
            res.append((t, d, err))
    os.environ.pop("IFK_SOLVE_CFG", None)
    for t, d, err in sorted(res)[:12]:
        print("%8.2f us  %s  maxdiff=%.1e" % (t, d, err))
    print("... %d configs; worst %.2f us" % (len(res), max(r[0] for r in res) if res else 0))


if __name__ == "__main__":
    main()
