"""Run one op a few times at one shape (to be wrapped by ncu).

    python tools/profile_one.py B C H W k groups [op] [reps]
op in {inverse, bwd_input, bwd_weight, conv, prepare, all}
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inverse_flow_b200 import functional as IF  # noqa: E402
from inverse_flow_b200.stack import reference_init_weight  # noqa: E402


def main():
    B, C, H, W, k, g = (int(v) for v in sys.argv[1:7])
    op = sys.argv[7] if len(sys.argv) > 7 else "all"
    reps = int(sys.argv[8]) if len(sys.argv) > 8 else 4
    torch.manual_seed(0)
    x = torch.randn(B, C, H, W, device="cuda")
    grad = torch.randn(B, C, H, W, device="cuda")
    w = reference_init_weight(C, k).cuda()
    prep = IF.Prepared(w, g)
    y = IF.inverse(x, w, prepared=prep)
    dx = IF.bwd_input(grad, w, prepared=prep)
    out = torch.empty_like(x)
    dw = torch.empty_like(w)
    torch.cuda.synchronize()
    for _ in range(reps):
        if op in ("prepare", "all"):
            IF.Prepared(w, g)
        if op in ("inverse", "all"):
            IF.inverse(x, w, out=out, prepared=prep)
        if op in ("bwd_input", "all"):
            IF.bwd_input(grad, w, out=out, prepared=prep)
        if op in ("bwd_weight", "all"):
            IF.bwd_weight(dx, y, w, groups=g, out=dw)
        if op in ("conv", "all"):
            IF.conv(y, w, groups=g, out=out)
    torch.cuda.synchronize()
    print("ok", float(out.abs().sum()))


if __name__ == "__main__":
    main()
