"""Phase timing of the resident and shuffle solve kernels (clock64 stamps of CTA 0, ifk_debug_set_probe).
    python tools/probe_solve.py B C H W k groups
One launch after a device sync: the phases ahead of the loop run with cold caches.
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inverse_flow_b200 import _native, functional as IF  # noqa: E402
from inverse_flow_b200.stack import reference_init_weight  # noqa: E402

NAMES = ["start -> weights/T loaded, sync", "owner bookkeeping", "loop constants", "wait for the image (TMA)",
         "pre-pass z = T x", "diagonal loop", "ybuf re-zero / store issue", "store read wait"]
# the shuffle kernel has no bookkeeping / pre-pass phases: those stamps coincide
NAMES_SHFL = ["start -> weights/T in registers, PDL wait", "-", "-", "wait for the image (TMA)", "-",
              "diagonal loop", "fence + TMA store issue", "store read wait"]


def main():
    B, C, H, W, k, g = (int(v) for v in sys.argv[1:7])
    lib = _native.load()
    lib.ifk_debug_set_probe.argtypes = [ctypes.c_void_p]
    lib.ifk_debug_set_probe.restype = None
    x = torch.randn(B, C, H, W, device="cuda")
    w = reference_init_weight(C, k).cuda()
    prep = IF.Prepared(w, g)
    out = torch.empty_like(x)
    probe = torch.zeros(16, dtype=torch.int64, device="cuda")
    for _ in range(3):
        IF.inverse(x, w, out=out, prepared=prep)
    lib.ifk_debug_set_probe(ctypes.c_void_p(probe.data_ptr()))
    IF.inverse(x, w, out=out, prepared=prep)
    torch.cuda.synchronize()
    lib.ifk_debug_set_probe(None)
    t = probe.cpu().tolist()
    print(_native.describe_solve(_native.problem(B, C, H, W, k, k, C, g)))
    # stamps: 0 start, 1 tables, 2 weights, 3 ybuf zeroed, 4 landed, 5 loop start, 6 loop end, 7 store issued, 8 end
    order = [0, 1, 2, 3, 4, 5, 6, 7, 8]
    ndiag = H + W - 1
    desc = _native.describe_solve(_native.problem(B, C, H, W, k, k, C, g))
    for a, b_, name in zip(order[:-1], order[1:], NAMES_SHFL if desc.startswith("shfl") else NAMES):
        if name != "-":
            print("%-42s %8d cycles" % (name, t[b_] - t[a]))
    print("total %d cycles; %.1f cycles per diagonal (%d diagonals)" % (t[8] - t[0], (t[6] - t[5]) / ndiag, ndiag))


if __name__ == "__main__":
    main()
