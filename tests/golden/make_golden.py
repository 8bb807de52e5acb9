"""Generate golden vectors for the inverse-conv hot path FROM THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference; the GPU box has no reference):

    python tests/golden/make_golden.py

Sources of truth, all executed from the reference checkout, float64:
  * inf/utils/solve_mc.py:88-114  `solve`           (raster back-substitution)
  * inf/utils/solve_mc.py:8-50    `solve_parallel`  (anti-diagonal wavefront; H == W only)
  * inf/utils/fastflow_inverse/solve_parallel_mc.pyx:77-126, compiled by oracle/build_ref.py
  * (dX, dW): the reference has no backward on the CPU and its solver is written with
    in-place updates that torch.autograd cannot trace, so its gradient is defined here as
    the derivative of ITS OWN solver, taken without any formula of ours: the solver is
    linear in x, so dX = (L^-1)^T g with L^-1 obtained column by column by solving for
    the unit images; dW by central finite differences (float64, eps 1e-6) of
    sum(g * solve(x, W)).
  * F.conv2d with top-left zero padding for the sampling direction x = L y.
groups > 1 (cinc_kernel_level2.cu:59-72 semantics) is produced by running the reference
solver on each channel group separately.

The vectors are small (a few KB each) and are committed as tests/golden/*.npz.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

REF = os.environ.get("IFK_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

# (name, B, C, H, W, KH, KW, groups, tap scale)
CASES = [
    ("ref_test_1x4x5x5_k3_g4", 1, 4, 5, 5, 3, 3, 4, 0.01),   # tests/inf/test_layers.py:182-190
    ("ref_test_1x4x5x5_k3_g1", 1, 4, 5, 5, 3, 3, 1, 0.01),
    ("mnist_2x1x7x7_k3_g1", 2, 1, 7, 7, 3, 3, 1, 0.05),
    ("mnist_2x4x6x6_k2_g1", 2, 4, 6, 6, 2, 2, 1, 0.05),
    ("rgb_2x3x4x5_k3_g1", 2, 3, 4, 5, 3, 3, 1, 0.1),          # H != W
    ("wide_1x2x3x8_k3_g1", 1, 2, 3, 8, 3, 3, 1, 0.1),
    ("tall_1x2x8x3_k2x3_g1", 1, 2, 8, 3, 2, 3, 1, 0.1),       # KH != KW
    ("glow_2x8x4x4_k3_g4", 2, 8, 4, 4, 3, 3, 4, 0.05),        # 4 groups of 2 channels
    ("glow_1x6x5x5_k3_g2", 1, 6, 5, 5, 3, 3, 2, 0.05),
    ("big_k_1x2x6x6_k5_g1", 1, 2, 6, 6, 5, 5, 1, 0.02),
    ("single_pixel_3x5x1x1_k3_g1", 3, 5, 1, 1, 3, 3, 1, 0.2),
    ("k1_2x4x3x3_k1_g1", 2, 4, 3, 3, 1, 1, 1, 0.2),
]


def load_reference():
    if "matplotlib" not in sys.modules:     # solve_mc.py:3 imports pyplot at module top
        mpl = types.ModuleType("matplotlib")
        mpl.pyplot = types.ModuleType("matplotlib.pyplot")
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = mpl.pyplot
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "ref_solve_mc", os.path.join(REF, "inf/utils/solve_mc.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    import solve_parallel_mc as cy
    return mod, cy


def make_weight(rng, C, Cw, KH, KW, scale):
    """inv_flow_*.reset_parameters-style weight (inv_conv.py:153-170): identity at the last
    tap + small noise, then W[c, -1, -1, -1] = 1.  All C input columns are populated so the
    masked / unused entries are exercised."""
    w = rng.standard_normal((C, Cw, KH, KW)) * scale
    for c in range(min(C, Cw)):
        w[c, c, KH - 1, KW - 1] += 1.0
    w[:, -1, -1, -1] = 1.0
    return w


def per_group(fn, x, w, groups):
    C = x.shape[1]
    Cg = C // groups
    outs = []
    for g in range(groups):
        sl = slice(g * Cg, (g + 1) * Cg)
        outs.append(fn(x[:, sl], w[sl, :Cg]))
    return torch.cat(outs, dim=1)


def main():
    ref, cy = load_reference()
    for name, B, C, H, W, KH, KW, groups, scale in CASES:
        rng = np.random.default_rng(sum(ord(ch) * (i + 1) for i, ch in enumerate(name)))
        x = rng.standard_normal((B, C, H, W))
        g = rng.standard_normal((B, C, H, W))
        w = make_weight(rng, C, C, KH, KW, scale)
        Cg = C // groups

        xt = torch.tensor(x, dtype=torch.float64)
        wt = torch.tensor(w, dtype=torch.float64)

        def ref_solve(xa, wa):
            """reference solver on numpy arrays, all groups; Cython when it is valid
            (its step count needs H <= W, .pyx:95-98), else the Python `solve`."""
            outs = []
            for gi in range(groups):
                sl = slice(gi * Cg, (gi + 1) * Cg)
                if H <= W:
                    buf = np.ascontiguousarray(xa[:, sl]).copy()
                    outs.append(np.asarray(cy.solve_parallel(
                        buf, np.ascontiguousarray(wa[sl, :Cg]), (KH, KW))))
                else:
                    outs.append(ref.solve(torch.tensor(xa[:, sl]), torch.tensor(wa[sl, :Cg]),
                                          (KH, KW)).numpy())
            return np.concatenate(outs, axis=1)

        y = per_group(lambda a, b: ref.solve(a, b, (KH, KW)), xt, wt, groups)
        n = C * H * W
        Linv = ref_solve(np.eye(n).reshape(n, C, H, W), w).reshape(n, n).T   # y = Linv @ x
        dx = (g.reshape(B, n) @ Linv).reshape(B, C, H, W)                     # Linv^T g
        dw = np.zeros_like(w)
        eps = 1e-6
        for idx in np.ndindex(*w.shape):
            wp, wm = w.copy(), w.copy()
            wp[idx] += eps
            wm[idx] -= eps
            dw[idx] = np.sum(g * (ref_solve(x, wp) - ref_solve(x, wm))) / (2 * eps)
        out = {
            "x": x, "w": w, "g": g, "groups": np.int64(groups),
            "y_solve": y.numpy(),
            "dx_ref": dx,
            "dw_ref_fd": dw,
        }
        with torch.no_grad():
            if H == W:      # solve_parallel's step count assumes a square image
                out["y_solve_parallel"] = per_group(
                    lambda a, b: ref.solve_parallel(a, b, (KH, KW)), xt, wt, groups).numpy()
                ycy = []
                for gi in range(groups):
                    sl = slice(gi * Cg, (gi + 1) * Cg)
                    buf = np.ascontiguousarray(x[:, sl]).copy()       # solver works in place
                    ycy.append(np.asarray(cy.solve_parallel(
                        buf, np.ascontiguousarray(w[sl, :Cg]), (KH, KW))))
                out["y_cython"] = np.concatenate(ycy, axis=1)
            # sampling direction: x_rec = L y through torch's own convolution
            sys.path.insert(0, ROOT)
            from oracle.oracle import masked_weight
            mw = torch.tensor(masked_weight(w, groups))
            out["conv_of_y"] = F.conv2d(
                F.pad(y, (KW - 1, 0, KH - 1, 0)), mw, groups=groups).numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        rec = np.abs(out["conv_of_y"] - x).max()
        print("%-32s y %s  |conv(y)-x| %.2e" % (name, out["y_solve"].shape, rec))


if __name__ == "__main__":
    main()
