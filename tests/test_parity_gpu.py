"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle.

Bar (BASELINE.json north_star): float32 results within max relative error 1e-5 of the
reference semantics, where max relative error := max|a - ref| / max|ref| with ref the
float64 oracle (oracle.max_rel_err) -- a MAX-NORM metric: one denominator per tensor.  Beside it
every case asserts an ELEMENTWISE relative error < 1e-3 on the entries with |ref| >= 1e-2 max|ref|
(oracle.max_elem_rel_err; float32 rounding of the largest terms bounds what smaller entries can reach).
conv(inverse(x)) reconstruction error reported too.
Weights follow the reference initialisation (small taps); a few cases use larger taps.
"""
import os
import re

import numpy as np
import pytest
import torch

from conftest import make_weight
from oracle import oracle

pytestmark = pytest.mark.gpu

TOL = 1e-5
ELEM_FLOOR, ELEM_TOL = 1e-2, 1e-3      # elementwise relative error on entries with |ref| >= 1e-2 max|ref|


@pytest.fixture(scope="module")
def IF():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from inverse_flow_b200 import functional
    return functional


def dev(a):
    return torch.tensor(np.ascontiguousarray(a), dtype=torch.float32, device="cuda")


def run_all(IF, x, w, g, groups, orient=0):
    """inverse, conv, dX, dW on the GPU; float64 oracle beside it.  Returns error dict."""
    xd, wd, gd = dev(x), dev(w), dev(g)
    y = IF.inverse(xd, wd, groups=groups, orient=orient)
    rec = IF.conv(y, wd, groups=groups, orient=orient)
    dx, dw = IF.backward(gd, y, wd, groups=groups, orient=orient)
    torch.cuda.synchronize()
    x64, w64, g64 = (np.asarray(a, dtype=np.float64) for a in (x.astype(np.float32), w.astype(np.float32),
                                                                g.astype(np.float32)))
    y_ref = oracle.inverse(x64, w64, groups, threads=4, orient=orient)
    dx_ref, dw_ref = oracle.backward(g64, y_ref, w64, groups, threads=4, orient=orient)
    return {
        "y": oracle.max_rel_err(y.cpu().numpy(), y_ref),
        "rec": oracle.max_rel_err(rec.cpu().numpy(), x64),
        "conv": oracle.max_rel_err(rec.cpu().numpy(),
                                   oracle.conv(y.cpu().numpy().astype(np.float64), w64, groups, orient=orient)),
        "dx": oracle.max_rel_err(dx.cpu().numpy(), dx_ref),
        "dw": oracle.max_rel_err(dw.cpu().numpy(), dw_ref),
        "dw_masked_zero": bool(np.all(dw.cpu().numpy()[dw_ref == 0.0] == 0.0)),
        "elem": max(oracle.max_elem_rel_err(y.cpu().numpy(), y_ref, ELEM_FLOOR),
                    oracle.max_elem_rel_err(dx.cpu().numpy(), dx_ref, ELEM_FLOOR),
                    oracle.max_elem_rel_err(dw.cpu().numpy(), dw_ref, ELEM_FLOOR)),
    }




def assert_parity(err, tol=TOL):
    for k in ("y", "rec", "conv", "dx", "dw"):
        assert err[k] < tol, (k, err)
    assert err["elem"] < ELEM_TOL, err
    assert err["dw_masked_zero"], err


def test_golden_vectors(IF, golden):
    """outputs of the reference itself (tests/golden/make_golden.py)."""
    x, w, g, groups = golden["x"], golden["w"], golden["g"], golden["groups"]
    xd, wd, gd = dev(x), dev(w), dev(g)
    y = IF.inverse(xd, wd, groups=groups)
    assert oracle.max_rel_err(y.cpu().numpy(), golden["y_solve"]) < TOL
    assert oracle.max_rel_err(IF.conv(dev(golden["y_solve"]), wd, groups=groups).cpu().numpy(),
                              golden["conv_of_y"]) < TOL
    dx, dw = IF.backward(gd, dev(golden["y_solve"]), wd, groups=groups)
    assert oracle.max_rel_err(dx.cpu().numpy(), golden["dx_ref"]) < TOL
    assert oracle.max_rel_err(dw.cpu().numpy(), golden["dw_ref_fd"]) < TOL
    # separate entry points agree with the fused one bit for bit
    dx2 = IF.bwd_input(gd, wd, groups=groups)
    dw2 = IF.bwd_weight(dx2, dev(golden["y_solve"]), wd, groups=groups)
    assert torch.equal(dx, dx2) and torch.equal(dw, dw2)


# (B, C, H, W, KH, KW, groups, tap scale): BASELINE model shapes, the reference test shape,
# ragged / degenerate cases
SHAPES = [
    (64, 1, 28, 28, 3, 3, 1, 0.05),       # configs[0]: if_cnn_mnist
    (100, 4, 14, 14, 2, 2, 1, 0.05),      # if_glow_mnist stage 1
    (100, 4, 14, 14, 2, 2, 4, 0.05),
    (100, 8, 7, 7, 2, 2, 1, 0.05),        # if_glow_mnist stage 2
    (100, 8, 7, 7, 2, 2, 4, 0.05),
    (32, 12, 16, 16, 3, 3, 1, 0.02),      # cifar / imagenet32 level 1
    (32, 12, 16, 16, 3, 3, 4, 0.02),
    (32, 24, 8, 8, 3, 3, 1, 0.02),
    (32, 24, 8, 8, 3, 3, 4, 0.02),
    (32, 48, 4, 4, 3, 3, 1, 0.02),
    (32, 48, 4, 4, 3, 3, 4, 0.02),
    (1, 4, 5, 5, 3, 3, 4, 0.01),          # tests/inf/test_layers.py:182-190
    (3, 3, 16, 16, 5, 5, 1, 0.02),
    (2, 3, 9, 13, 3, 3, 1, 0.1),          # H != W
    (2, 5, 13, 9, 2, 3, 1, 0.1),          # KH != KW, odd channel count
    (2, 6, 10, 10, 3, 3, 2, 0.05),
    (2, 6, 10, 10, 3, 3, 3, 0.05),
    (4, 7, 1, 1, 3, 3, 1, 0.2),           # single pixel: only the centre tap acts
    (2, 4, 1, 17, 3, 3, 1, 0.1),          # single row
    (2, 4, 17, 1, 3, 3, 1, 0.1),          # single column
    (2, 4, 6, 6, 1, 1, 1, 0.2),           # 1x1 kernel: pure channel triangular solve
    (2, 2, 12, 12, 7, 7, 1, 0.01),        # kernel larger than half the image
    (1, 1, 3, 3, 5, 5, 1, 0.05),          # kernel larger than the image
    (5, 16, 32, 32, 3, 3, 1, 0.01),
]


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s[:7])))
def test_parity_against_oracle(IF, shape):
    B, C, H, W, KH, KW, groups, scale = shape
    rng = np.random.default_rng(hash(shape[:7]) % (2 ** 31))
    x = rng.standard_normal((B, C, H, W)).astype(np.float32)
    g = rng.standard_normal((B, C, H, W)).astype(np.float32)
    w = make_weight(rng, C, C, KH, KW, scale)
    assert_parity(run_all(IF, x, w, g, groups))


@pytest.mark.parametrize("shape", [(3, 12, 16, 16, 3, 3, 1, 0.02), (2, 8, 7, 7, 2, 2, 4, 0.05),
                                   (2, 3, 9, 13, 3, 3, 1, 0.1), (2, 5, 13, 9, 2, 3, 1, 0.1)],
                         ids=lambda s: "x".join(map(str, s[:7])))
def test_global_fallback_kernel(IF, shape, monkeypatch):
    """Force the kernel used for images that do not fit in shared memory."""
    monkeypatch.setenv("IFK_SOLVE_GLOBAL", "1")
    B, C, H, W, KH, KW, groups, scale = shape
    rng = np.random.default_rng(11)
    x = rng.standard_normal((B, C, H, W)).astype(np.float32)
    g = rng.standard_normal((B, C, H, W)).astype(np.float32)
    w = make_weight(rng, C, C, KH, KW, scale)
    assert_parity(run_all(IF, x, w, g, groups))


@pytest.mark.parametrize("shape", [(3, 12, 16, 16, 3, 3, 1, 0.02), (2, 8, 7, 7, 2, 2, 4, 0.05),
                                   (2, 3, 9, 13, 3, 3, 1, 0.1), (2, 5, 13, 9, 2, 3, 1, 0.1),
                                   (2, 24, 8, 8, 3, 3, 1, 0.02), (2, 4, 6, 6, 5, 5, 1, 0.02)],
                         ids=lambda s: "x".join(map(str, s[:7])))
def test_stream_kernel(IF, shape, monkeypatch):
    """Force the kernel for images that exceed shared memory (reads neighbours back from the output)."""
    monkeypatch.setenv("IFK_SOLVE_STREAM", "1")
    from inverse_flow_b200 import _native
    B, C, H, W, KH, KW, groups, scale = shape
    assert _native.describe_solve(_native.problem(B, C, H, W, KH, KW, C, groups)).startswith("stream<")
    rng = np.random.default_rng(12)
    x = rng.standard_normal((B, C, H, W)).astype(np.float32)
    g = rng.standard_normal((B, C, H, W)).astype(np.float32)
    w = make_weight(rng, C, C, KH, KW, scale)
    assert_parity(run_all(IF, x, w, g, groups))


ORIENT_SHAPES = [(3, 12, 16, 16, 3, 3, 1, 0.02), (2, 8, 7, 7, 2, 2, 4, 0.05), (2, 3, 9, 13, 3, 3, 1, 0.1),
                 (2, 5, 13, 9, 2, 3, 1, 0.1), (4, 1, 28, 28, 3, 3, 1, 0.1), (2, 4, 6, 6, 5, 5, 1, 0.02),
                 (3, 20, 10, 10, 3, 3, 1, 0.02)]


@pytest.mark.parametrize("kernel", ["default", "split", "wave", "resident", "window", "stream", "global"])
@pytest.mark.parametrize("orient", ["TR", "BL", "BR"])
@pytest.mark.parametrize("shape", ORIENT_SHAPES, ids=lambda s: "x".join(map(str, s[:7])))
def test_orientations_by_index_reflection(IF, shape, orient, kernel, monkeypatch):
    """ifk_problem.orient: every operator in the frame reflected along the order's axes, against
    the oracle applied to explicitly flipped copies (what the reference layers do with torch.flip,
    inf/layers/inv_conv.py:198-214) -- for each of the three solve kernels."""
    if kernel == "stream":
        if shape[1] // shape[6] == 1:
            pytest.skip("single-channel groups have no stream variant")
        monkeypatch.setenv("IFK_SOLVE_STREAM", "1")
    elif kernel == "window":
        monkeypatch.setenv("IFK_SOLVE_WINDOW", "1")
    elif kernel == "resident":
        monkeypatch.setenv("IFK_SOLVE_SHFL", "0")       # the shared-memory kernel also where the shuffle kernel applies
        monkeypatch.setenv("IFK_SOLVE_WAVE", "0")       # ... and where the pipelined wavefront kernel does
    elif kernel == "wave":
        from inverse_flow_b200 import _native
        if not _native.describe_solve(_native.problem(*shape[:6], shape[1], shape[6])).startswith("wave<"):
            pytest.skip("no pipelined wavefront variant for this group width / kernel size")
    elif kernel == "split":
        from inverse_flow_b200 import _native
        monkeypatch.setenv("IFK_SOLVE_SPLIT", "1")      # opt-in kernel
        if not _native.describe_solve(_native.problem(*shape[:6], shape[1], shape[6])).startswith("split<"):
            pytest.skip("no warp-specialised variant for this group width / kernel size / image height")
    elif kernel == "global":
        monkeypatch.setenv("IFK_SOLVE_GLOBAL", "1")
    B, C, H, W, KH, KW, groups, scale = shape
    rng = np.random.default_rng(31)
    x = rng.standard_normal((B, C, H, W)).astype(np.float32)
    g = rng.standard_normal((B, C, H, W)).astype(np.float32)
    w = make_weight(rng, C, C, KH, KW, scale)
    assert_parity(run_all(IF, x, w, g, groups, orient=orient))


# (B, C, H, W, KH, KW, groups, tap scale) x kernel family "cc,ns,vec" (None = the default family)
WAVE_CASES = [
    ((100, 12, 16, 16, 3, 3, 1, 0.02), None), ((100, 12, 16, 16, 3, 3, 1, 0.02), "6,8,2"),      # imagenet32 level 1
    ((256, 12, 16, 16, 3, 3, 1, 0.02), None), ((256, 12, 16, 16, 3, 3, 1, 0.02), "6,8,2"),      # cifar batch
    ((100, 24, 8, 8, 3, 3, 1, 0.02), None), ((100, 24, 8, 8, 3, 3, 1, 0.02), "6,16,2"),
    ((256, 24, 8, 8, 3, 3, 1, 0.02), None),
    ((100, 48, 4, 4, 3, 3, 1, 0.02), None), ((100, 48, 4, 4, 3, 3, 1, 0.02), "6,32,2"),
    ((100, 24, 8, 8, 3, 3, 4, 0.02), None), ((100, 48, 4, 4, 3, 3, 4, 0.02), None),               # reference's 4 groups
    ((5, 12, 5, 7, 3, 3, 1, 0.05), None), ((5, 12, 7, 5, 3, 3, 1, 0.05), "6,8,2"),               # ragged
    ((3, 12, 32, 32, 3, 3, 1, 0.01), None), ((3, 24, 16, 16, 3, 3, 1, 0.01), None),              # several rows per thread
    ((2, 12, 20, 9, 3, 3, 1, 0.02), None), ((2, 48, 3, 4, 3, 3, 1, 0.02), None), ((7, 24, 5, 11, 3, 3, 1, 0.02), None),
    ((300, 12, 6, 6, 3, 3, 1, 0.02), None), ((301, 48, 4, 4, 3, 3, 4, 0.02), None),              # stripes of images per CTA
    ((3, 12, 2, 2, 3, 3, 1, 0.05), None), ((2, 6, 9, 9, 3, 3, 1, 0.05), None),
]


@pytest.mark.parametrize("case", WAVE_CASES, ids=lambda c: "x".join(map(str, c[0][:7])) + ("-" + c[1] if c[1] else ""))
def test_wave_kernel(IF, case, monkeypatch):
    """the software-pipelined wavefront kernel (ifk_solve_wave.cu): every compiled family, BASELINE batches,
    ragged images, several rows per thread, batches beyond the grid, all orientations, with and without TMA"""
    from inverse_flow_b200 import _native
    shape, family = case
    if family:
        monkeypatch.setenv("IFK_WAVE_CFG", family)
    B, C, H, W, KH, KW, groups, scale = shape
    d = _native.describe_solve(_native.problem(B, C, H, W, KH, KW, C, groups))
    assert d.startswith("wave<"), d
    if family:
        cc, ns, vec = family.split(",")
        assert "cc=%s,ns=%s,vec=%s" % (cc, ns, vec) in d, d
    rng = np.random.default_rng(29)
    x = rng.standard_normal((B, C, H, W)).astype(np.float32)
    g = rng.standard_normal((B, C, H, W)).astype(np.float32)
    w = make_weight(rng, C, C, KH, KW, scale)
    assert_parity(run_all(IF, x, w, g, groups))
    for orient in ("TR", "BL", "BR"):
        assert_parity(run_all(IF, x[:5], w, g[:5], groups, orient=orient))
    monkeypatch.setenv("IFK_SOLVE_NOBULK", "1")
    assert_parity(run_all(IF, x[:3], w, g[:3], groups))


# (B, C, H, W, KH, KW, groups, tap scale) x reduction split of the helper warps (None = the default family)
SPLIT_CASES = [
    ((100, 12, 16, 16, 3, 3, 1, 0.02), None), ((100, 12, 16, 16, 3, 3, 1, 0.02), "8"),          # imagenet32 level 1
    ((256, 12, 16, 16, 3, 3, 1, 0.02), None), ((256, 12, 16, 16, 3, 3, 1, 0.02), "8"),          # cifar batch
    ((100, 24, 8, 8, 3, 3, 1, 0.02), None), ((256, 24, 8, 8, 3, 3, 1, 0.02), None),
    ((100, 48, 4, 4, 3, 3, 4, 0.02), None), ((301, 48, 4, 4, 3, 3, 4, 0.02), None),             # reference's 4 groups: Cg = 12
    ((5, 12, 5, 7, 3, 3, 1, 0.05), None), ((5, 12, 7, 5, 3, 3, 1, 0.05), "8"),                  # ragged
    ((2, 12, 16, 9, 3, 3, 1, 0.02), None), ((7, 24, 5, 11, 3, 3, 1, 0.02), None), ((2, 24, 8, 30, 3, 3, 1, 0.02), None),
    ((300, 12, 6, 6, 3, 3, 1, 0.02), None),                                                    # stripes of images per CTA
    ((3, 12, 2, 2, 3, 3, 1, 0.05), None), ((3, 12, 1, 9, 3, 3, 1, 0.05), None), ((3, 12, 16, 40, 3, 3, 1, 0.02), None),
]


@pytest.mark.parametrize("case", SPLIT_CASES, ids=lambda c: "x".join(map(str, c[0][:7])) + ("-" + c[1] if c[1] else ""))
def test_split_kernel(IF, case, monkeypatch):
    """the warp-specialised wavefront kernel (ifk_solve_split.cu): every compiled family, BASELINE batches,
    ragged images, batches beyond the grid, all orientations, with and without TMA"""
    from inverse_flow_b200 import _native
    shape, family = case
    monkeypatch.setenv("IFK_SOLVE_SPLIT", "1")          # opt-in kernel
    if family:
        monkeypatch.setenv("IFK_SPLIT_CFG", family)
    B, C, H, W, KH, KW, groups, scale = shape
    d = _native.describe_solve(_native.problem(B, C, H, W, KH, KW, C, groups))
    assert d.startswith("split<"), d
    if family:
        assert "ns=%s," % family in d, d
    rng = np.random.default_rng(37)
    x = rng.standard_normal((B, C, H, W)).astype(np.float32)
    g = rng.standard_normal((B, C, H, W)).astype(np.float32)
    w = make_weight(rng, C, C, KH, KW, scale)
    assert_parity(run_all(IF, x, w, g, groups))
    for orient in ("TR", "BL", "BR"):
        assert_parity(run_all(IF, x[:5], w, g[:5], groups, orient=orient))
    monkeypatch.setenv("IFK_SOLVE_NOBULK", "1")
    assert_parity(run_all(IF, x[:3], w, g[:3], groups))


SHFL_SHAPES = [(100, 4, 14, 14, 2, 2, 1, 0.05), (100, 8, 7, 7, 2, 2, 1, 0.05), (64, 1, 28, 28, 3, 3, 1, 0.1),
               (9, 4, 14, 14, 3, 3, 1, 0.03), (5, 4, 14, 9, 2, 2, 1, 0.05), (5, 4, 9, 14, 3, 3, 1, 0.03),
               (7, 8, 7, 7, 2, 2, 4, 0.1), (7, 12, 16, 16, 3, 3, 4, 0.03), (6, 8, 5, 6, 3, 3, 4, 0.02),
               (3, 2, 32, 5, 2, 2, 1, 0.1), (3, 1, 1, 7, 3, 3, 1, 0.1), (3, 3, 6, 1, 3, 3, 1, 0.1),
               (1500, 4, 6, 6, 2, 2, 1, 0.05), (4, 3, 10, 12, 3, 3, 1, 0.02), (300, 16, 7, 7, 2, 2, 4, 0.05)]


@pytest.mark.parametrize("nct", [0, 1])
@pytest.mark.parametrize("shape", SHFL_SHAPES, ids=lambda s: "x".join(map(str, s[:7])))
def test_shuffle_kernel(IF, shape, nct, monkeypatch):
    """the register/shuffle wavefront kernel (rows on the lanes of one warp, neighbours by warp
    shuffle), with the widest channel split that fits the warp (default) and with one lane per row"""
    from inverse_flow_b200 import _native
    if nct:
        monkeypatch.setenv("IFK_SHFL_NCT", str(nct))
    B, C, H, W, KH, KW, groups, scale = shape
    d = _native.describe_solve(_native.problem(B, C, H, W, KH, KW, C, groups))
    if nct and not d.startswith("shfl<"):
        pytest.skip("no one-lane-per-row variant for this group width")
    assert d.startswith("shfl<"), d
    rng = np.random.default_rng(23)
    x = rng.standard_normal((B, C, H, W)).astype(np.float32)
    g = rng.standard_normal((B, C, H, W)).astype(np.float32)
    w = make_weight(rng, C, C, KH, KW, scale)
    assert_parity(run_all(IF, x, w, g, groups))
    for orient in ("TR", "BL"):
        assert_parity(run_all(IF, x[:4], w, g[:4], groups, orient=orient))


@pytest.mark.parametrize("shape", [(3, 20, 10, 10, 3, 3, 1, 0.02), (2, 48, 9, 13, 3, 3, 1, 0.01), (2, 24, 8, 8, 2, 2, 1, 0.05),
                                   (2, 32, 6, 7, 5, 5, 1, 0.005), (1, 17, 5, 3, 7, 7, 1, 0.005), (2, 80, 5, 6, 3, 3, 4, 0.02),
                                   (200, 24, 8, 8, 3, 3, 1, 0.02)],
                         ids=lambda s: "x".join(map(str, s[:7])))
def test_wide_group_conv_kernel(IF, shape, monkeypatch):
    """the register-tiled masked convolution for wide groups (4 pixels x 4 channels per thread, sliding
    row window), pinned on small shapes, in every orientation, against the oracle and the plain kernel"""
    B, C, H, W, KH, KW, groups, scale = shape
    rng = np.random.default_rng(29)
    y = rng.standard_normal((B, C, H, W)).astype(np.float32)
    w = make_weight(rng, C, C, KH, KW, scale)
    yd, wd = dev(y), dev(w)
    for orient in ("TL", "TR", "BL", "BR"):
        monkeypatch.setenv("IFK_CONV_WIDE", "1")
        a = IF.conv(yd, wd, groups=groups, orient=orient)
        monkeypatch.setenv("IFK_CONV_WIDE", "0")
        b = IF.conv(yd, wd, groups=groups, orient=orient)
        torch.cuda.synchronize()
        ref = oracle.conv(y.astype(np.float64), w.astype(np.float64), groups, threads=4, orient=orient)
        assert oracle.max_rel_err(a.cpu().numpy(), ref) < TOL, orient
        assert oracle.max_rel_err(b.cpu().numpy(), ref) < TOL, orient


def test_bad_orient_is_rejected(IF):
    x = torch.randn(1, 4, 5, 5, device="cuda")
    w = torch.zeros(4, 4, 3, 3, device="cuda")
    with pytest.raises(ValueError):
        IF.inverse(x, w, groups=1, orient=7)
    with pytest.raises(ValueError):
        IF.conv(x, w, groups=1, orient="XX")


def test_layer_orders_match_autograd_through_flips():
    """order='TR'/'BL'/'BR' forward, dX and dW equal the top-left op wrapped in torch.flip"""
    from inverse_flow_b200.layers import inv_flow_with_pad
    from inverse_flow_b200.layers.inv_conv import inv_conv_4d
    flips = {"TR": [3], "BL": [2], "BR": [2, 3]}
    for order, dims in flips.items():
        torch.manual_seed(5)
        m = inv_flow_with_pad(8, 8, (3, 3), order=order, groups=1).to("cuda")
        x = torch.randn(3, 8, 9, 7, device="cuda", requires_grad=True)
        g = torch.randn(3, 8, 9, 7, device="cuda")
        y, _ = m(x)
        y.backward(g)
        dx, dw = x.grad.clone(), m.weight_fwd.grad.clone()
        x2 = x.detach().clone().requires_grad_(True)
        w2 = m.weight_fwd.detach().clone().requires_grad_(True)
        y2 = torch.flip(inv_conv_4d(torch.flip(x2, dims), w2, 1), dims)
        y2.backward(g)
        for a, b in ((y, y2), (dx, x2.grad), (dw, w2.grad)):
            assert oracle.max_rel_err(a.detach().cpu().numpy(), b.detach().cpu().numpy().astype(np.float64)) < TOL
        rec = m.reverse(y.detach())
        assert oracle.max_rel_err(rec.cpu().numpy(), x.detach().cpu().numpy().astype(np.float64)) < TOL


def test_large_images_take_the_window_path(IF):
    """(2,12,64,64) k=3 and (2,48,32,32) k=3 exceed shared memory: window kernel; dW staged or not."""
    from inverse_flow_b200 import _native
    for (B, C, H, W, k, scale) in [(2, 12, 64, 64, 3, 0.01), (2, 48, 32, 32, 3, 0.005), (3, 3, 128, 96, 3, 0.05),
                                   (2, 1, 200, 260, 3, 0.05), (3, 12, 64, 48, 5, 0.003)]:
        assert _native.describe_solve(_native.problem(B, C, H, W, k, k, C, 1)).startswith("window<")
        rng = np.random.default_rng(6)
        x = rng.standard_normal((B, C, H, W)).astype(np.float32)
        g = rng.standard_normal((B, C, H, W)).astype(np.float32)
        w = make_weight(rng, C, C, k, k, scale)
        assert_parity(run_all(IF, x, w, g, 1))


@pytest.mark.parametrize("kernel", ["window", "stream"])
def test_wide_group_runs_as_a_thread_block_cluster(IF, kernel, monkeypatch):
    """Cg = 96, k = 3: 73.7K prepared weights exceed one SM's register file, so the large-image
    kernels run as a cluster of CTAs that split the output channels: the window kernel exchanges y
    through distributed shared memory, the stream kernel through L2; one cluster barrier per diagonal."""
    from inverse_flow_b200 import _native
    if kernel == "stream":
        monkeypatch.setenv("IFK_SOLVE_STREAM", "1")
    B, C, H, W, k = 2, 96, 8, 8, 3
    d = _native.describe_solve(_native.problem(B, C, H, W, k, k, C, 1))
    assert d.startswith(kernel + "<") and int(re.search(r"cluster=(\d+)", d).group(1)) >= 4
    rng = np.random.default_rng(7)
    x = rng.standard_normal((B, C, H, W)).astype(np.float32)
    g = rng.standard_normal((B, C, H, W)).astype(np.float32)
    w = make_weight(rng, C, C, k, k, 0.003)
    assert_parity(run_all(IF, x, w, g, 1))


@pytest.mark.parametrize("shape", [(3, 48, 16, 16, 5, 5, 0.002, 2), (2, 48, 12, 10, 7, 7, 0.001, 8),
                                   (45, 48, 6, 5, 7, 7, 0.001, 8),      # 18 clusters: every one solves several images
                                   (160, 48, 8, 8, 5, 5, 0.002, 2),     # small clusters, several images each
                                   (2, 96, 16, 16, 5, 5, 0.001, 16), (5, 96, 16, 12, 3, 3, 0.003, 4)],
                         ids=lambda s: "x".join(map(str, s[:6])))
def test_wide_groups_with_large_kernels_run_in_clusters(IF, shape):
    """C >= 48 with k >= 5: the weights only fit the register files of 2..16 SMs together"""
    from inverse_flow_b200 import _native
    B, C, H, W, KH, KW, scale, min_cluster = shape
    d = _native.describe_solve(_native.problem(B, C, H, W, KH, KW, C, 1))
    assert d.startswith("window<") and int(re.search(r"cluster=(\d+)", d).group(1)) >= min_cluster, d
    rng = np.random.default_rng(17)
    x = rng.standard_normal((B, C, H, W)).astype(np.float32)
    g = rng.standard_normal((B, C, H, W)).astype(np.float32)
    w = make_weight(rng, C, C, KH, KW, scale)
    assert_parity(run_all(IF, x, w, g, 1))


def test_large_image_takes_the_global_path(IF):
    """(2, 96, 20, 20) k=7: neither shared memory nor 16 register files hold it -> plain fallback."""
    from inverse_flow_b200 import _native
    rng = np.random.default_rng(5)
    B, C, H, W, k = 2, 96, 20, 20, 7
    assert _native.describe_solve(_native.problem(B, C, H, W, k, k, C, 1)).startswith("global")
    x = rng.standard_normal((B, C, H, W)).astype(np.float32)
    g = rng.standard_normal((B, C, H, W)).astype(np.float32)
    w = make_weight(rng, C, C, k, k, 0.0002)
    assert_parity(run_all(IF, x, w, g, 1))


@pytest.mark.parametrize("shape", [(3, 12, 16, 16, 3, 3, 1, 0.02), (2, 8, 7, 7, 2, 2, 4, 0.05),
                                   (2, 3, 9, 13, 3, 3, 1, 0.1), (2, 5, 13, 9, 2, 3, 1, 0.1),
                                   (2, 24, 8, 8, 3, 3, 1, 0.02), (2, 4, 6, 6, 5, 5, 1, 0.02),
                                   (4, 1, 28, 28, 3, 3, 1, 0.1), (3, 20, 10, 10, 3, 3, 1, 0.02),
                                   (2, 2, 12, 12, 7, 7, 1, 0.01), (1, 6, 1, 9, 3, 3, 1, 0.1), (1, 6, 9, 1, 3, 3, 1, 0.1),
                                   (150, 12, 20, 20, 3, 3, 1, 0.02), (700, 4, 9, 7, 3, 3, 1, 0.05),
                                   (1300, 6, 5, 8, 2, 3, 2, 0.05)],
                         ids=lambda s: "x".join(map(str, s[:7])))
def test_window_kernel(IF, shape, monkeypatch):
    """Force the ring-of-diagonals kernel on small shapes (several images per CTA for the last ones)."""
    monkeypatch.setenv("IFK_SOLVE_WINDOW", "1")
    from inverse_flow_b200 import _native
    B, C, H, W, KH, KW, groups, scale = shape
    assert _native.describe_solve(_native.problem(B, C, H, W, KH, KW, C, groups)).startswith("window<")
    rng = np.random.default_rng(13)
    x = rng.standard_normal((B, C, H, W)).astype(np.float32)
    g = rng.standard_normal((B, C, H, W)).astype(np.float32)
    w = make_weight(rng, C, C, KH, KW, scale)
    assert_parity(run_all(IF, x, w, g, groups))


@pytest.mark.parametrize("shape", [(1, 1, 3, 3, 1), (3, 3, 2, 2, 1), (4, 4, 2, 2, 1), (8, 8, 2, 2, 1), (5, 7, 3, 3, 1), (12, 12, 3, 3, 1), (13, 13, 3, 2, 1),
                                   (24, 24, 3, 3, 4), (48, 48, 3, 3, 1), (48, 48, 5, 5, 4), (96, 96, 3, 3, 1), (20, 20, 1, 3, 2)],
                         ids=lambda s: "x".join(map(str, s)))
def test_prepared_rows_match_the_float64_fold_for_every_slab_size(IF, shape, monkeypatch):
    """ifk_prepare.cu: the canonical prepared rows (T kernel + register-tiled tap products) for every slab size of
    the tap launch (IFK_PREP_CFG=0,n), against T = (I + A0)^-1, -T W_q formed in float64 from the same weights"""
    C, Cw, KH, KW, groups = shape
    rng = np.random.default_rng(41)
    w = make_weight(rng, C, Cw, KH, KW, 0.05)
    wt = dev(w)
    Cg, K = C // groups, KH * KW
    KD = K * Cg
    KDP = (KD + 3) // 4 * 4

    def canonical(cfg):
        if cfg:
            monkeypatch.setenv("IFK_PREP_CFG", cfg)
        buf = IF.Prepared(wt, groups).buffer
        torch.cuda.synchronize()
        return buf[:2 * C * KDP].cpu().numpy().reshape(2, C, KDP)[:, :, :KD].reshape(2, groups, Cg, K, Cg)

    tiled = canonical(None)
    for cfg in ("0,1", "0,2", "0,%d" % K):
        assert np.array_equal(canonical(cfg), tiled), cfg           # the slab size only changes which CTA forms a tap
    w64 = w.astype(np.float64)
    for G in range(groups):
        blk = w64[G * Cg:(G + 1) * Cg, :Cg]                      # rows of the group, its Cg input columns
        A0 = np.tril(blk[:, :, KH - 1, KW - 1], -1)
        T = np.linalg.inv(np.eye(Cg) + A0)
        for t in range(K):
            Wq = blk[:, :, KH - 1 - t // KW, KW - 1 - t % KW]
            fwd = T if t == 0 else -(T @ Wq)
            adj = T.T if t == 0 else -(T.T @ Wq.T)
            for d, ref in ((0, fwd), (1, adj)):
                err = np.abs(tiled[d, G, :, t, :] - ref).max() / max(np.abs(ref).max(), 1e-30)
                assert err < 1e-5, (G, t, d, err)


def test_weight_with_fewer_input_columns(IF):
    """groups=4 only reads W[:, :C/4]; a (C, C/4, k, k) weight must give the same result."""
    rng = np.random.default_rng(2)
    x = rng.standard_normal((4, 8, 6, 6)).astype(np.float32)
    w = make_weight(rng, 8, 8, 3, 3, 0.05)
    a = IF.inverse(dev(x), dev(w), groups=4)
    b = IF.inverse(dev(x), dev(w[:, :2]), groups=4)
    assert torch.equal(a, b)
    _, dwa = IF.backward(dev(x), a, dev(w), groups=4)
    _, dwb = IF.backward(dev(x), a, dev(w[:, :2]), groups=4)
    assert torch.equal(dwa[:, :2], dwb) and torch.all(dwa[:, 2:] == 0)


def test_masked_weight_entries_are_ignored(IF):
    """the diagonal tap and the centre-tap upper triangle never influence the result
    (solve_mc.py:104-108; the CUDA kernels skip k_c == c, .cu:58-60)."""
    rng = np.random.default_rng(3)
    x = rng.standard_normal((3, 6, 8, 8)).astype(np.float32)
    w = make_weight(rng, 6, 6, 3, 3, 0.05)
    w2 = w.copy()
    for c in range(6):
        w2[c, c:, -1, -1] = rng.standard_normal(6 - c) * 100
    assert torch.equal(IF.inverse(dev(x), dev(w)), IF.inverse(dev(x), dev(w2)))
    assert torch.equal(IF.conv(dev(x), dev(w)), IF.conv(dev(x), dev(w2)))


def test_empty_batch(IF):
    w = dev(make_weight(np.random.default_rng(0), 4, 4, 3, 3))
    x = torch.zeros((0, 4, 5, 5), device="cuda")
    assert IF.inverse(x, w).shape == (0, 4, 5, 5)
    assert IF.conv(x, w).shape == (0, 4, 5, 5)
    dx, dw = IF.backward(x, x, w)
    assert dx.shape == (0, 4, 5, 5) and dw.shape == w.shape and torch.all(dw == 0)


def test_linearity_and_determinism_at_full_batch(IF):
    """size-independent properties at a BASELINE-sized batch: L^-1 is linear, dW is
    bit-reproducible (fixed-order reduction), conv(inverse(x)) == x."""
    torch.manual_seed(0)
    B, C, H, W, k = 512, 12, 16, 16, 3
    w = dev(make_weight(np.random.default_rng(1), C, C, k, k, 0.02))
    x1 = torch.randn(B, C, H, W, device="cuda")
    x2 = torch.randn(B, C, H, W, device="cuda")
    y1, y2 = IF.inverse(x1, w, groups=1), IF.inverse(x2, w, groups=1)
    y12 = IF.inverse(x1 + 2.0 * x2, w, groups=1)
    scale = float(y12.abs().max())
    assert float((y12 - (y1 + 2.0 * y2)).abs().max()) / scale < 1e-5
    assert float((IF.conv(y1, w, groups=1) - x1).abs().max()) / float(x1.abs().max()) < 1e-5
    g = torch.randn(B, C, H, W, device="cuda")
    dx_a, dw_a = IF.backward(g, y1, w, groups=1)
    dx_b, dw_b = IF.backward(g, y1, w, groups=1)
    assert torch.equal(dx_a, dx_b) and torch.equal(dw_a, dw_b)
    # adjoint identity <g, L^-1 x> == <L^-T g, x>
    lhs = float((g.double() * y1.double()).sum())
    rhs = float((dx_a.double() * x1.double()).sum())
    assert abs(lhs - rhs) / max(abs(lhs), 1.0) < 1e-5


@pytest.mark.parametrize("shape", [(512, 3, 64, 64, 3, 1), (256, 12, 32, 32, 3, 1), (64, 48, 16, 16, 3, 4),
                                   (512, 1, 28, 28, 3, 1)], ids=str)
def test_round_trip_at_sweep_sizes(IF, shape):
    """encode -> decode round trip at microbenchmark-sweep sizes (too big for the CPU oracle)."""
    B, C, H, W, k, groups = shape
    torch.manual_seed(1)
    w = dev(make_weight(np.random.default_rng(4), C, C, k, k, 0.01))
    x = torch.randn(B, C, H, W, device="cuda")
    y = IF.inverse(x, w, groups=groups)
    assert float((IF.conv(y, w, groups=groups) - x).abs().max()) / float(x.abs().max()) < 1e-5
    # a slice of the batch against the oracle
    y_ref = oracle.inverse(x[:2].cpu().numpy().astype(np.float64), w.cpu().numpy().astype(np.float64), groups)
    assert oracle.max_rel_err(y[:2].cpu().numpy(), y_ref) < TOL


def test_autograd_function_matches_oracle(IF):
    from inverse_flow_b200.layers import inv_conv_4d
    rng = np.random.default_rng(8)
    x = rng.standard_normal((4, 8, 7, 7)).astype(np.float32)
    g = rng.standard_normal((4, 8, 7, 7)).astype(np.float32)
    w = make_weight(rng, 8, 8, 3, 3, 0.05)
    for groups in (1, 4):
        xt = dev(x).requires_grad_(True)
        wt = dev(w).requires_grad_(True)
        y = inv_conv_4d(xt, wt, groups)
        y.backward(dev(g))
        y_ref = oracle.inverse(x.astype(np.float64), w.astype(np.float64), groups)
        dx_ref, dw_ref = oracle.backward(g.astype(np.float64), y_ref, w.astype(np.float64), groups)
        assert oracle.max_rel_err(y.detach().cpu().numpy(), y_ref) < TOL
        assert oracle.max_rel_err(xt.grad.cpu().numpy(), dx_ref) < TOL
        assert oracle.max_rel_err(wt.grad.cpu().numpy(), dw_ref) < TOL


def test_layers_round_trip_like_the_reference_test():
    """tests/inf/test_layers.py:19-36, 182-190: reverse(forward(x)) == x, atol 1e-3."""
    from inverse_flow_b200.layers import Inv_FlowUnit, inv_flow_no_pad, inv_flow_with_pad
    torch.manual_seed(0)
    input_size = (1, 4, 5, 5)
    mods = [inv_flow_no_pad(4, 4, (3, 3)), inv_flow_with_pad(4, 4, (3, 3), order='TL')]
    mods += [inv_flow_with_pad(8, 8, (3, 3), order=o, groups=1) for o in ('TR', 'BL', 'BR')]
    mods += [Inv_FlowUnit(8, 8, (3, 3)), inv_flow_no_pad(12, 12, (2, 2), groups=1)]
    for m in mods:
        m = m.to('cuda')
        C = m.conv_tl.in_channels if isinstance(m, Inv_FlowUnit) else m.in_channels
        x = torch.randn(3, C, *input_size[2:], device='cuda')
        out = m(x)
        fwd, logdet = out
        assert logdet == 0.0
        rev = m.reverse(fwd.detach())
        np.testing.assert_allclose(rev.cpu().numpy(), x.cpu().numpy(), atol=1e-3)


def test_layer_order_matches_flipped_oracle():
    from inverse_flow_b200.layers import inv_flow_with_pad
    torch.manual_seed(0)
    m = inv_flow_with_pad(4, 4, (3, 3), order='BR', groups=1).to('cuda')
    x = torch.randn(2, 4, 6, 6, device='cuda')
    y, _ = m(x)
    w = m.weight_fwd.detach().cpu().numpy().astype(np.float64)
    xf = x.cpu().numpy().astype(np.float64)[:, :, ::-1, ::-1]
    y_ref = oracle.inverse(np.ascontiguousarray(xf), w, 1)[:, :, ::-1, ::-1]
    assert oracle.max_rel_err(y.detach().cpu().numpy(), y_ref) < TOL


def test_reference_module_call_style():
    """inv_conv.py:48-56, 262-267: caller-allocated zero output, result is list[0]."""
    from inverse_flow_b200 import inv_conv_with_bp
    rng = np.random.default_rng(9)
    x = dev(rng.standard_normal((2, 4, 5, 5)))
    w = dev(make_weight(rng, 4, 4, 3, 3))
    out = x * 0.0
    z = inv_conv_with_bp.inverse(x, w, out)
    assert z[0].data_ptr() == out.data_ptr()
    back = inv_conv_with_bp.forward(z[0], w, torch.zeros_like(x))
    np.testing.assert_allclose(back[0].cpu().numpy(), x.cpu().numpy(), atol=1e-5)
    with pytest.raises(RuntimeError):
        inv_conv_with_bp.inverse(x.transpose(2, 3), w, out)        # CHECK_CONTIGUOUS
    # the reference's backward calls dw(input_x, kernel, output_grad, M, out) (inv_conv.py:79): that argument
    # order is not what the weight gradient depends on -- it must fail loudly, never compute something else
    g = dev(rng.standard_normal((2, 4, 5, 5)))
    with pytest.raises(TypeError):
        inv_conv_with_bp.dw(x, w, g, None, torch.zeros_like(w))
    dx = inv_conv_with_bp.dy(g, w, None, torch.zeros_like(g))[0]
    dw = inv_conv_with_bp.dw(saved_output=z[0], kernel=w, grad_input=dx, output=torch.zeros_like(w))[0]
    x64, w64, g64 = (t.cpu().numpy().astype(np.float64) for t in (x, w, g))
    dx_ref, dw_ref = oracle.backward(g64, oracle.inverse(x64, w64, 4), w64, 4)
    assert oracle.max_rel_err(dx.cpu().numpy(), dx_ref) < TOL and oracle.max_rel_err(dw.cpu().numpy(), dw_ref) < TOL


def test_caller_allocated_outputs_are_validated(IF):
    """a wrong-shaped, wrong-dtype or aliasing `out`, a channels_last activation, or prepared weights of another
    kernel must raise -- never become an out-of-bounds device write"""
    rng = np.random.default_rng(19)
    x = dev(rng.standard_normal((3, 4, 6, 6)))
    g = dev(rng.standard_normal((3, 4, 6, 6)))
    w = dev(make_weight(rng, 4, 4, 3, 3))
    y = IF.inverse(x, w, groups=1)
    for bad in (torch.empty(3, 4, 6, 5, device="cuda"), torch.empty(3, 4, 6, 6, device="cuda", dtype=torch.float64),
                torch.empty(3, 4, 6, 6)):
        with pytest.raises((ValueError, RuntimeError)):
            IF.bwd_input(g, w, groups=1, out=bad)
        with pytest.raises((ValueError, RuntimeError)):
            IF.inverse(x, w, groups=1, out=bad)
    with pytest.raises(ValueError):
        IF.bwd_input(g, w, groups=1, out=g)                        # aliasing
    dx = IF.bwd_input(g, w, groups=1)
    for bad in (torch.empty(4, 4, 3, 2, device="cuda"), torch.empty(4, 4, 3, 3, device="cuda", dtype=torch.float16)):
        with pytest.raises((ValueError, RuntimeError)):
            IF.bwd_weight(dx, y, w, groups=1, out=bad)
    with pytest.raises(RuntimeError, match="channels_last"):
        IF.inverse(x.contiguous(memory_format=torch.channels_last), w, groups=1)
    other = IF.Prepared(dev(make_weight(rng, 4, 4, 2, 2)), 1)
    with pytest.raises(ValueError):
        IF.inverse(x, w, groups=1, prepared=other)


def test_inverse_once_matches_prepare_then_inverse(IF):
    """ifk_inverse_once_f32: the reference's one-call inverse(input, kernel, output) (inv_conv_with_bp_general.cpp:19-28)"""
    import ctypes
    from inverse_flow_b200 import _native
    lib = _native.load()
    rng = np.random.default_rng(20)
    for (B, C, H, W, k, groups) in [(5, 12, 8, 8, 3, 1), (7, 4, 14, 14, 2, 4)]:
        x = dev(rng.standard_normal((B, C, H, W)))
        w = dev(make_weight(rng, C, C, k, k, 0.05))
        p = _native.problem(B, C, H, W, k, k, C, groups)
        scratch = torch.empty(lib.ifk_prepared_floats(ctypes.byref(p)), device="cuda")
        y = torch.empty_like(x)
        _native.check(lib.ifk_inverse_once_f32(ctypes.byref(p), x.data_ptr(), w.data_ptr(), scratch.data_ptr(), y.data_ptr(),
                                               _native.current_stream(x.device)))
        assert torch.equal(y, IF.inverse(x, w, groups=groups))


CHAIN_CASES = [
    # (B, C, H, W, k, groups), layer orientations
    ((100, 12, 16, 16, 3, 1), ("TL", "TR", "BL", "BR")),          # Inv_FlowUnit at the imagenet32 / cifar level-1 shape
    ((100, 24, 8, 8, 3, 1), ("TL", "TR", "BL", "BR")),
    ((100, 48, 4, 4, 3, 1), ("TL", "TR", "BL", "BR")),
    ((37, 48, 4, 4, 3, 4), ("BR", "TL", "TL", "BL", "TR")),
    ((300, 12, 6, 6, 3, 1), ("TL", "BR", "TR")),                   # more images than CTAs: a stripe per CTA
    ((5, 12, 5, 7, 3, 1), ("TL",) * 11),                           # more layers than one launch takes
    ((4, 8, 7, 7, 2, 1), ("TL", "TR", "BL", "BR")),               # no resident chain kernel: served layer by layer
]


@pytest.mark.parametrize("case", CHAIN_CASES, ids=lambda c: "x".join(map(str, c[0])) + "-" + "".join(c[1]))
def test_inverse_chain_is_bit_identical_to_single_launches(IF, case):
    """ifk_inverse_chain_f32: consecutive layers in one launch, the image never leaving shared memory -- every
    layer's output bit-identical to the per-layer launches (and hence inside the oracle's tolerance)"""
    (B, C, H, W, k, groups), orients = case
    rng = np.random.default_rng(41)
    x = dev(rng.standard_normal((B, C, H, W)))
    ws = [dev(make_weight(rng, C, C, k, k, 0.03)) for _ in orients]
    prepared = [IF.Prepared(w, groups) for w in ws]
    ys = IF.inverse_chain(x, prepared, orients)
    cur = x
    for w, pr, o, y in zip(ws, prepared, orients, ys):
        ref = IF.inverse(cur, w, prepared=pr, orient=o)
        assert torch.equal(y, ref)
        cur = ref
    x64 = x[:3].cpu().numpy().astype(np.float64)
    for w, o in zip(ws, orients):
        x64 = oracle.inverse(x64, w.cpu().numpy().astype(np.float64), groups, orient=o)
    assert oracle.max_rel_err(ys[-1][:3].cpu().numpy(), x64) < TOL


def test_inv_flow_unit_is_one_chained_launch_with_the_layers_gradients():
    """Inv_FlowUnit (reference inf/layers/inv_flow.py:28-53) = TL -> TR -> BL -> BR in one launch; outputs and all
    gradients equal those of the four layers applied one after the other"""
    from inverse_flow_b200.layers import Inv_FlowUnit
    torch.manual_seed(3)
    unit = Inv_FlowUnit(12, 12, (3, 3), groups=1).cuda()
    x = torch.randn(9, 12, 16, 16, device="cuda", requires_grad=True)
    g = torch.randn(9, 12, 16, 16, device="cuda")
    y, ldj = unit(x)
    y.backward(g)
    got = [x.grad.clone()] + [c.weight_fwd.grad.clone() for c in unit._convs()]
    x.grad = None
    for c in unit._convs():
        c.weight_fwd.grad = None
    cur = x
    for c in unit._convs():
        cur, _ = c(cur)
    cur.backward(g)
    want = [x.grad] + [c.weight_fwd.grad for c in unit._convs()]
    assert ldj == 0.0 and torch.equal(y, cur)
    for a, b in zip(got, want):
        assert torch.equal(a, b)
    # (round trip through four layers of the reference initialisation -- W[:, -1, -1, -1] = 1 makes each operator
    #  amplify by an order of magnitude, |y| ~ 5e2 here -- so the reconstruction carries that conditioning)
    assert oracle.max_rel_err(unit.reverse(y.detach()).cpu().numpy(), x.detach().cpu().numpy()) < 5e-3


def test_runs_on_the_callers_stream_and_in_a_graph(IF):
    """asynchronous on the current stream, capturable in a CUDA graph (no device syncs)."""
    rng = np.random.default_rng(10)
    x = dev(rng.standard_normal((8, 12, 16, 16)))
    w = dev(make_weight(rng, 12, 12, 3, 3, 0.02))
    ref = IF.inverse(x, w, groups=1)
    prepared = IF.Prepared(w, 1)
    out = torch.empty_like(x)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        IF.inverse(x, w, out=out, prepared=prepared)          # warm-up on the side stream
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    out.zero_()
    with torch.cuda.graph(graph):
        IF.inverse(x, w, out=out, prepared=prepared)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, ref)


def test_batched_entry_points_match_single_calls(IF):
    """ifk_prepare_many_f32 / ifk_bwd_weight_partial_f32 + ifk_bwd_weight_reduce_many_f32 give
    bit-identical results to the per-layer calls, and the stack runner matches the oracle."""
    import ctypes
    from inverse_flow_b200 import _native
    from inverse_flow_b200.stack import InvConvStack
    lib = _native.load()
    stack = InvConvStack([(8, 7, 7, 2, 3), (12, 6, 6, 3, 2)], batch=5, groups=1, seed=3)
    torch.manual_seed(0)
    for st in stack.stages:
        st.act[0].normal_()
        st.grad_in.normal_()
    stack.forward_backward()
    torch.cuda.synchronize()
    for st in stack.stages:
        x, g = st.act[0], st.grad_in
        for i in range(st.n):
            single = IF.Prepared(st.w[i], 1)
            assert torch.equal(single.buffer, st.prepared[i])
            y = IF.inverse(x, st.w[i], groups=1)
            assert torch.equal(y, st.act[i + 1])
            x = y
        for i in reversed(range(st.n)):
            dx, dw = IF.backward(g, st.act[i + 1], st.w[i], groups=1)
            assert torch.equal(dx, st.dxs[i])
            assert torch.equal(dw, st.dw[i])
            g = dx
        w64 = [w.cpu().numpy().astype(np.float64) for w in st.w]
        cur = st.act[0].cpu().numpy().astype(np.float64)
        for w in w64:
            cur = oracle.inverse(cur, w, 1)
        assert oracle.max_rel_err(st.act[st.n].cpu().numpy(), cur) < TOL
    # the captured graph (fork/join of the dW side stream) reproduces the eager run
    ref_bucket = stack.grad_bucket.clone()
    stack.grad_bucket.zero_()
    stack.capture()
    stack.step()
    torch.cuda.synchronize()
    assert torch.equal(stack.grad_bucket, ref_bucket)


def test_if_glow_model_trains_and_inverts():
    """the layers inside a Glow-style model: autograd end to end, a few Adam steps
    lower the loss, and reverse(forward(x)) reconstructs the input (reference plot_recon check,
    inf/train/experiment.py:440-473)."""
    from inverse_flow_b200 import glow
    torch.manual_seed(0)
    model = glow.IFGlow((1, 28, 28), num_blocks=2, block_size=3, kernel_size=2, coupling_width=32, groups=1).cuda()
    x = torch.rand(16, 1, 28, 28, device="cuda") - 0.5
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    losses = []
    for _ in range(8):
        opt.zero_grad()
        loss = model.loss(x)
        loss.backward()
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < losses[0]
    model.eval()
    latents, logp = model(x)
    assert torch.isfinite(logp).all()
    rec = model.reverse([z.detach() for z in latents])
    np.testing.assert_allclose(rec.cpu().numpy(), x.cpu().numpy(), atol=1e-3)


def test_fincflow_layers_reverse_through_the_inverse_kernels():
    """PaddedConv2d.reverse (level 1, full C) and Finc_FlowUnit.reverse_level2
    (one 4-group solve) invert the padded nn.Conv2d forward, for every corner order."""
    from inverse_flow_b200.layers import Finc_FlowUnit, PaddedConv2d
    torch.manual_seed(0)
    for order in ("TL", "TR", "BL", "BR"):
        m = PaddedConv2d(6, 6, (3, 3), order=order).cuda()
        x = torch.randn(3, 6, 9, 7, device="cuda")
        out, ld = m(x)
        assert ld == 0.0 and out.shape == x.shape
        rec, ld2 = m.reverse(out.detach())
        np.testing.assert_allclose(rec.cpu().numpy(), x.cpu().numpy(), atol=1e-4)
        # against the oracle in the layer's orientation
        dims = {"TL": None, "TR": [3], "BL": [2], "BR": [2, 3]}[order]
        flip = (lambda t: t) if dims is None else (lambda t: torch.flip(t, dims))
        y_ref = oracle.inverse(flip(out.detach()).cpu().numpy().astype(np.float64),
                               m.tl_weight().cpu().numpy().astype(np.float64), 1)
        assert oracle.max_rel_err(flip(rec).cpu().numpy(), y_ref) < TOL
    unit = Finc_FlowUnit(8, 8, (3, 3)).cuda()
    x = torch.randn(2, 8, 6, 6, device="cuda")
    out, _ = unit(x)
    r2 = unit.reverse(out.detach())
    r1 = unit.reverse_level1(out.detach())
    np.testing.assert_allclose(r2.cpu().numpy(), x.cpu().numpy(), atol=1e-4)
    np.testing.assert_allclose(r1.cpu().numpy(), r2.cpu().numpy(), atol=1e-5)


def test_host_to_host_step_matches_the_resident_step():
    """the end-to-end path (pinned host buffers, copies overlapped on their own stream, one graph)
    returns exactly what the device-resident step computes."""
    from inverse_flow_b200.stack import InvConvStack
    stack = InvConvStack([(4, 6, 6, 2, 3), (8, 3, 3, 2, 2)], batch=7, groups=1, seed=5)
    hb = stack.make_host_buffers()
    for st, x, g in zip(stack.stages, hb["x"], hb["g"]):
        st.act[0].copy_(x)
        st.grad_in.copy_(g)
    stack.forward_backward()
    torch.cuda.synchronize()
    ref_y = [st.act[st.n].clone() for st in stack.stages]
    ref_dx = [st.dx.clone() for st in stack.stages]
    ref_dw = stack.grad_bucket.clone()
    for st in stack.stages:
        st.act[st.n].zero_()
    stack.grad_bucket.zero_()
    for _ in range(2):
        h2d, d2h = stack.step_host(hb)
    assert h2d > 0 and d2h > h2d
    for k in range(2):
        assert torch.equal(hb["y"][k], ref_y[k].cpu()) and torch.equal(hb["dx"][k], ref_dx[k].cpu())
    assert torch.equal(hb["dw"], ref_dw.cpu())


# ---- fused neighbours: ActNorm affine + Squeeze re-indexing inside the solve --------------
FUSED_CASES = [
    # (B, C, H, W, k, groups, squeeze, orient)       C, H, W: the SQUEEZED geometry the layer works on
    (100, 12, 16, 16, 3, 1, True, "TL"), (100, 12, 16, 16, 3, 1, False, "TL"), (9, 24, 8, 8, 3, 1, True, "TL"),
    (7, 48, 4, 4, 3, 1, True, "TL"), (5, 12, 16, 16, 3, 1, True, "BR"), (5, 12, 6, 10, 3, 1, False, "TR"),
    (6, 48, 4, 4, 3, 4, True, "TL"), (300, 12, 6, 6, 3, 1, True, "BL"),
    (4, 4, 14, 14, 2, 1, True, "TL"),        # shuffle-kernel geometry: served by the unfused sequence
]


@pytest.mark.parametrize("case", FUSED_CASES, ids=lambda c: "x".join(map(str, c)))
def test_fused_actnorm_squeeze_matches_oracle_composition(IF, case):
    """y = L^-1(s * S(x) + b) and dz = S^T(s * L^-T g) in one launch each against the float64 oracle applied to the
    explicitly squeezed / scaled tensors (reference actnorm.py:36, squeeze.py:5-24), plus the raw dX the dW kernel reads"""
    B, C, H, W, k, groups, squeeze, orient = case
    rng = np.random.default_rng(41)
    xin = rng.standard_normal((B, C // 4, 2 * H, 2 * W) if squeeze else (B, C, H, W)).astype(np.float32)
    g = rng.standard_normal((B, C, H, W)).astype(np.float32)
    s = np.exp(-0.3 * rng.standard_normal(C)).astype(np.float32)
    b = (0.5 * rng.standard_normal(C)).astype(np.float32)
    w = make_weight(rng, C, C, k, k, 0.02)

    def sq(a):
        Bn, Cn, Hn, Wn = a.shape
        return a.reshape(Bn, Cn, Hn // 2, 2, Wn // 2, 2).transpose(0, 1, 3, 5, 2, 4).reshape(Bn, Cn * 4, Hn // 2, Wn // 2)

    def unsq(a):
        Bn, Cn, Hn, Wn = a.shape
        return a.reshape(Bn, Cn // 4, 2, 2, Hn, Wn).transpose(0, 1, 4, 2, 5, 3).reshape(Bn, Cn // 4, Hn * 2, Wn * 2)

    u = (sq(xin) if squeeze else xin).astype(np.float64) * s.astype(np.float64)[None, :, None, None] + \
        b.astype(np.float64)[None, :, None, None]
    y_ref = oracle.inverse(u, w.astype(np.float64), groups, threads=4, orient=orient)
    dx_ref, _ = oracle.backward(g.astype(np.float64), y_ref, w.astype(np.float64), groups, threads=4, orient=orient)
    t = dx_ref * s.astype(np.float64)[None, :, None, None]
    dz_ref = unsq(t) if squeeze else t
    wd = dev(w)
    y = IF.inverse_fused(dev(xin), wd, in_scale=dev(s), in_bias=dev(b), squeeze=squeeze, groups=groups, orient=orient)
    dx, dz = IF.bwd_input_fused(dev(g), wd, out_scale=dev(s), squeeze=squeeze, groups=groups, orient=orient)
    torch.cuda.synchronize()
    assert oracle.max_rel_err(y.cpu().numpy(), y_ref) < TOL
    assert oracle.max_rel_err(dx.cpu().numpy(), dx_ref) < TOL
    assert oracle.max_rel_err(dz.cpu().numpy(), dz_ref) < TOL
    assert tuple(dz.shape) == xin.shape


def test_fused_layer_matches_the_unfused_layers_through_autograd():
    """ActNormInvFlow (one fused launch per direction) against ActNorm ops + Squeeze + inv_flow_with_pad composed in
    PyTorch around the unfused kernels: outputs, log-det and the gradients of x, translation, log_scale and the weight"""
    from inverse_flow_b200.layers import ActNormInvFlow, inv_conv_4d
    from inverse_flow_b200 import functional as F2
    torch.manual_seed(3)
    for squeeze in (True, False):
        layer = ActNormInvFlow(12, (3, 3), order="TL", squeeze=squeeze, groups=1).cuda()
        x = torch.randn(20, 3, 32, 32, device="cuda") if squeeze else torch.randn(20, 12, 16, 16, device="cuda")
        x = (x * 1.7 + 0.3).requires_grad_(True)
        out, ldj = layer(x)
        gsum = torch.randn_like(out)
        (out * gsum).sum().backward()
        got = [out.detach(), x.grad.clone(), layer.translation.grad.clone(), layer.log_scale.grad.clone(),
               layer.conv.weight_fwd.grad.clone()]
        # unfused reference composition with the same (now initialised) parameters
        x2 = x.detach().clone().requires_grad_(True)
        t2 = layer.translation.detach().clone().requires_grad_(True)
        ls2 = layer.log_scale.detach().clone().requires_grad_(True)
        w2 = layer.conv.weight_fwd.detach().clone().requires_grad_(True)
        u = F2.space_to_depth(x2) if squeeze else x2
        u = (u - t2.view(1, -1, 1, 1)) * torch.exp(-ls2).view(1, -1, 1, 1)
        out2 = inv_conv_4d(u.contiguous(), w2, 1, "TL")
        (out2 * gsum).sum().backward()
        want = [out2.detach(), x2.grad, t2.grad, ls2.grad, w2.grad]
        for a, b_, name in zip(got, want, ("out", "dx", "dtranslation", "dlog_scale", "dW")):
            err = float((a - b_).abs().max() / b_.abs().max())
            assert err < 2e-5, (squeeze, name, err)
        assert torch.allclose(ldj, -layer.log_scale.sum().expand(20) * 16 * 16)
        # reverse undoes forward
        back = layer.reverse(out.detach())
        assert float((back - x.detach()).abs().max() / x.detach().abs().max()) < 1e-4
