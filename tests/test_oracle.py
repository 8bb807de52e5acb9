"""CPU tests: the oracle (oracle/ifk_oracle.c) against the reference's own outputs.

Golden vectors come from the reference itself (tests/golden/make_golden.py): solve_mc.py
`solve` / `solve_parallel`, the compiled Cython solver, the derivative of that solver and
torch's F.conv2d.  float64 inverse must match BIT FOR BIT (same operation order, no FMA
contraction); gradients to 1e-7 relative (the golden dW is a finite difference).
"""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, make_weight
from oracle import oracle


def test_inverse_bit_exact_vs_reference_solvers(golden):
    y = oracle.inverse(golden["x"], golden["w"], golden["groups"])
    assert np.array_equal(y, golden["y_solve"]), "raster solve_mc.solve"
    yw = oracle.inverse(golden["x"], golden["w"], golden["groups"], wavefront=True)
    assert np.array_equal(yw, golden["y_solve"]), "wavefront order gives the same bits"
    if "y_solve_parallel" in golden:
        assert np.array_equal(yw, golden["y_solve_parallel"]), "solve_mc.solve_parallel"
        assert np.array_equal(yw, golden["y_cython"]), "solve_parallel_mc.pyx"


def test_conv_matches_torch_conv2d(golden):
    x = oracle.conv(golden["y_solve"], golden["w"], golden["groups"])
    assert oracle.max_rel_err(x, golden["conv_of_y"]) < 1e-14
    assert oracle.max_rel_err(x, golden["x"]) < 1e-13           # conv(inverse(x)) == x


def test_backward_matches_reference_derivative(golden):
    dx, dw = oracle.backward(golden["g"], golden["y_solve"], golden["w"], golden["groups"])
    assert oracle.max_rel_err(dx, golden["dx_ref"]) < 1e-12
    assert oracle.max_rel_err(dw, golden["dw_ref_fd"]) < 1e-7
    # entries the solver never reads get exactly zero gradient
    assert np.all(dw[golden["dw_ref_fd"] == 0.0] == 0.0)


def test_dense_toeplitz_cross_check(golden):
    """Independent of the loops above: dense L, numpy.linalg.solve."""
    if golden["x"][0].size > 200:
        pytest.skip("dense check only for tiny images")
    B, C, H, W = golden["x"].shape
    L = oracle.dense_L(golden["w"], H, W, golden["groups"])
    xf = golden["x"].reshape(B, -1).T
    y = np.linalg.solve(L, xf).T.reshape(golden["x"].shape)
    assert oracle.max_rel_err(oracle.inverse(golden["x"], golden["w"], golden["groups"]), y) < 1e-12
    dx = np.linalg.solve(L.T, golden["g"].reshape(B, -1).T).T.reshape(golden["x"].shape)
    assert oracle.max_rel_err(oracle.bwd_input(golden["g"], golden["w"], golden["groups"]), dx) < 1e-12
    assert abs(np.linalg.slogdet(L)[1]) < 1e-12               # unit diagonal: log|det| = 0


def test_f32_oracle_close_to_f64(golden):
    x32, w32 = golden["x"].astype(np.float32), golden["w"].astype(np.float32)
    y32 = oracle.inverse(x32, w32, golden["groups"])
    assert y32.dtype == np.float32
    assert oracle.max_rel_err(y32, golden["y_solve"]) < 5e-6


def test_threads_do_not_change_results():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((6, 4, 9, 7))
    g = rng.standard_normal(x.shape)
    w = make_weight(rng, 4, 4, 3, 3, 0.1, np.float64)
    for groups in (1, 2, 4):
        y1 = oracle.inverse(x, w, groups, threads=1)
        y4 = oracle.inverse(x, w, groups, threads=4)
        assert np.array_equal(y1, y4)
        d1 = oracle.backward(g, y1, w, groups, threads=1)
        d4 = oracle.backward(g, y1, w, groups, threads=4)
        assert np.array_equal(d1[0], d4[0]) and np.array_equal(d1[1], d4[1])


def test_round_trip_at_reference_test_shape():
    """tests/inf/test_layers.py:182-190: (1,4,5,5), k=3, atol 1e-3 round trip."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal((1, 4, 5, 5)).astype(np.float32)
    w = make_weight(rng, 4, 4, 3, 3, 0.01)
    for groups in (1, 4):
        y = oracle.inverse(x, w, groups)
        np.testing.assert_allclose(oracle.conv(y, w, groups), x, atol=1e-3)


def test_literal_kernels_coincide_with_contract_at_C4():
    """for C == 4 the shipped kernels are 4 independent single-channel
    solves == groups=4 of the math contract; for C < 4 they write nothing (zeros)."""
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 4, 6, 5))
    w = make_weight(rng, 4, 4, 3, 3, 0.1, np.float64)
    assert np.array_equal(oracle.literal_inverse(x, w), oracle.inverse(x, w, groups=4))
    # literal dy is L^-1 g (causal), i.e. inverse(g) -- not the true gradient L^-T g
    g = rng.standard_normal(x.shape)
    assert oracle.max_rel_err(oracle.literal_dy(g, w), oracle.inverse(g, w, groups=4)) < 1e-13
    assert oracle.max_rel_err(oracle.literal_dy(g, w), oracle.bwd_input(g, w, groups=4)) > 1e-3
    x3 = rng.standard_normal((2, 3, 4, 4))
    w3 = make_weight(rng, 3, 3, 3, 3, 0.1, np.float64)
    assert np.all(oracle.literal_inverse(x3, w3) == 0.0)


@pytest.mark.skipif(not os.path.isdir(os.path.join(ROOT, "oracle", "_ref")),
                    reason="oracle/_ref not built (needs the reference checkout)")
def test_against_compiled_reference_solver_random_shapes():
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    import solve_parallel_mc as cy
    rng = np.random.default_rng(7)
    for (B, C, H, W, k) in [(3, 1, 28, 28, 3), (2, 4, 14, 14, 2), (2, 12, 16, 16, 3),
                            (1, 3, 8, 12, 5), (2, 24, 8, 8, 3)]:
        x = rng.standard_normal((B, C, H, W))
        w = make_weight(rng, C, C, k, k, 0.05, np.float64)
        ref = np.asarray(cy.solve_parallel(x.copy(), w, (k, k)))
        assert np.array_equal(oracle.inverse(x, w, 1, threads=2), ref)


def test_empty_batch():
    w = make_weight(np.random.default_rng(0), 4, 4, 3, 3)
    x = np.zeros((0, 4, 5, 5), np.float32)
    assert oracle.inverse(x, w).shape == (0, 4, 5, 5)
    dw = oracle.bwd_weight(x, x, w.shape)
    assert dw.shape == w.shape and np.all(dw == 0)


@pytest.mark.parametrize("orient,pad,flip_axes", [
    ("TL", (2, 0, 2, 0), ()),          # F.pad order: (left, right, top, bottom)
    ("TR", (0, 2, 2, 0), (3,)),
    ("BL", (2, 0, 0, 2), (2,)),
    ("BR", (0, 2, 0, 2), (2, 3)),
])
def test_orientations_match_torch_conv2d_padded_towards_the_other_corner(orient, pad, flip_axes):
    """orientation semantics, checked without any flip of the data: F L F is the masked convolution with
    the zero padding on the opposite side(s) and the kernel reflected along the same axes (torch's
    F.conv2d on CPU, float64); inverse and the gradients are then pinned to it through
    conv(inverse(x)) == x and the adjoint / derivative identities."""
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(3)
    for (B, C, H, W, groups) in [(2, 4, 6, 5, 1), (2, 8, 5, 7, 4)]:
        y = rng.standard_normal((B, C, H, W))
        w = make_weight(rng, C, C, 3, 3, 0.2).astype(np.float64)
        wt = torch.from_numpy(oracle.masked_weight(w, groups))
        if flip_axes:
            wt = torch.flip(wt, flip_axes)
        ref = F.conv2d(F.pad(torch.from_numpy(y), pad), wt, groups=groups).numpy()
        got = oracle.conv(y, w, groups, orient=orient)
        assert oracle.max_rel_err(got, ref) < 1e-13, orient
        # the inverse undoes it, and the backward is its derivative
        x = got
        yy = oracle.inverse(x, w, groups, orient=orient)
        assert oracle.max_rel_err(yy, y) < 1e-9
        g = rng.standard_normal((B, C, H, W))
        dx = oracle.bwd_input(g, w, groups, orient=orient)
        xp = rng.standard_normal((B, C, H, W))
        # adjoint identity <g, L^-1 x'> == <L^-T g, x'>
        lhs = float(np.sum(g * oracle.inverse(xp, w, groups, orient=orient)))
        rhs = float(np.sum(dx * xp))
        assert abs(lhs - rhs) <= 1e-9 * max(1.0, abs(lhs))
        # dW by central differences of sum(g * inverse(x, W)) on a few live entries
        dw = oracle.bwd_weight(dx, yy, w.shape, groups, orient=orient)
        eps = 1e-6
        Cg = C // groups
        for (c, kc, a, b_) in [(1, 0, 0, 0), (C - 1, Cg - 1, 1, 2), (2, 1, 2, 1), (C - 1, 0, 2, 2)]:
            wp, wm = w.copy(), w.copy()
            wp[c, kc, a, b_] += eps
            wm[c, kc, a, b_] -= eps
            fd = (np.sum(g * oracle.inverse(x, wp, groups, orient=orient)) -
                  np.sum(g * oracle.inverse(x, wm, groups, orient=orient))) / (2 * eps)
            assert abs(fd - dw[c, kc, a, b_]) <= 1e-6 * max(1.0, abs(fd)), (orient, c, kc, a, b_)
