"""Property tests of the CPU oracle over random geometries (hypothesis): the identities that the GPU
parity tests rely on at sizes the golden vectors do not cover."""
import numpy as np
from hypothesis import given, settings, strategies as st

from conftest import make_weight
from oracle import oracle


@st.composite
def problems(draw):
    groups = draw(st.sampled_from([1, 2, 4]))
    cg = draw(st.integers(1, 4))
    C = groups * cg
    H, W = draw(st.integers(1, 7)), draw(st.integers(1, 7))
    KH, KW = draw(st.integers(1, 4)), draw(st.integers(1, 4))
    B = draw(st.integers(1, 3))
    orient = draw(st.sampled_from(["TL", "TR", "BL", "BR"]))
    seed = draw(st.integers(0, 2 ** 16))
    return B, C, H, W, KH, KW, groups, orient, seed


@settings(max_examples=40, deadline=None)
@given(problems())
def test_round_trip_linearity_and_adjoint(p):
    B, C, H, W, KH, KW, groups, orient, seed = p
    rng = np.random.default_rng(seed)
    w = make_weight(rng, C, C, KH, KW, 0.2).astype(np.float64)
    x = rng.standard_normal((B, C, H, W))
    x2 = rng.standard_normal((B, C, H, W))
    g = rng.standard_normal((B, C, H, W))
    y = oracle.inverse(x, w, groups, orient=orient)
    # conv(inverse(x)) == x, and inverse(conv(x)) == x
    assert oracle.max_rel_err(oracle.conv(y, w, groups, orient=orient), x) < 1e-9
    assert oracle.max_rel_err(oracle.inverse(oracle.conv(x, w, groups, orient=orient), w, groups, orient=orient), x) < 1e-9
    # linear in x
    lin = oracle.inverse(2.0 * x - 3.0 * x2, w, groups, orient=orient)
    assert oracle.max_rel_err(lin, 2.0 * y - 3.0 * oracle.inverse(x2, w, groups, orient=orient)) < 1e-9
    # <g, L^-1 x> == <L^-T g, x>
    dx = oracle.bwd_input(g, w, groups, orient=orient)
    lhs, rhs = float(np.sum(g * y)), float(np.sum(dx * x))
    assert abs(lhs - rhs) <= 1e-9 * max(1.0, abs(lhs), abs(rhs))
    # the raster and the wavefront order of the solve agree bit for bit; threads do not change results
    assert np.array_equal(y, oracle.inverse(x, w, groups, wavefront=True, orient=orient))
    assert np.array_equal(y, oracle.inverse(x, w, groups, threads=3, orient=orient))
    # masked weight entries (centre tap, diagonal and above) get exactly zero gradient
    dw = oracle.bwd_weight(dx, y, w.shape, groups, orient=orient)
    cgp = C // groups
    for c in range(C):
        assert np.all(dw[c, (c % cgp):, KH - 1, KW - 1] == 0.0)
    assert np.all(dw[:, cgp:] == 0.0)
