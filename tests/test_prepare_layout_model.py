"""CPU models of the index algebra of the weight-preparation kernels (inverse_flow_b200/csrc/ifk_prepare.cu,
wave_pack_kernel in ifk_solve_wave.cu).  They do NOT run the CUDA code -- the GPU suite does that
(tests/test_parity_gpu.py::test_prepared_rows_match_the_float64_fold_for_every_slab_size and every solve test) -- they
pin the layout contracts those kernels were written against, in plain numpy, so that an edit of the staging strides,
the tile shape or the thread decomposition can be checked without a GPU:

* the substitution with the column of T in registers (four partial sums, entries above the diagonal forced to zero)
  is (I + A0)^-1 in float32;
* the staged tap of the tiled product -- row stride RS, transposed for the adjoint, a thread forming four adjacent
  columns -- yields -T W_q and -T^T W_q^T;
* the pack kernel's decomposition (one thread per 16-byte word, tap table) writes every word of the packed copy exactly
  once and is the same gather as one thread per pair with a tap search (its first version).
"""
import numpy as np
import pytest


def tap_stride(cg):
    """prepare_tap_stride(): rows 16-byte aligned, an odd multiple of 4 floats"""
    rs = (cg + 3) // 4 * 4
    return rs + 4 if rs % 8 == 0 else rs


@pytest.mark.parametrize("cg", [4, 6, 8, 12, 24, 48])
def test_register_substitution_is_the_inverse_of_the_unit_lower_matrix(cg):
    rng = np.random.default_rng(cg)
    A = np.tril(rng.standard_normal((cg, cg)) * 0.3, -1).astype(np.float32)
    T = np.zeros((cg, cg), np.float32)
    for j in range(cg):                        # t_column_in_registers<CG>(A, T, TS, j)
        t = np.zeros(cg, np.float32)
        for i in range(cg):
            s = np.zeros(4, np.float32)
            for k in range(i):
                s[k & 3] = np.float32(A[i, k] * t[k] + s[k & 3])
            v = -((s[0] + s[1]) + (s[2] + s[3]))
            t[i] = 0.0 if i < j else (1.0 if i == j else v)
        T[:, j] = t
    ref = np.linalg.inv(np.eye(cg) + A.astype(np.float64))
    assert np.abs(T - ref).max() / np.abs(ref).max() < 1e-6
    assert np.all(np.triu(T, 1) == 0.0) and np.all(np.diag(T) == 1.0)


def tiled_tap_product(cg, cw, kh, kw, direction, t, w, L):
    """prepare_taps_kernel for one tap: L = tap 0 of this direction as written by prepare_t_kernel"""
    K, ts, rs = kh * kw, cg + 1, tap_stride(cg)
    qh, qw = divmod(t, kw)
    a = (kh - 1 - qh) * kw + (kw - 1 - qw)
    Wq = np.full(cg * rs + 8, np.nan)
    for e in range(cg * (rs - cg)):            # pad columns
        k = e // (rs - cg)
        Wq[k * rs + cg + (e - k * (rs - cg))] = 0.0
    flat, row_stride = w.reshape(-1), cw * K
    for e in range(cg * cg):
        r, c = divmod(e, cg)
        Wq[(r * rs + c) if direction == 0 else (c * rs + r)] = flat[r * row_stride + c * K + a]
    Ls = np.zeros(cg * ts)
    for co in range(cg):
        Ls[co * ts:co * ts + cg] = L[co]
    out = np.full((cg, cg), np.nan)
    cg4 = (cg + 3) >> 2
    for e in range(cg * cg4):
        co, j = divmod(e, cg4)
        j4 = 4 * j
        k0, k1 = (0, co + 1) if direction == 0 else (co, cg)
        acc = np.zeros(4)
        for k in range(k0, k1):
            r4 = Wq[j4 + k * rs:j4 + k * rs + 4]
            assert not np.isnan(r4).any()      # only staged or padded cells are read
            acc += Ls[co * ts + k] * r4
        for q in range(4):
            if j4 + q < cg:
                out[co, j4 + q] = -acc[q]
    return out


@pytest.mark.parametrize("cg,cw", [(1, 1), (3, 3), (4, 6), (6, 6), (12, 12), (13, 16), (48, 48)])
def test_tiled_tap_product_layout(cg, cw):
    rng = np.random.default_rng(7)
    kh, kw = 3, 2
    w = rng.standard_normal((cg, cw, kh, kw))
    T = np.linalg.inv(np.eye(cg) + np.tril(w[:, :cg, kh - 1, kw - 1], -1))
    for direction, L in ((0, T), (1, T.T)):
        for t in range(1, kh * kw):
            qh, qw = divmod(t, kw)
            Wt = w[:, :cg, kh - 1 - qh, kw - 1 - qw]
            ref = -(T @ Wt) if direction == 0 else -(T.T @ Wt.T)
            np.testing.assert_allclose(tiled_tap_product(cg, cw, kh, kw, direction, t, w, L), ref, atol=1e-10)


def wave_code(qh, qw, chan, is_x):
    return qh | (qw << 8) | (chan << 16) | ((1 << 30) if is_x else 0)


def pack_dims(cg, kh, kw, cc, ns, vec):
    K, cgv, nct = kh * kw, cg // vec, cg // cc
    lpp = nct * ns
    nft = (kw > 1) + (kh > 1)
    not_ = K - 1 - nft
    NF, NO = nft * cgv, (not_ + 1) * cgv
    nvf, nvo = (NF + ns - 1) // ns, (NO + ns - 1) // ns
    pf, po = nvf * vec // 2, nvo * vec // 2
    npairs = cc * (pf + po)
    return K, cgv, lpp, not_, NF, NO, nvf, nvo, pf, po, npairs, (npairs + 1) // 2


def pack_entry(prep, dims, cg, kw, cc, ns, vec, KDP, C, pstride, layer, direction, G, l, f, tap_of):
    """the two floats and the code of flat pair f of lane l (shared by both decompositions)"""
    K, cgv, lpp, not_, NF, NO, nvf, nvo, pf, po, npairs, nw4 = dims
    ks, ct = l % ns, l // ns
    i, c_ = divmod(f, cc)
    fresh = i < pf
    ip = i if fresh else i - pf
    hv = vec >> 1
    j, e2 = (ip, 0) if hv == 1 else (ip >> 1, ip & 1)
    ent = j * ns + ks
    slot = (0 if fresh else nvf) + j
    if ent >= (NF if fresh else NO):
        return 0.0, 0.0, 0, slot, c_, e2
    ti, qv = divmod(ent, cgv)
    t, is_x = 0, False
    if fresh:
        t = 1 if (kw > 1 and ti == 0) else kw
    elif ti == not_:
        is_x = True
    else:
        t = tap_of(ti)
    src = layer * pstride + ((direction * C + G * cg + ct * cc + c_) * KDP + t * cg + qv * vec + 2 * e2)
    return prep[src], prep[src + 1], wave_code(t // kw, t % kw, qv * vec, is_x), slot, c_, e2


@pytest.mark.parametrize("cg,cc,ns,vec,groups", [(12, 6, 4, 2, 1), (24, 6, 8, 2, 1), (48, 6, 16, 2, 1), (6, 6, 4, 2, 4),
                                                 (12, 3, 2, 2, 2), (12, 6, 8, 2, 1)])
def test_pack_kernel_decomposition_covers_the_packed_copy_once(cg, cc, ns, vec, groups):
    kh = kw = 3
    C, count = cg * groups, 2
    KDP = (kh * kw * cg + 3) // 4 * 4
    pstride = 2 * C * KDP + 7
    rng = np.random.default_rng(0)
    prep = rng.standard_normal(count * pstride)
    dims = pack_dims(cg, kh, kw, cc, ns, vec)
    K, cgv, lpp, not_, NF, NO, nvf, nvo, pf, po, npairs, nw4 = dims
    pack_floats = 2 * groups * nw4 * lpp * 4

    def searched_tap(ti):                       # first version: a search per pair
        n = 0
        for tt in range(1, K):
            if tt // kw + tt % kw < 2:
                continue
            if n == ti:
                return tt
            n += 1
        return 0

    table = [tt for tt in range(1, K) if tt // kw + tt % kw >= 2]        # the host-filled tap table

    # one thread per pair
    a_pack = np.full((count, pack_floats), np.nan)
    a_codes = np.full((count, (nvf + nvo) * lpp), -1, dtype=np.int64)
    for e in range(count * 2 * groups * lpp * 2 * nw4):
        r, f = divmod(e, 2 * nw4)
        r, l = divmod(r, lpp)
        r, G = divmod(r, groups)
        layer, direction = divmod(r, 2)
        w0 = w1 = 0.0
        if f < npairs:
            w0, w1, code, slot, c_, e2 = pack_entry(prep, dims, cg, kw, cc, ns, vec, KDP, C, pstride, layer, direction, G, l, f,
                                                    searched_tap)
            if c_ == 0 and e2 == 0 and direction == 0 and G == 0:
                a_codes[layer, slot * lpp + l] = code
        d = ((direction * groups + G) * nw4 + f // 2) * lpp * 4 + l * 4 + (f & 1) * 2
        a_pack[layer, d], a_pack[layer, d + 1] = w0, w1

    # one thread per 16-byte word: grid.y = (layer, direction, group), grid.x over nw4 * lpp
    b_pack = np.full((count, pack_floats), np.nan)
    b_codes = np.full((count, (nvf + nvo) * lpp), -1, dtype=np.int64)
    writes = np.zeros((count, pack_floats), dtype=np.int64)
    for by in range(count * 2 * groups):
        G, r = by % groups, by // groups
        direction, layer = r & 1, r >> 1
        for e in range(nw4 * lpp):
            f2, l = divmod(e, lpp)
            w = [0.0] * 4
            for h in range(2):
                f = 2 * f2 + h
                if f >= npairs:
                    continue
                w[2 * h], w[2 * h + 1], code, slot, c_, e2 = pack_entry(prep, dims, cg, kw, cc, ns, vec, KDP, C, pstride, layer,
                                                                        direction, G, l, f, lambda ti: table[ti])
                if c_ == 0 and e2 == 0 and direction == 0 and G == 0:
                    b_codes[layer, slot * lpp + l] = code
            d = ((direction * groups + G) * nw4 + f2) * lpp * 4 + l * 4
            b_pack[layer, d:d + 4] = w
            writes[layer, d:d + 4] += 1
    assert np.all(writes == 1)
    assert np.array_equal(a_pack, b_pack) and not np.isnan(b_pack).any()
    assert np.array_equal(a_codes, b_codes) and np.all(b_codes >= 0)
