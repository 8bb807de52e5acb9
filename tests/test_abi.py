"""CPU tests of the C-ABI boundary: the library loads, exports every symbol include/ifk.h
declares, and validates arguments before touching the device (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from inverse_flow_b200 import _native


@pytest.fixture(scope="module")
def lib():
    from inverse_flow_b200.build import build_native
    build_native()
    return _native.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ifk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ifk_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    names = declared_symbols()
    assert set(names) == set(_native.EXPORTS), "keep _native.EXPORTS in sync with include/ifk.h"
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n


def test_version_and_status_strings(lib):
    assert lib.ifk_version() == 400
    assert lib.ifk_status_string(0) == b"ok"
    for code in (-1, -2, -3, -4, -5, -6):
        assert lib.ifk_status_string(code) not in (b"ok", b"unknown ifk status")


def test_sizes(lib):
    p = _native.problem(100, 12, 16, 16, 3, 3, 12, 1)
    # canonical rows: 2 directions x C rows x round_up(K*Cg, 4); geometries the pipelined wavefront kernel serves
    # carry its lane-major packed copy and entry codes behind them (a multiple of 4 floats: 16-byte strides)
    n = lib.ifk_prepared_floats(ctypes.byref(p))
    assert n > 2 * 12 * 108 and n % 4 == 0
    p4 = _native.problem(100, 12, 16, 16, 3, 3, 12, 4)
    assert lib.ifk_prepared_floats(ctypes.byref(p4)) == 2 * 12 * 28          # Cg = 3: shuffle kernel, no packed copy
    p2 = _native.problem(100, 4, 14, 14, 2, 2, 4, 1)
    assert lib.ifk_prepared_floats(ctypes.byref(p2)) == 2 * 4 * 16
    assert lib.ifk_prepared_floats(ctypes.byref(_native.problem(7, 12, 5, 9, 3, 3, 12, 1))) == n   # size-independent
    ws = lib.ifk_bwd_weight_workspace_bytes(ctypes.byref(p))
    assert ws > 0 and ws % (12 * 12 * 9 * 4) == 0


@pytest.mark.parametrize("kwargs,code", [
    (dict(C=0), -2), (dict(H=0), -2), (dict(KH=0), -2), (dict(B=-1), -2),
    (dict(groups=0), -3), (dict(groups=5), -3), (dict(groups=4, Cw=2), -3),
    (dict(orient=4), -6), (dict(orient=-1), -6),
])
def test_bad_geometry_is_rejected_before_any_launch(lib, kwargs, code):
    base = dict(B=2, C=12, H=8, W=8, KH=3, KW=3, Cw=12, groups=1)
    base.update(kwargs)
    p = _native.problem(**base)
    dummy = ctypes.c_void_p(16)
    assert lib.ifk_inverse_f32(ctypes.byref(p), dummy, dummy, dummy, None) == code
    assert lib.ifk_conv_f32(ctypes.byref(p), dummy, dummy, dummy, None) == code
    assert lib.ifk_prepared_floats(ctypes.byref(p)) == 0
    with pytest.raises(ValueError):
        _native.check(code)


def test_problem_struct_matches_the_header():
    """ifk_problem is ten ints, `orient` then `flags` last (include/ifk.h); names map to IFK_ORIENT_*"""
    assert ctypes.sizeof(_native.Problem) == 10 * ctypes.sizeof(ctypes.c_int)
    assert [f[0] for f in _native.Problem._fields_][-2:] == ["orient", "flags"]
    header = open(os.path.join(ROOT, "include", "ifk.h")).read()
    struct = header[header.index("typedef struct ifk_problem"):header.index("} ifk_problem;")]
    assert re.findall(r"int ([A-Za-z, ]+);", struct) == ["B, C, H, W", "KH, KW", "Cw", "groups", "orient", "flags"]
    assert [_native.problem(1, 4, 5, 5, 3, 3, 4, 1, o).orient for o in ("TL", "TR", "BL", "BR")] == [0, 1, 2, 3]
    with pytest.raises(ValueError):
        _native.problem(1, 4, 5, 5, 3, 3, 4, 1, "XX")


def test_null_pointers_are_rejected(lib):
    p = _native.problem(2, 4, 5, 5, 3, 3, 4, 1)
    dummy = ctypes.c_void_p(16)
    assert lib.ifk_inverse_f32(ctypes.byref(p), None, dummy, dummy, None) == -1
    assert lib.ifk_inverse_f32(None, dummy, dummy, dummy, None) == -1
    assert lib.ifk_bwd_weight_f32(ctypes.byref(p), dummy, dummy, dummy, None, None) == -1
    assert lib.ifk_prepare_f32(ctypes.byref(p), None, dummy, None) == -1


def test_empty_batch_is_a_noop_on_the_solve(lib):
    p = _native.problem(0, 4, 5, 5, 3, 3, 4, 1)
    dummy = ctypes.c_void_p(16)
    assert lib.ifk_inverse_f32(ctypes.byref(p), None, dummy, None, None) == 0
    assert lib.ifk_bwd_input_f32(ctypes.byref(p), None, dummy, None, None) == 0
    assert lib.ifk_conv_f32(ctypes.byref(p), None, dummy, None, None) == 0


def test_describe_solve_picks_the_on_chip_kernels_for_model_shapes(lib, monkeypatch):
    """model shapes stay on chip: the register/shuffle kernel where one warp's lanes cover the rows
    and a lane can hold its weights (MNIST-sized layers), the shared-memory resident kernel otherwise"""
    for shape, kind in [((100, 4, 14, 14, 2), "shfl<"), ((100, 8, 7, 7, 2), "shfl<"), ((64, 1, 28, 28, 3), "shfl<"),
                        ((100, 12, 16, 16, 3), "wave<"), ((100, 24, 8, 8, 3), "wave<"), ((100, 48, 4, 4, 3), "wave<"),
                        ((100, 12, 32, 32, 3), "wave<")]:
        B, C, H, W, k = shape
        d = _native.describe_solve(_native.problem(B, C, H, W, k, k, C, 1))
        assert d.startswith(kind), (shape, d)
    monkeypatch.setenv("IFK_SOLVE_SHFL", "0")
    assert _native.describe_solve(_native.problem(100, 4, 14, 14, 2, 2, 4, 1)).startswith("smem<")
    monkeypatch.delenv("IFK_SOLVE_SHFL")
    big = _native.describe_solve(_native.problem(8, 96, 64, 64, 7, 7, 96, 1))
    assert big.startswith("global")


def test_batched_entry_points_validate_arguments(lib):
    p = _native.problem(2, 4, 5, 5, 3, 3, 4, 1)
    dummy = ctypes.c_void_p(16)
    # count == 0 is a no-op; negative counts and NULL buffers are refused before any launch
    assert lib.ifk_prepare_many_f32(ctypes.byref(p), 0, None, 0, None, 0, None) == 0
    assert lib.ifk_prepare_many_f32(ctypes.byref(p), -1, dummy, 0, dummy, 0, None) == -2
    assert lib.ifk_prepare_many_f32(ctypes.byref(p), 2, None, 0, dummy, 0, None) == -1
    assert lib.ifk_bwd_weight_reduce_many_f32(ctypes.byref(p), 0, None, 0, None, 0, None) == 0
    assert lib.ifk_bwd_weight_reduce_many_f32(ctypes.byref(p), 3, dummy, 6, dummy, 0, None) == -2   # stride % 4
    assert lib.ifk_bwd_weight_reduce_many_f32(ctypes.byref(p), 3, None, 8, dummy, 0, None) == -1
    assert lib.ifk_bwd_weight_partial_f32(ctypes.byref(p), dummy, dummy, None, None) == -1
    bad = _native.problem(2, 4, 5, 5, 3, 3, 4, 3)
    assert lib.ifk_prepare_many_f32(ctypes.byref(bad), 1, dummy, 0, dummy, 0, None) == -3


def test_describe_solve_variants(lib, monkeypatch):
    """resident kernel for model shapes; the window kernel (ring of diagonals in shared memory, with
    and without a cluster) for images beyond shared memory; the older stream kernel where even the
    ring does not fit; the plain fallback only where neither registers nor clusters hold the weights"""
    d = lambda *a: _native.describe_solve(_native.problem(*a))
    assert d(64, 12, 64, 64, 3, 3, 12, 1).startswith("window<") and "cluster=1" in d(64, 12, 64, 64, 3, 3, 12, 1)
    csize = lambda text: int(re.search(r"cluster=(\d+)", text).group(1))
    # weights beyond one SM's register file: a cluster; wider when the batch leaves SMs idle
    assert d(8, 96, 32, 32, 3, 3, 96, 1).startswith("window<") and csize(d(8, 96, 32, 32, 3, 3, 96, 1)) >= 4
    assert d(512, 96, 32, 32, 3, 3, 96, 1).startswith("window<") and csize(d(512, 96, 32, 32, 3, 3, 96, 1)) == 4
    assert d(512, 48, 16, 16, 5, 5, 48, 1).startswith("window<") and csize(d(512, 48, 16, 16, 5, 5, 48, 1)) == 2
    assert d(8, 96, 128, 128, 3, 3, 96, 1).startswith("stream<")
    assert d(8, 96, 64, 64, 7, 7, 96, 1).startswith("global")
    monkeypatch.setenv("IFK_SOLVE_STREAM", "1")
    assert d(64, 12, 64, 64, 3, 3, 12, 1).startswith("stream<")
    monkeypatch.delenv("IFK_SOLVE_STREAM")
    monkeypatch.setenv("IFK_SOLVE_WINDOW", "1")
    assert d(3, 12, 16, 16, 3, 3, 12, 1).startswith("window<")


def test_kernel_selection_is_well_formed_over_the_sweep_grid(lib):
    """every cell of the microbenchmark grid (and around it) dispatches to a kernel whose launch
    configuration is legal: threads <= 1024, shared memory <= 227 KB, cluster size <= 16 and dividing the
    grid; the dW workspace and the prepared buffer have sizes; nothing needs a GPU"""
    kinds = {}
    for C in (1, 3, 4, 8, 12, 24, 48, 96):
        for H in (1, 4, 7, 14, 16, 28, 32, 64, 128):
            for k in (1, 2, 3, 5, 7):
                for B in (1, 8, 100, 512):
                    for groups in (1, 4):
                        if C % groups:
                            continue
                        p = _native.problem(B, C, H, H, k, k, C, groups)
                        d = _native.describe_solve(p)
                        kind = d.split("<")[0].split(" ")[0]
                        kinds[kind] = kinds.get(kind, 0) + 1
                        assert kind in ("shfl", "split", "wave", "smem", "window", "stream", "global"), d
                        m = re.search(r"smem=(\d+)B", d)
                        if m:
                            assert int(m.group(1)) <= 227 * 1024, d
                        m = re.search(r"threads=(\d+)(?:\+(\d+))?", d)
                        assert m and int(m.group(1)) + int(m.group(2) or 0) <= 1024, d
                        m = re.search(r"cluster=(\d+)", d)
                        if m:
                            cs = int(m.group(1))
                            gx = int(re.search(r"grid=(\d+)x", d).group(1))
                            assert cs in (1, 2, 4, 8, 16) and gx % cs == 0, d
                        assert re.search(r"grid=(\d+)x(\d+)", d), d
                        assert lib.ifk_prepared_floats(ctypes.byref(p)) > 0
                        assert lib.ifk_bwd_weight_workspace_bytes(ctypes.byref(p)) > 0
    # the grid exercises every solve kernel
    assert set(kinds) == {"shfl", "wave", "smem", "window", "stream", "global"}, kinds


def test_fused_entry_points_validate_before_any_launch(lib):
    """ifk_inverse_fused_f32 / ifk_bwd_input_fused_f32 (include/ifk.h, struct ifk_fused): argument errors and the
    'not served by the wavefront kernel' answer come back without touching a device"""
    assert ctypes.sizeof(_native.Fused) == 3 * ctypes.sizeof(ctypes.c_void_p) + 2 * ctypes.sizeof(ctypes.c_int) or \
        ctypes.sizeof(_native.Fused) == 4 * ctypes.sizeof(ctypes.c_void_p)            # 3 pointers + int (+ padding)
    header = open(os.path.join(ROOT, "include", "ifk.h")).read()
    struct = header[header.index("typedef struct ifk_fused"):header.index("} ifk_fused;")]
    assert re.findall(r"(\w+)[;,]", struct.replace("*", "")) == ["in_scale", "in_bias", "out_scale", "squeeze"]
    dummy = ctypes.c_void_p(16)
    ok = _native.problem(2, 12, 16, 16, 3, 3, 12, 1)
    f = _native.Fused(None, None, None, 1)
    assert lib.ifk_inverse_fused_f32(ctypes.byref(ok), None, dummy, dummy, dummy, None) == -1          # no ifk_fused
    assert lib.ifk_inverse_fused_f32(ctypes.byref(ok), ctypes.byref(f), None, dummy, dummy, None) == -1
    assert lib.ifk_bwd_input_fused_f32(ctypes.byref(ok), ctypes.byref(f), dummy, dummy, dummy, None, None) == -1   # dz is required
    bad = _native.Fused(None, None, None, 2)
    assert lib.ifk_inverse_fused_f32(ctypes.byref(ok), ctypes.byref(bad), dummy, dummy, dummy, None) == -8
    # squeeze needs (C / groups) % 4 == 0
    g6 = _native.problem(2, 12, 16, 16, 3, 3, 12, 2)
    assert lib.ifk_inverse_fused_f32(ctypes.byref(g6), ctypes.byref(f), dummy, dummy, dummy, None) == -2
    # geometries without a pipelined wavefront variant: the caller runs the unfused sequence
    mnist = _native.problem(2, 4, 14, 14, 2, 2, 4, 1)
    assert lib.ifk_inverse_fused_f32(ctypes.byref(mnist), ctypes.byref(f), dummy, dummy, dummy, None) == _native.ERR_UNSUPPORTED
    empty = _native.problem(0, 12, 16, 16, 3, 3, 12, 1)
    assert lib.ifk_inverse_fused_f32(ctypes.byref(empty), ctypes.byref(f), None, dummy, None, None) == 0


def test_split_kernel_is_opt_in_and_limited_to_the_rows_it_holds(lib, monkeypatch):
    d = lambda *a: _native.describe_solve(_native.problem(*a))
    assert d(100, 12, 16, 16, 3, 3, 12, 1).startswith("wave<")
    monkeypatch.setenv("IFK_SOLVE_SPLIT", "1")
    assert d(100, 12, 16, 16, 3, 3, 12, 1).startswith("split<") and "helper cc=6 ns=4" in d(100, 12, 16, 16, 3, 3, 12, 1)
    assert d(100, 24, 8, 8, 3, 3, 24, 1).startswith("split<")
    assert d(100, 12, 32, 32, 3, 3, 12, 1).startswith("wave<")          # 32 rows: more than the split kernel's 16 slots
    assert d(100, 48, 4, 4, 3, 3, 48, 1).startswith("wave<")            # no split variant for Cg = 48
    monkeypatch.setenv("IFK_SPLIT_CFG", "8")
    assert "ns=8" in d(100, 12, 16, 16, 3, 3, 12, 1)
    # the prepared buffer grows by the split kernel's packed copy
    p = _native.problem(1, 12, 16, 16, 3, 3, 12, 1)
    with_split = lib.ifk_prepared_floats(ctypes.byref(p))
    monkeypatch.setenv("IFK_SOLVE_SPLIT", "0")
    assert lib.ifk_prepared_floats(ctypes.byref(p)) < with_split


def test_library_override_fails_loudly_when_the_file_is_missing():
    """IFK_LIBRARY (development: another build of the same ABI, tools/variant_build.py) replaces the in-tree path; a
    missing file raises -- there is no fallback to the default build, let alone to a CPU path"""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from inverse_flow_b200 import _native\n"
            "try:\n    _native.load()\nexcept _native.IfkError as e:\n    print('IfkError', 'not found' in str(e))\n" % ROOT)
    env = dict(os.environ, IFK_LIBRARY="/nonexistent/libifk_variant.so")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.stdout.strip().endswith("IfkError True"), (out.stdout, out.stderr[-500:])
    env = dict(os.environ, IFK_LIBRARY=_native.LIB_PATH)
    code_ok = ("import sys; sys.path.insert(0, %r)\n"
               "from inverse_flow_b200 import _native\nprint(_native.load().ifk_version())\n" % ROOT)
    out = subprocess.run([sys.executable, "-c", code_ok], env=env, capture_output=True, text=True, timeout=300)
    assert out.stdout.strip().isdigit(), (out.stdout, out.stderr[-500:])
