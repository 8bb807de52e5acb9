import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0]
                  for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        d = {k: z[k] for k in z.files}
    d["groups"] = int(d["groups"])
    return d


@pytest.fixture(params=golden_names())
def golden(request):
    d = load_golden(request.param)
    d["name"] = request.param
    return d


def make_weight(rng, C, Cw, KH, KW, scale=0.05, dtype=np.float32):
    """inv_flow_*.reset_parameters-shaped weight (reference inf/layers/inv_conv.py:153-170):
    identity at the last tap + noise, W[c,-1,-1,-1] = 1; every column populated."""
    w = rng.standard_normal((C, Cw, KH, KW)) * scale
    for c in range(min(C, Cw)):
        w[c, c, KH - 1, KW - 1] += 1.0
    w[:, -1, -1, -1] = 1.0
    return w.astype(dtype)


class _KnobPatch:
    """monkeypatch whose setenv / delenv of IFK_* knobs also tells the library to re-read them: the
    library parses its environment once and caches it (include/ifk.h, ifk_debug_reload_env)."""

    def __init__(self, inner):
        self._inner = inner

    def _reload(self, name):
        if name.startswith("IFK_"):
            from inverse_flow_b200 import _native
            _native.reload_env()

    def setenv(self, name, value, *a, **k):
        self._inner.setenv(name, value, *a, **k)
        self._reload(name)

    def delenv(self, name, *a, **k):
        self._inner.delenv(name, *a, **k)
        self._reload(name)

    def __getattr__(self, attr):
        return getattr(self._inner, attr)


@pytest.fixture
def monkeypatch(monkeypatch):
    patch = _KnobPatch(monkeypatch)
    yield patch
    monkeypatch.undo()
    patch._reload("IFK_")
