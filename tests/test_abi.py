"""The C-ABI library loads without a GPU, exports every symbol include/twoace.h declares, and the
product path fails loudly (no CPU fallback) when no CUDA device exists."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    import twoace_b200
    return twoace_b200


def _declared():
    src = open(os.path.join(ROOT, "include", "twoace.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(twoace_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(built):
    lib = ctypes.CDLL(built.lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in twoace.h but not exported"
    assert sorted(built.lib.EXPORTS) == names, "lib.py EXPORTS out of sync with twoace.h"


def test_default_params_match_reference_defaults(built):
    lib = built.lib.load()
    p = built.lib.Params()
    lib.twoace_default_params(ctypes.byref(p))
    # inferLowRankV4.m:2-9
    assert (p.lam, p.r, p.mu0, p.rho, p.cc_frac, p.tol_rel, p.tol_abs, p.maxiter) == \
        (0.0, 20, 1e-3, 1.03, 0.95, 1e-4, 1e-8, 500)
    assert lib.twoace_version() >= 100


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(built.TwoaceError):
        built.Context(0)
    import numpy as np
    with pytest.raises(built.TwoaceError):
        built.inferLowRankV4(np.ones((8, 16), complex), np.ones(8), 4, 4, train_idx=np.arange(7)[None])


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "2ace-mmwave-channel-estimation_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                s = open(os.path.join(dp, f)).read()
                assert "import oracle" not in s and "from oracle" not in s, f
