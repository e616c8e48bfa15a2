"""Oracle golden vectors (tests/golden/oracle_golden.npz, made by tests/golden/make_oracle_golden.py).

They are outputs of the NumPy ORACLE, not of the MATLAB reference (which could not be run: parity unpinned).  The
CPU test pins the oracle against drift; the GPU test compares the CUDA library with the committed vectors without
re-running the oracle."""
import importlib.util
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _maker():
    spec = importlib.util.spec_from_file_location("make_oracle_golden", os.path.join(HERE, "golden", "make_oracle_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _golden():
    return np.load(os.path.join(HERE, "golden", "oracle_golden.npz"))


def _err(a, b):
    from twoace_b200 import harness as hz
    return hz.aligned_rel_err(a, b)


def test_oracle_reproduces_its_golden_vectors():
    mk, g = _maker(), _golden()
    for name, (kind, A, B, tx, rx, tr, it) in mk.cases().items():
        X, Y, q = mk.run_oracle(kind, A, B, tx, rx, tr, it)
        assert _err(X, g[name + "/X"]) < 1e-7, name
        if kind != "PHASELIFT":
            assert abs(q - float(g[name + "/quality"])) < 1e-7, name
            assert Y.shape == g[name + "/Y"].shape, name


@pytest.mark.gpu
def test_cuda_library_matches_the_golden_vectors(gpu_ctx):
    import twoace_b200 as tw
    mk, g = _maker(), _golden()
    for name, (kind, A, B, tx, rx, tr, it) in mk.cases().items():
        if kind == "PHASELIFT":
            sig, _ = tw.phaselift_batch([A], [B], tw.PlOpts.default(maxIts=it), gpu_ctx)
            assert _err(sig[0], g[name + "/X"]) < 1e-8, name
            continue
        p = tw.Params.default(maxiter=it).fixed_iters()
        res = tw.solve_batch(getattr(tw, kind), [A], [B], tx, rx, [tr], p, gpu_ctx)
        assert _err(res.X[0], g[name + "/X"]) < 1e-6, name
        assert abs(res.quality[0] - float(g[name + "/quality"])) < 1e-6, name
        assert len(res.Y[0]) == len(g[name + "/Y"]), name
