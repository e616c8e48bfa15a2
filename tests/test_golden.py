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


# ---- round-2 components (tests/golden/oracle_golden_r2.npz, made by tests/golden/make_oracle_golden_r2.py) ------------
def _maker_r2():
    spec = importlib.util.spec_from_file_location("make_oracle_golden_r2", os.path.join(HERE, "golden", "make_oracle_golden_r2.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


def test_oracle_reproduces_its_round2_golden_vectors():
    g = np.load(os.path.join(HERE, "golden", "oracle_golden_r2.npz"))
    now = _maker_r2().compute()
    assert set(now) == set(g.files)
    for k in g.files:
        if np.issubdtype(g[k].dtype, np.integer):
            assert np.array_equal(now[k], g[k]), k            # index bookkeeping: bit-exact
        else:
            assert _rel(now[k], g[k]) < 1e-9, k


@pytest.mark.gpu
def test_cuda_library_matches_the_round2_golden_vectors(gpu_ctx):
    import twoace_b200 as tw
    from twoace_b200 import harness as hz, solvers as sv
    mk = _maker_r2()
    g = np.load(os.path.join(HERE, "golden", "oracle_golden_r2.npz"))
    cb = hz.load_codebook()
    gpu_ctx.set_codebook(cb)
    sp = tw.SynthParams.default(ntrain=3)
    m, snr, lo, hi, tid = zip(*mk.SYNTH_CASES)
    out = sv.synth_batch(list(m), list(snr), list(lo), list(hi), list(tid), sp, gpu_ctx)
    for k in range(len(m)):
        assert np.array_equal(out["rows"][k], g[f"synth{k}/rows"]) and np.array_equal(out["train_idx"][k], g[f"synth{k}/train_idx"])
        assert _rel(out["B"][k], g[f"synth{k}/B"]) < 1e-12 and _rel(out["vecH"][k], g[f"synth{k}/vecH"]) < 1e-12
        assert np.allclose(out["angles"][k], np.concatenate([g[f"synth{k}/aod"], g[f"synth{k}/aoa"]]), atol=1e-12, rtol=0)
    ang = np.concatenate([g["synth0/aod"], g["synth0/aoa"]])[None, :]
    a = sv.angle_evaluation_batch(np.stack([g["synth0/vecH"], g["angles/x_noisy"]]), np.repeat(ang, 2, axis=0), 16, 16, ctx=gpu_ctx)
    assert np.allclose(a[0], g["angles/exact"], atol=1e-9) and np.allclose(a[1], g["angles/noisy"], atol=1e-9)
    ins = hz.make_batch(1, cb, 361, 20.0)[0]
    from oracle import admm          # (pre-processing and the spectral start point only; the iterates come from the file)
    A, B, _, _ = admm._preprocess(ins.A, ins.B, 1e-8)
    tr = g["stage_M361/train_idx"]
    At, Bt = A[tr], B[tr]
    X0 = admm.spectral_initialize(At, Bt, 20)
    p = tw.Params.default(maxiter=5, tol_rel=0.0, tol_abs=0.0)
    _, _, S, _ = sv.infer_admm_batch([At], [Bt], [X0], True, False, 16, 16, p, nuclear=2, ctx=gpu_ctx)
    assert _rel(S[0]["X"], g["minl2_M361_it5/X"]) < 1e-9 and _rel(S[0]["Y"], g["minl2_M361_it5/Y"]) < 1e-9
    _, _, S, _ = sv.infer_admm_batch([At], [Bt], [X0], True, False, 16, 16, p, ctx=gpu_ctx)
    assert _rel(S[0]["X"], g["v4_M361_it5/X"]) < 1e-9 and _rel(S[0]["Z"], g["v4_M361_it5/Z"]) < 1e-9
    from twoace_b200 import twostage as ts
    P, C, mcs = ts.svd_reduction(ins.A[:121] @ (np.eye(256)[:, ::2]), 3)
    assert mcs == int(g["twostage/mCS"]) and _rel(P @ C, g["twostage/PC"]) < 1e-9
