"""Parity bundles: the .mat exchange format between this repository and a MATLAB run of the
unmodified reference (tools/matlab/run_bundle.m, tools/pin_parity.py)."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _pin():
    spec = importlib.util.spec_from_file_location("pin_parity", os.path.join(ROOT, "tools", "pin_parity.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_bundle_round_trip_and_compare(tmp_path, codebook):
    from scipy.io import savemat
    import twoace_b200 as tw
    from oracle import admm
    hz = tw.harness
    insts = hz.make_batch(2, codebook, 36, 20.0)
    A = [i.A for i in insts] + [insts[0].A]
    B = [i.B for i in insts] + [(insts[0].B / 2) ** 2]
    tr = [insts[0].train_idx[:1], insts[1].train_idx[:3], np.zeros((0, 0), np.int32)]
    solver = ["inferLowRankV4", "inferLowRankV4_multi", "MyPhaseLift"]
    path = str(tmp_path / "bundle.mat")
    p = tw.Params.default(maxiter=15)
    hz.export_bundle(path, A, B, 16, 16, solver, tr, p)
    S = hz.import_bundle(path)
    assert S["solver"] == solver and list(S["tx"]) == [16, 16, 16]
    for b in range(3):
        assert np.array_equal(S["A"][b], A[b]) and np.array_equal(S["B"][b], B[b])
        assert np.array_equal(S["train_idx"][b], tr[b])               # 1-based on disk, 0-based here
    assert S["params"][0, 7] == 15 and S["params"][0, 1] == 20
    # a stand-in for MATLAB's result file (same layout as run_bundle.m's save): the oracle's own answers
    pin = _pin()
    S["A"], S["B"], S["train_idx"], S["solver"] = S["A"][:2], S["B"][:2], S["train_idx"][:2], S["solver"][:2]
    X, q = pin.solve_bundle(S, "oracle")
    Xc = np.empty((2, 1), dtype=object)
    Yc = np.empty((2, 1), dtype=object)
    for b in range(2):
        Xc[b, 0] = X[b].reshape(-1, 1) * np.exp(0.3j)                  # MATLAB's eig may return another phase
        Yc[b, 0] = np.zeros((1, 1), complex)
    res = str(tmp_path / "bundle_ref.mat")
    savemat(res, {"X": Xc, "Y": Yc, "quality": q.reshape(-1, 1), "matlab_version": "fake"}, format="5")
    ref = hz.import_reference_results(res)
    errs, dq = pin.compare(S, ref, X, q)
    assert errs.max() < 1e-12 and dq.max() == 0.0
    Xo, _, qo = admm.infer_low_rank_v4(A[0], B[0], 16, 16, admm.Params(maxiter=15), train_idx=tr[0][0])
    # (the .mat round trip returns Fortran-ordered arrays: BLAS may sum in another order)
    assert np.linalg.norm(Xo - X[0]) < 1e-9 * np.linalg.norm(Xo) and abs(qo - q[0]) < 1e-9


def test_matlab_replay_scripts_cover_every_solver_name():
    """run_bundle.m dispatches on the same solver names export_bundle documents."""
    src = open(os.path.join(ROOT, "tools", "matlab", "run_bundle.m")).read()
    for name in ("inferLowRankV4", "inferLowRankV4_multi", "inferLowRank_Nuclear", "inferLowRankV3",
                 "inferLowRankV2", "inferLowRank", "MyPhaseLift"):
        assert "'" + name + "'" in src
    assert "TWOACE_DRAWS" in open(os.path.join(ROOT, "tools", "matlab", "randsample.m")).read()
