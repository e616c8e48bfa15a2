"""inferMinL2 (ADMM_v2.m version 0; ADMM_v2/inferMinL2.m) on the GPU against the oracle restatement: per-stage iterate
parity of its InferADMM (no low-rank variable, X = pinv(A)(Y - M/mu)), the 90 %-energy rank rule of its spectral
initialisation, and full solves through the MATLAB-signature wrapper."""
import math

import numpy as np
import pytest

from oracle import admm

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


def _case(codebook, M, trial=0):
    from twoace_b200 import harness as hz
    ins = hz.make_batch(trial + 1, codebook, M, 20.0)[trial]
    A, B, _, _ = admm._preprocess(ins.A, ins.B, 1e-8)
    return ins, A, B


@pytest.mark.parametrize("M,r,sbr,iters", [(64, 20, True, 5), (64, 7, False, 5), (361, 20, True, 1), (361, 20, True, 40),
                                           (361, 6, False, 40), (529, 1, True, 60), (1024, 20, False, 25)])
def test_minl2_stage_parity(codebook, gpu_ctx, M, r, sbr, iters):
    import twoace_b200 as tw
    from twoace_b200 import solvers as sv
    _, A, B = _case(codebook, M)
    X0 = admm.spectral_initialize(A, B, 20)[:, :r]
    snap = {iters: None}
    tro = admm.StageTrace()
    Xo, Yo, _ = admm.infer_admm_minl2(A, B, X0, sbr, 0.0, 0.0, 0.0, iters, tro, snap)
    p = tw.Params.default(maxiter=iters, tol_rel=0.0, tol_abs=0.0)
    Xg, Yg, Sg, W = sv.infer_admm_batch([A], [B], [X0], sbr, False, 16, 16, p, nuclear=2, ctx=gpu_ctx)
    s = snap[iters]
    # pinv(A) is formed through the smaller Gram matrix here and by SVD in the oracle: cond(A)^2 ~ 2e2 at m = 342
    tol = 1e-7
    assert rel(Sg[0]["X"], s["X"]) < tol and rel(Sg[0]["Y"], s["Y"]) < tol
    assert np.linalg.norm(Sg[0]["M"] - s["M"]) < tol * max(1.0, np.linalg.norm(s["M"]))
    assert int(W[0][2]) == iters
    if A.shape[0] > A.shape[1]:
        # (for m <= n every column fits the magnitudes exactly after one step -- objective ~1e-16 -- so the best iterate
        # and best column of :296-312 are decided by rounding noise in the reference itself)
        assert rel(Xg[0], Xo) < tol and rel(Yg[0], Yo) < tol
        assert int(W[0][3]) == tro.opt_iter and int(W[0][4]) == tro.opt_col
        assert int(W[0][5]) == tro.n_mu_bumps and abs(W[0][0] - tro.mu) <= 1e-12 * tro.mu


def test_minl2_is_stationary_after_one_step_when_m_le_n(codebook, gpu_ctx):
    """m <= n: A pinv(A) = I, so A X = Y - M/mu exactly, Y does not move and the stopping test fires at iteration 1
    (the reference behaves the same way: inferMinL2.m:282-330)."""
    import twoace_b200 as tw
    from twoace_b200 import solvers as sv
    _, A, B = _case(codebook, 64)
    X0 = admm.spectral_initialize(A, B, 20)
    tro = admm.StageTrace()
    Xo, _, conv = admm.infer_admm_minl2(A, B, X0, True, 0.0, 1e-4, 1e-8, 500, tro)
    Xg, _, _, W = sv.infer_admm_batch([A], [B], [X0], True, False, 16, 16, tw.Params.default(), nuclear=2, ctx=gpu_ctx)
    assert tro.iters == 1 and conv and int(W[0][2]) == 1 and bool(W[0][6])
    assert rel(Xg[0], Xo) < 1e-10


def _oracle_stage_b_columns(A, B, tr, finish=False):
    """All columns of the oracle's parallel-refinement iterate (inferMinL2.m:222-224), scaled like the output.  With
    `finish` every column is carried through the rest of inferMinL2 (:42-58: held-out quality, refinement on all rows if
    quality > 0.6, roll-back), i.e. the candidates are what inferMinL2 returns for each possible choice of min()."""
    A = np.asarray(A, dtype=np.complex128)
    m, n = A.shape
    An, Bn, A_norm, B_norm = admm._preprocess(A, B, 1e-8)
    At, Bt = An[tr], Bn[tr]
    X = admm.spectral_initialize_minl2(At, Bt, min(20, m, n))
    X, _, _ = admm.infer_admm_minl2(At, Bt, X, True, 0.0, 1e-4, 1e-8, 500)
    G = X.conj().T @ X
    X = X @ np.linalg.eigh(0.5 * (G + G.conj().T))[1]
    snap = {1: None}
    admm.infer_admm_minl2(At, Bt, X, False, 0.0, 1e-4, 1e-8, 500, None, snap)
    cols = snap[1]["X"]
    if finish:
        te = admm.test_index_set(m, tr)
        out = []
        for c in range(cols.shape[1]):
            x0 = cols[:, c:c + 1]
            q = admm.quality_score(An[te], Bn[te], x0) if te.size else float("nan")
            x = x0
            if q > 0.6:
                x, _, _ = admm.infer_admm_minl2(An, Bn, x0, True, 0.0, 1e-4, 1e-8, 500)
                if abs(np.vdot(x0, x)) / np.linalg.norm(x0) / np.linalg.norm(x) < 0.6:
                    x = x0
            out.append((x.reshape(-1) * (B_norm / A_norm), q))
        return out
    return cols * (B_norm / A_norm)


@pytest.mark.parametrize("M", [10, 36, 64, 225, 361, 529])
def test_inferMinL2_full_solve(codebook, gpu_ctx, M):
    import twoace_b200 as tw
    from twoace_b200 import harness as hz
    insts = hz.make_batch(4, codebook, M, 20.0)
    rng = np.random.default_rng(3)
    for ins in insts:
        tr = rng.permutation(M)[:math.ceil(M * 0.95)].astype(np.int32)
        info = admm.SolveInfo()
        Xo, Yo, qo = admm.infer_min_l2(ins.A, ins.B, train_idx=tr, info=info)
        Xg, Yg, qg = tw.inferMinL2(ins.A, ins.B, train_idx=tr, ctx=gpu_ctx)
        if M > 256:
            # (pinv through the Gram matrix vs SVD, cond(A)^2 ~ 1e2, over a few hundred iterations)
            # The column-wise iteration of this solver (no low-rank coupling between the columns) amplifies rounding
            # differences exponentially: from identical start points the GPU and the oracle agree to 1e-13 after 10
            # iterations, 1e-11 ... 1e-7 after 50 and up to 5e-5 after 100, column by column, with identical iteration
            # counts, best columns and mu (profiles/r02_minl2_stage_deviation_growth_M529.log; tight per-stage bars in
            # test_minl2_stage_parity).  After the ~200 iterations of a full solve the per-instance bar is therefore 1e-2,
            # with bit-exact bookkeeping.
            assert Yg.shape == Yo.shape
            assert abs(qg - qo) < 1e-2, (qg, qo)
            assert hz.aligned_rel_err(Xg, Xo) < 1e-2
            from twoace_b200 import solvers as sv
            sw = sv.solve_batch(tw.MINL2, [ins.A], [ins.B], 16, 16, [tr], tw.Params.default(), gpu_ctx).stage_words[0]
            got = [(int(sw[k, 2]), int(sw[k, 3]), int(sw[k, 4]), int(sw[k, 6])) for k in (0, 1)]
            want = [(t.iters, t.opt_iter, t.opt_col, int(t.converged)) for t in info.traces[:2]]
            assert got == want          # iteration counts, best iterate, best column, convergence flag: bit-exact
        else:
            # m_train <= n: after one step EVERY column of the parallel refinement fits the train magnitudes exactly
            # (objective ~1e-16), so which column min() returns (:304-310) is decided by rounding noise in the reference
            # itself; the GPU must return what inferMinL2 returns for ONE of the oracle's columns (held-out quality,
            # refinement if that quality exceeds 0.6, roll-back), with that column's quality
            cands = _oracle_stage_b_columns(ins.A, ins.B, tr, finish=True)
            errs = [hz.aligned_rel_err(Xg, x) for x, _ in cands]
            assert min(errs) < 1e-6, errs
            qc = cands[int(np.argmin(errs))][1]
            assert (np.isnan(qc) and np.isnan(qg)) or abs(qg - qc) < 1e-7
    # version switch row 0 (ADMM_v2.m:23)
    X0, _, q0 = tw.ADMM_v2(insts[0].B, insts[0].A, 16, 16, 0, train_idx=tr, ctx=gpu_ctx)
    assert X0.shape == (256,)


def test_minl2_rank_rule(gpu_ctx):
    """inferMinL2.m:181-185 with 16 rows: 16 nonzero eigenvalues, the rule fires and cuts the columns down; the test
    set is empty (ceil(0.95 * 16) = 16), so quality = 1 - 0/0 = NaN and the refinement is skipped, as in MATLAB."""
    import twoace_b200 as tw
    rng = np.random.default_rng(0)
    n, m = 64, 16
    A = (rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n))) / np.sqrt(2)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    B = np.abs(A @ x)
    tr = rng.permutation(m)[:math.ceil(m * 0.95)].astype(np.int32)
    info = admm.SolveInfo()
    Xo, _, qo = admm.infer_min_l2(A, B, train_idx=tr, info=info)
    Xg, _, qg = tw.inferMinL2(A, B, train_idx=tr, ctx=gpu_ctx)
    assert np.isnan(qo) and np.isnan(qg)
    assert admm.spectral_initialize_minl2(*admm._preprocess(A, B, 1e-8)[:2], 16).shape[1] < 16
    from twoace_b200 import harness as hz
    cols = _oracle_stage_b_columns(A, B, tr)        # m <= n: the returned column is noise-decided (see above)
    assert cols.shape[1] < 16
    assert min(hz.aligned_rel_err(Xg, cols[:, c]) for c in range(cols.shape[1])) < 1e-7
