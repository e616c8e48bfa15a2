"""CPU oracle for the 2ACE ADMM phase-retrieval / low-rank CSI recovery loop.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path imports this module: it
is used by ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` as the checker and the CPU baseline.

PARITY UNPINNED: the reference (MATLAB) ships no golden vectors, no
known-answer tests and no seeds reproducible outside MATLAB for this path, and
neither MATLAB nor Octave exists in the build container.  The
oracle is therefore a line-by-line NumPy complex128 restatement of the ``.m``
files, anchored on (i) the authors' commented self-test recipe
(``main/src/my_recovery_algorithms/ADMM_v2.m:13-19,47-48``), (ii) analytic
identities of every sub-step and (iii) the held-out "quality" bar of
``inferLowRankV4.m:59,68``; see ``tests/test_oracle_*.py``.

Reference files restated here (all under
``main/src/my_recovery_algorithms/ADMM_v2/`` of the reference tree):

* ``inferLowRankV4.m``        -> :func:`infer_low_rank_v4`
* ``inferLowRankV4_multi.m``  -> :func:`infer_low_rank_v4_multi`
* ``inferLowRank_Nuclear.m``  -> :func:`infer_low_rank_nuclear`
* ``../ADMM_v2.m`` / ``../ADMM_v2_nuclear.m`` -> :func:`admm_v2`

The MATLAB global RNG is externalised: the train index sets that the reference
draws with ``randsample`` (``inferLowRankV4.m:37``) are explicit inputs
(0-based, in the drawn order).  Everything else is deterministic.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

__all__ = [
    "Params", "normalize_rows", "argmin_y", "argmin_x", "rank_profile",
    "argmin_z", "argmin_z_nuclear", "spectral_initialize", "infer_admm",
    "infer_low_rank_impl", "infer_low_rank_v4", "infer_low_rank_v4_multi",
    "infer_low_rank_nuclear", "admm_v2", "test_index_set", "quality_score",
]


@dataclass
class Params:
    """Optional positional arguments of inferLowRankV4.m:1-9 with their nargin defaults."""
    lam: float = 0.0
    r: int = 20
    mu0: float = 1e-3
    rho: float = 1.03
    cc_frac: float = 0.95
    tol_rel: float = 1e-4
    tol_abs: float = 1e-8
    maxiter: int = 500

    def fixed_iters(self) -> "Params":
        """Fixed-iteration mode of SURVEY.md §8(d): thresholds 0 so :351 never fires."""
        return Params(self.lam, self.r, self.mu0, self.rho, self.cc_frac, 0.0, 0.0, self.maxiter)


@dataclass
class StageTrace:
    """Book-keeping of one InferADMM call (not returned by the reference; used for parity)."""
    iters: int = 0
    converged: bool = False
    mu: float = 0.0
    opt_obj: float = math.inf
    opt_iter: int = -1
    opt_col: int = -1
    n_mu_bumps: int = 0
    res_comb: list = field(default_factory=list)
    # conditioning of the column orthonormalisation that produced this stage's start point
    # (lambda_min / lambda_max of X'*X at inferLowRankV4.m:242; NaN when not applicable)
    ortho_cond: float = float("nan")


def _fro(x) -> float:
    return float(np.sqrt(np.sum(np.abs(x) ** 2)))


# ----------------------------------------------------------------------------
# inferLowRankV4.m:517-538
def normalize_rows(Y, B, scale_by_row):
    Y = np.array(Y, dtype=np.complex128, copy=True)
    B = np.asarray(B, dtype=np.float64).reshape(-1)
    r = Y.shape[1]
    if scale_by_row:
        D = np.sqrt(np.sum(np.abs(Y) ** 2, axis=1))
        I = D == 0
        if I.any():
            Y[I, :] = 1.0 / math.sqrt(r)
            D[I] = 1.0
        return Y * (B / D)[:, None]
    D = np.abs(Y)
    I = D == 0
    if I.any():
        Y[I] = 1.0
        D[I] = 1.0
    return Y * (B[:, None] / D)


# inferLowRankV4.m:490-512
def argmin_y(AX, B, M, mu, scale_by_row):
    B = np.asarray(B, dtype=np.float64).reshape(-1)
    Y = AX + M / mu
    r = Y.shape[1]
    if scale_by_row:
        D = np.sqrt(np.sum(np.abs(Y) ** 2, axis=1))
        I = D == 0
        if I.any():
            Y[I, :] = 1.0 / math.sqrt(r)
            D[I] = 1.0
        BD = B / D
        return Y * ((BD + mu) / (1 + mu))[:, None]
    D = np.abs(Y)
    I = D == 0
    if I.any():
        Y[I] = 1.0
        D[I] = 1.0
    BD = B[:, None] / D
    return Y * ((BD + mu) / (1 + mu))


# inferLowRankV4.m:380-388
def argmin_x(A, Y, Z, M, N, mu, lam, U, D):
    rhs = A.conj().T @ (Y - M / mu) + (Z - N / mu)
    if lam == 0:
        return U @ rhs
    D_inv = 1.0 / (D + (1 + lam / mu))
    return U @ (D_inv[:, None] * (U.conj().T @ rhs))


# inferLowRankV4.m:416-443
def rank_profile(tx, rx, m, n, use_rank_one):
    """inferLowRankV4.m:416-443.  ``use_rank_one``: False/True as in V4, or the profile code of an older
    version: 2 = inferLowRank.m:407-418,437 (single stage [r2]), 3 = inferLowRankV2.m:418-431 (as V3/V4 except
    the small-array fallback [r2 r3])."""
    sz = min(rx, tx)
    r0 = math.ceil(math.sqrt(sz) * 0.5)
    r1 = math.ceil(math.sqrt(sz) * 0.7)
    r2 = math.ceil(math.sqrt(sz))
    r3 = min(sz, math.ceil(math.sqrt(sz) * 2.0))
    f0, f1, f2, f3 = 0.8, 0.9, 0.95, 0.995
    if use_rank_one == 1:
        return [1], [0.95]
    if use_rank_one == 2:
        return [min(tx, rx, r2)], [f2]
    if m >= n * 3:
        return [r3], [f3]
    if r1 <= 2:
        return ([r2, r3], [f2, f3]) if use_rank_one == 3 else ([r2], [f2])
    if r0 <= 2:
        return [r1, r2, r3], [f1, f2, f3]
    return [r0, r1, r2, r3], [f0, f1, f2, f3]


def _eigh_desc(G):
    """MATLAB eig of an exactly-Hermitian matrix (ascending) followed by
    max(0,real(.)) and a stable descending sort (inferLowRankV4.m:407-409)."""
    G = 0.5 * (G + G.conj().T)
    w, V = np.linalg.eigh(G)
    s2 = np.maximum(0.0, w)
    idx = np.argsort(-s2, kind="stable")
    return s2[idx], idx, V


# inferLowRankV4.m:402-464
def argmin_z(X, N, mu, tx, rx, m, n, use_rank_one):
    Z = X + N / mu
    ncol = Z.shape[1]
    E = Z.reshape(tx, -1, order="F")
    s2, idx, U = _eigh_desc(E @ E.conj().T)
    r_list, f_list = rank_profile(tx, rx, m, n, use_rank_one)
    s2 = s2.copy()
    s2_scale = np.ones_like(s2)
    for r, f in zip(r_list, f_list):
        vr = float(np.sum(s2[:r]))
        v = float(np.sum(s2))
        if vr < v * f:
            scale = min(1.0, vr / (v - vr) * (1.0 / f - 1.0))
            s2[r:] = s2[r:] * scale
            s2_scale[idx[r:]] = s2_scale[idx[r:]] * scale
    if np.any(s2_scale < 1):
        P = (U * np.sqrt(s2_scale)[None, :]) @ U.conj().T
        Z = (P @ E).reshape(tx * rx, ncol, order="F")
    return Z


# inferLowRank_Nuclear.m:411-439 (Shrink with soft=1 on the singular values)
def argmin_z_nuclear(X, N, mu, tx, rx, m, n, use_rank_one):
    Z = X + N / mu
    U, s, Vh = np.linalg.svd(Z, full_matrices=False)
    s = np.sign(s) * np.maximum(0.0, np.abs(s) - 1.0 / mu)
    return (U * s[None, :]) @ Vh


# inferLowRankV4.m:540-553
def spectral_initialize(A, B, r):
    A = np.asarray(A, dtype=np.complex128)
    B = np.asarray(B, dtype=np.float64).reshape(-1)
    As = A.copy()
    an = np.sqrt(np.sum(np.abs(A) ** 2, axis=1))
    nz = an != 0
    As[nz, :] = A[nz, :] * (B[nz] / an[nz])[:, None]
    AtA = As.conj().T @ As
    s2, idx, V = _eigh_desc(AtA)
    return V[:, idx[:r]] * np.sqrt(s2[:r])[None, :]


def _min_skip_nan(objs):
    """MATLAB [obj,j]=min(objs): first minimiser, NaN skipped (all-NaN -> NaN, index 0)."""
    ok = ~np.isnan(objs)
    if not ok.any():
        return float("nan"), 0
    j = int(np.nanargmin(objs))
    return float(objs[j]), j


# inferLowRankV4.m:260-365
def infer_admm(A, B, X0, scale_by_row, use_rank_one, tx, rx, lam, mu0, rho,
               tol_rel, tol_abs, maxiter, U=None, D=None, argmin_z_fn=argmin_z,
               trace: StageTrace | None = None, snapshot_iters=None):
    """One InferADMM call.  Returns (X, Y, converged).

    ``snapshot_iters`` (test hook, not in the reference): dict iteration ->
    filled with a copy of the state after that iteration (1-based).
    """
    A = np.asarray(A, dtype=np.complex128)
    B = np.asarray(B, dtype=np.float64).reshape(-1)
    X0 = np.asarray(X0, dtype=np.complex128)
    if X0.ndim == 1:
        X0 = X0[:, None]
    m, n = A.shape
    r = X0.shape[1]
    Ah = A.conj().T
    if U is None:
        if lam == 0:
            U = np.linalg.inv(Ah @ A + np.eye(n))
            D = None
        else:
            G = Ah @ A
            w, U = np.linalg.eigh(0.5 * (G + G.conj().T))
            D = np.maximum(0.0, w)

    M = np.zeros((m, r), dtype=np.complex128)
    N = np.zeros((n, r), dtype=np.complex128)
    X = X0.copy()
    AX = A @ X
    with np.errstate(divide="ignore", invalid="ignore"):
        if scale_by_row:
            X = X * (np.linalg.norm(B) / _fro(AX))
        else:
            cn = np.sqrt(np.sum(np.abs(AX) ** 2, axis=0))
            X = X * (np.linalg.norm(B) / cn)[None, :]
        AX = A @ X
        Y = normalize_rows(AX, B, scale_by_row)
        Z = argmin_z_fn(X, N, 1.0, tx, rx, m, n, use_rank_one)
        AtY = Ah @ Y

        mu = mu0
        opt_obj = math.inf
        converged = False
        last_res = math.inf
        opt_X = None
        opt_Y = None
        if trace is None:
            trace = StageTrace()

        for it in range(1, maxiter + 1):
            Y0, Z0, AtY0 = Y, Z, AtY
            X = argmin_x(A, Y, Z, M, N, mu, lam, U, D)
            AX = A @ X
            Y = argmin_y(AX, B, M, mu, scale_by_row)
            AtY = Ah @ Y
            Z = argmin_z_fn(X, N, mu, tx, rx, m, n, use_rank_one)
            J_M = AX - Y
            M = M + mu * J_M
            J_N = X - Z
            N = N + mu * J_N

            if scale_by_row:
                obj = float(np.linalg.norm(np.sqrt(np.sum(np.abs(AX) ** 2, axis=1)) - B))
                if obj < opt_obj:
                    opt_obj, opt_X, opt_Y = obj, X, Y
                    trace.opt_iter, trace.opt_col = it, -1
            else:
                objs = np.sqrt(np.sum((np.abs(AX) - B[:, None]) ** 2, axis=0))
                obj, j = _min_skip_nan(objs)
                if obj < opt_obj:
                    opt_obj, opt_X, opt_Y = obj, X[:, j:j + 1], Y[:, j:j + 1]
                    trace.opt_iter, trace.opt_col = it, j

            nJM, nJN = _fro(J_M), _fro(J_N)
            nZd, nYd = _fro(Z - Z0), _fro(Y - Y0)
            nAX, nY, nX, nZ = _fro(AX), _fro(Y), _fro(X), _fro(Z)
            nAtY = _fro(AtY)
            res_prim = math.sqrt(nJM ** 2 + nJN ** 2)
            res_dual = mu * math.sqrt(_fro(AtY - AtY0) ** 2 + nZd ** 2)
            res_comb = math.sqrt(res_prim ** 2 + nYd ** 2 + nZd ** 2)
            mx1, mx2 = max(nAX, nY), max(nX, nZ)
            thresh_prim = tol_abs * math.sqrt((m + n) * r) + tol_rel * math.sqrt(mx1 ** 2 + mx2 ** 2)
            thresh_dual = tol_abs * math.sqrt(n * r * 2) + tol_rel * math.sqrt(nAtY ** 2 + nZ ** 2)
            thresh_comb = tol_abs * math.sqrt((m + n) * r * 2) + tol_rel * math.sqrt(
                mx1 ** 2 + mx2 ** 2 + nY ** 2 + nZ ** 2)
            trace.iters = it
            trace.res_comb.append(res_comb)
            if snapshot_iters is not None and it in snapshot_iters:
                snapshot_iters[it] = dict(X=X.copy(), Y=Y.copy(), Z=Z.copy(), M=M.copy(), N=N.copy(),
                                          mu=mu, opt_obj=opt_obj, res_comb=res_comb)
            if (res_prim < thresh_prim and res_dual < thresh_dual) or (res_comb < thresh_comb):
                converged = True
                break
            if res_comb > last_res * 0.9:
                mu = mu * rho
                trace.n_mu_bumps += 1
            last_res = res_comb

    trace.converged = converged
    trace.mu = mu
    trace.opt_obj = opt_obj
    if opt_X is None:  # every obj was NaN (degenerate columns): MATLAB would error on
        # the undefined opt_X; the entry points turn NaN into 0 (A2only.m:176), so return NaN.
        ncol = r if scale_by_row else 1
        opt_X = np.full((n, ncol), np.nan + 0j)
        opt_Y = np.full((m, ncol), np.nan + 0j)
    return opt_X, opt_Y, converged


# inferLowRankV4.m:90-250
def infer_low_rank_impl(A, B, Xs, tx, rx, lam, r, mu0, rho, tol_rel, tol_abs, maxiter,
                        use_rank_one, argmin_z_fn=argmin_z, traces=None):
    A = np.asarray(A, dtype=np.complex128)
    m, n = A.shape
    r = min(r, m, n)
    Ah = A.conj().T
    if lam == 0:
        U = np.linalg.inv(Ah @ A + np.eye(n))
        D = None
    else:
        G = Ah @ A
        w, U = np.linalg.eigh(0.5 * (G + G.conj().T))
        D = np.maximum(0.0, w)
    tA, tB = StageTrace(), StageTrace()
    X, Y, _ = infer_admm(A, B, Xs, True, use_rank_one, tx, rx, lam, mu0, rho, tol_rel, tol_abs,
                         maxiter, U, D, argmin_z_fn, tA)
    G = X.conj().T @ X
    wx, Vx = np.linalg.eigh(0.5 * (G + G.conj().T))
    with np.errstate(divide="ignore", invalid="ignore"):
        tB.ortho_cond = float(max(wx[0], 0.0) / wx[-1]) if wx[-1] > 0 else 0.0
    X = X @ Vx
    X, Y, conv = infer_admm(A, B, X, False, use_rank_one, tx, rx, lam, mu0, rho, tol_rel, tol_abs,
                            maxiter, U, D, argmin_z_fn, tB)
    if traces is not None:
        traces.extend([tA, tB])
    return X, Y, conv


def test_index_set(m, train_idx):
    """setdiff(1:m, train_idx) (inferLowRankV4.m:38): sorted ascending, 0-based here."""
    mask = np.ones(m, dtype=bool)
    mask[np.asarray(train_idx, dtype=np.int64)] = False
    return np.nonzero(mask)[0]


test_index_set.__test__ = False  # not a pytest test


def quality_score(A_test, B_test, X):
    """inferLowRankV4.m:54."""
    with np.errstate(divide="ignore", invalid="ignore"):
        return float(1 - np.linalg.norm(np.abs(A_test @ X).reshape(-1) - B_test) / np.linalg.norm(B_test))


@dataclass
class SolveInfo:
    """Integer/flag bookkeeping of one solve (for the bit-exact part of the parity bar)."""
    quality: float = float("nan")
    used_rank_one: bool = False
    rolled_back: bool = False
    similarity: float = float("nan")
    best_trial: int = 0
    trial_quality: list = field(default_factory=list)
    trial_rank_one: list = field(default_factory=list)
    traces: list = field(default_factory=list)


def _preprocess(A, B, tol_abs):
    """inferLowRankV4.m:11-31."""
    A = np.asarray(A, dtype=np.complex128)
    B = np.asarray(B, dtype=np.float64).reshape(-1)
    m, n = A.shape
    A_norm = _fro(A) / math.sqrt(m)
    if A_norm < tol_abs:
        A_norm = 1.0
    B_norm = float(np.linalg.norm(B))
    if B_norm < tol_abs:
        B_norm = 1.0
    return A / A_norm, B / B_norm, A_norm, B_norm


def _train_solve(A, B, train_idx, tx, rx, p: Params, r, argmin_z_fn, info: SolveInfo):
    """inferLowRankV4.m:36-63 for one train/test split."""
    m = A.shape[0]
    train_idx = np.asarray(train_idx, dtype=np.int64)
    assert train_idx.size == int(math.floor(m * p.cc_frac)), "train_idx must have floor(m*cc_frac) entries"
    test_idx = test_index_set(m, train_idx)
    A_train, B_train = A[train_idx, :], B[train_idx]
    A_test, B_test = A[test_idx, :], B[test_idx]
    Xs = spectral_initialize(A_train, B_train, r)
    use_rank_one = False
    X, Y, _ = infer_low_rank_impl(A_train, B_train, Xs, tx, rx, p.lam, r, p.mu0, p.rho, p.tol_rel,
                                  p.tol_abs, p.maxiter, use_rank_one, argmin_z_fn, info.traces)
    quality = quality_score(A_test, B_test, X)
    if quality < 0.6:
        use_rank_one = True
        X, Y, _ = infer_low_rank_impl(A_train, B_train, Xs, tx, rx, p.lam, r, p.mu0, p.rho, p.tol_rel,
                                      p.tol_abs, p.maxiter, use_rank_one, argmin_z_fn, info.traces)
        quality = quality_score(A_test, B_test, X)
    return X, Y, quality, use_rank_one


def _refine(A, B, X_in, Y_in, quality, use_rank_one, tx, rx, p: Params, argmin_z_fn, info: SolveInfo):
    """inferLowRankV4.m:68-80."""
    t = StageTrace()
    if quality > 0.6:
        X0, Y0 = X_in, Y_in
        X, Y, _ = infer_admm(A, B, X0, True, use_rank_one, tx, rx, p.lam, p.mu0, p.rho, p.tol_rel,
                             p.tol_abs, p.maxiter, None, None, argmin_z_fn, t)
        with np.errstate(divide="ignore", invalid="ignore"):
            similarity = float(np.abs(np.vdot(X0, X)) / np.linalg.norm(X0) / np.linalg.norm(X))
        info.similarity = similarity
        if similarity < 0.6:
            X, Y = X0, Y0
            info.rolled_back = True
    else:
        X, Y, _ = infer_admm(A, B, X_in, True, use_rank_one, tx, rx, p.lam, p.mu0, p.rho, p.tol_rel,
                             p.tol_abs, p.maxiter, None, None, argmin_z_fn, t)
    info.traces.append(t)
    return X, Y


def _v4_like(A, B, tx, rx, p: Params, train_idx, argmin_z_fn, info: SolveInfo):
    A = np.asarray(A, dtype=np.complex128)
    m, n = A.shape
    r = min(p.r, m, n)
    A, B, A_norm, B_norm = _preprocess(A, B, p.tol_abs)
    X, Y, quality, use_rank_one = _train_solve(A, B, train_idx, tx, rx, p, r, argmin_z_fn, info)
    info.quality, info.used_rank_one = quality, use_rank_one
    info.trial_quality, info.trial_rank_one = [quality], [use_rank_one]
    X, Y = _refine(A, B, X, Y, quality, use_rank_one, tx, rx, p, argmin_z_fn, info)
    s = B_norm / A_norm
    return X.reshape(-1) * s, Y.reshape(-1) * s, quality


def _older_version(A, B, tx, rx, p: Params, train_idx, info: SolveInfo, profile: int, refine_only_if_good: bool):
    """inferLowRankV3.m:1-70 / inferLowRankV2.m:1-66 / inferLowRank.m:1-66: the V4 flow without the rank-one
    retry; V1 / V2 skip the refine when quality <= 0.6 (V2.m:47-58), V3 refines either way (V3.m:50-62)."""
    A = np.asarray(A, dtype=np.complex128)
    m, n = A.shape
    A, B, A_norm, B_norm = _preprocess(A, B, p.tol_abs)
    train_idx = np.asarray(train_idx, dtype=np.int64)
    # the spectral init of these versions runs INSIDE inferLowRankImpl, after r = min([r m n]) has been
    # re-evaluated with m = m_train (V3.m:198-200, V2.m:193-195, inferLowRank.m:193-195): r <= m_train
    r = min(p.r, m, n, int(train_idx.size))
    test_idx = test_index_set(m, train_idx)
    A_train, B_train = A[train_idx, :], B[train_idx]
    Xs = spectral_initialize(A_train, B_train, r)            # V3.m:213-224 (inside inferLowRankImpl there)
    X, Y, _ = infer_low_rank_impl(A_train, B_train, Xs, tx, rx, p.lam, r, p.mu0, p.rho, p.tol_rel, p.tol_abs,
                                  p.maxiter, profile, argmin_z, info.traces)
    quality = quality_score(A[test_idx, :], B[test_idx], X)
    info.quality, info.used_rank_one = quality, False
    info.trial_quality, info.trial_rank_one = [quality], [False]
    if quality > 0.6 or not refine_only_if_good:
        X, Y = _refine(A, B, X, Y, quality, profile, tx, rx, p, argmin_z, info)
    s = B_norm / A_norm
    return X.reshape(-1) * s, Y.reshape(-1) * s, quality


def infer_low_rank_v3(A, B, tx, rx, params: Params | None = None, *, train_idx, info: SolveInfo | None = None):
    """inferLowRankV3.m (main ADMM_v2.m version 3)."""
    return _older_version(A, B, tx, rx, params or Params(), train_idx, info or SolveInfo(), 0, False)


def infer_low_rank_v2(A, B, tx, rx, params: Params | None = None, *, train_idx, info: SolveInfo | None = None):
    """inferLowRankV2.m (main ADMM_v2.m version 2; mu0 = 1e-3, rho = 1.03, 95 % split are hard-coded there)."""
    return _older_version(A, B, tx, rx, params or Params(), train_idx, info or SolveInfo(), 3, True)


def infer_low_rank_v1(A, B, tx, rx, params: Params | None = None, *, train_idx, info: SolveInfo | None = None):
    """inferLowRank.m (main ADMM_v2.m version 1)."""
    return _older_version(A, B, tx, rx, params or Params(), train_idx, info or SolveInfo(), 2, True)


def infer_low_rank_v4(A, B, tx, rx, params: Params | None = None, *, train_idx, info: SolveInfo | None = None):
    """inferLowRankV4.m:1-88.  ``train_idx``: the randsample draw of :37 (0-based, drawn order)."""
    return _v4_like(A, B, tx, rx, params or Params(), train_idx, argmin_z, info or SolveInfo())


def infer_low_rank_nuclear(A, B, tx, rx, params: Params | None = None, *, train_idx,
                           info: SolveInfo | None = None):
    """inferLowRank_Nuclear.m:5-97 (init_mode is always 1: the 13th parameter does not exist)."""
    return _v4_like(A, B, tx, rx, params or Params(), train_idx, argmin_z_nuclear, info or SolveInfo())


def infer_low_rank_v4_multi(A, B, tx, rx, params: Params | None = None, *, train_idx,
                            info: SolveInfo | None = None):
    """inferLowRankV4_multi.m:5-109.  ``train_idx``: sequence of the 3 randsample draws of :48.

    Quirk reproduced: the refine branch tests the LAST trial's quality and uses the
    LAST trial's use_rank_one (:89,:92,:100), while X_max/Y_max come from the best trial.
    """
    p = params or Params()
    info = info or SolveInfo()
    A = np.asarray(A, dtype=np.complex128)
    m, n = A.shape
    r = min(p.r, m, n)
    A, B, A_norm, B_norm = _preprocess(A, B, p.tol_abs)
    assert len(train_idx) == 3
    max_quality = -1.0
    X_max = Y_max = None
    quality, use_rank_one = float("nan"), False
    for i in range(3):
        X, Y, quality, use_rank_one = _train_solve(A, B, train_idx[i], tx, rx, p, r, argmin_z, info)
        info.trial_quality.append(quality)
        info.trial_rank_one.append(use_rank_one)
        if max_quality < quality:
            X_max, Y_max, max_quality = X, Y, quality
            info.best_trial = i
    if X_max is None:  # all three qualities NaN: MATLAB would fail on undefined X_max
        X_max = np.full((n, 1), np.nan + 0j)
        Y_max = np.full((m, 1), np.nan + 0j)
    info.quality, info.used_rank_one = quality, use_rank_one
    X, Y = _refine(A, B, X_max, Y_max, quality, use_rank_one, tx, rx, p, argmin_z, info)
    s = B_norm / A_norm
    return X.reshape(-1) * s, Y.reshape(-1) * s, quality


# ------------------------------------------------------------------------------------------------
# inferMinL2.m (ADMM_v2.m:23, version 0): the same train / quality / refine shell around a plain
# least-squares phase retrieval -- no low-rank variable Z: X = pinv(A) (Y - M/mu)
def infer_admm_minl2(A, B, X0, scale_by_row, lam, tol_rel, tol_abs, maxiter, trace: StageTrace | None = None,
                     snapshot_iters=None):
    """inferMinL2.m:227-346 (InferADMM of that file) with lambda = 0: U = pinv(A) (:235-237)."""
    if lam != 0:
        raise NotImplementedError("lambda != 0 (eigen-path, inferMinL2.m:238-241) is dead in the reference")
    A = np.asarray(A, dtype=np.complex128)
    B = np.asarray(B, dtype=np.float64).reshape(-1)
    X0 = np.asarray(X0, dtype=np.complex128)
    if X0.ndim == 1:
        X0 = X0[:, None]
    m, n = A.shape
    r = X0.shape[1]
    U = np.linalg.pinv(A)                                       # :236
    X = X0.copy()
    AX = A @ X                                                  # :250
    normB = np.linalg.norm(B)
    with np.errstate(divide="ignore", invalid="ignore"):
        if scale_by_row:
            X = X * (normB / _fro(AX))                          # :252
        else:
            X = X * (normB / np.sqrt(np.sum(np.abs(AX) ** 2, axis=0)))[None, :]     # :254-256
    AX = A @ X
    Y = normalize_rows(AX, B, scale_by_row)                     # :262
    AtY = A.conj().T @ Y
    M = np.zeros((m, r), dtype=np.complex128)
    mu, rho = 0.001, 1.03                                       # :269-270
    opt_obj, last_res = np.inf, np.inf
    opt_X = opt_Y = None
    converged = False
    it = 0
    opt_iter, opt_col, bumps = -1, -1, 0
    for it in range(1, maxiter + 1):
        Y0, AtY0 = Y, AtY
        X = U @ (Y - M / mu)                                    # :282, :350
        AX = A @ X
        Y = argmin_y(AX, B, M, mu, scale_by_row)                # :286
        AtY = A.conj().T @ Y
        J_M = AX - Y
        M = M + mu * J_M                                        # :291-292
        if scale_by_row:                                        # :296-312
            obj = np.linalg.norm(np.sqrt(np.sum(np.abs(AX) ** 2, axis=1)) - B)
            if obj < opt_obj:
                opt_obj, opt_X, opt_Y, opt_iter, opt_col = obj, X.copy(), Y.copy(), it, -1
        else:
            objs = np.sqrt(np.sum((np.abs(AX) - B[:, None]) ** 2, axis=0))
            obj, j = _min_skip_nan(objs)
            if obj < opt_obj:
                opt_obj, opt_X, opt_Y, opt_iter, opt_col = obj, X[:, j:j + 1].copy(), Y[:, j:j + 1].copy(), it, j
        res_prim = _fro(J_M)                                    # :316-325
        res_dual = mu * _fro(AtY - AtY0)
        res_comb = math.sqrt(res_prim ** 2 + _fro(Y - Y0) ** 2)
        nAX, nY = _fro(AX), _fro(Y)
        th_prim = tol_abs * math.sqrt(m * r) + tol_rel * max(nAX, nY)
        th_dual = tol_abs * math.sqrt(n * r) + tol_rel * _fro(AtY)
        th_comb = tol_abs * math.sqrt(m * r * 2) + tol_rel * math.sqrt(max(nAX, nY) ** 2 + nY ** 2)
        if snapshot_iters is not None and it in snapshot_iters:
            snapshot_iters[it] = dict(X=X.copy(), Y=Y.copy(), M=M.copy())
        if (res_prim < th_prim and res_dual < th_dual) or (res_comb < th_comb):
            converged = True
            break
        if res_comb > last_res * 0.9:                           # :332-334
            mu *= rho
            bumps += 1
        last_res = res_comb
    if trace is not None:
        trace.iters, trace.converged, trace.mu = it, converged, mu
        trace.opt_iter, trace.opt_col, trace.n_mu_bumps = opt_iter, opt_col, bumps
    if opt_X is None:                                           # every objective NaN: MATLAB errors on opt_X undefined
        opt_X = np.full((n, r if scale_by_row else 1), np.nan, dtype=np.complex128)
        opt_Y = np.full((m, r if scale_by_row else 1), np.nan, dtype=np.complex128)
    return opt_X, opt_Y, converged


def spectral_initialize_minl2(A, B, r):
    """inferMinL2.m:163-196: SpectralInitialize plus the 90 %-energy rank rule (:181-185, applied twice: idempotent)."""
    A = np.asarray(A, dtype=np.complex128)
    B = np.asarray(B, dtype=np.float64).reshape(-1)
    m, n = A.shape
    As = A.copy()
    an = np.sqrt(np.sum(np.abs(A) ** 2, axis=1))
    nz = an != 0
    As[nz, :] = A[nz, :] * (B[nz] / an[nz])[:, None]
    s2, idx, V = _eigh_desc(As.conj().T @ As)
    for _ in range(2):
        if np.sum(s2[:r]) >= np.sum(s2) * 0.9:
            k = int(np.nonzero(np.cumsum(s2) >= np.sum(s2) * 0.9)[0][0]) + 1
            r = min(max(k, 3), m, n)
    return V[:, idx[:r]] * np.sqrt(s2[:r])[None, :]


def infer_min_l2(A, B, lam=0.0, r=20, tol_rel=1e-4, tol_abs=1e-8, maxiter=500, *, train_idx, info: SolveInfo | None = None):
    """[X, Y, quality] = inferMinL2(A, B, lambda, r, tol_rel, tol_abs, maxiter)  (inferMinL2.m:1-66).
    ``train_idx``: the randsample(m, ceil(m*0.95)) draw of :34 (0-based)."""
    info = info if info is not None else SolveInfo()
    A = np.asarray(A, dtype=np.complex128)
    m, n = A.shape
    r = min(r, m, n)                                            # :9
    A, B, A_norm, B_norm = _preprocess(A, B, tol_abs)           # :17-28
    train_idx = np.asarray(train_idx, dtype=np.int64)
    assert train_idx.size == math.ceil(m * 0.95)
    test_idx = test_index_set(m, train_idx)
    A_train, B_train = A[train_idx, :], B[train_idx]
    # inferMinL2Impl (:68-225)
    X = spectral_initialize_minl2(A_train, B_train, r)
    ta, tb = StageTrace(), StageTrace()
    X, Y, _ = infer_admm_minl2(A_train, B_train, X, True, lam, tol_rel, tol_abs, maxiter, ta)
    G = X.conj().T @ X                                          # :219-220  [Vx,Dx] = eig(X'*X); X = X*Vx
    _, Vx = np.linalg.eigh(0.5 * (G + G.conj().T))
    X = X @ Vx
    X, Y, _ = infer_admm_minl2(A_train, B_train, X, False, lam, tol_rel, tol_abs, maxiter, tb)
    info.traces += [ta, tb]
    quality = quality_score(A[test_idx, :], B[test_idx], X) if test_idx.size else float("nan")   # :42 (0/0 -> NaN)
    info.quality = quality
    if quality > 0.6:                                           # :47-58
        X0, Y0 = X, Y
        tr = StageTrace()
        X, Y, _ = infer_admm_minl2(A, B, X0, True, lam, tol_rel, tol_abs, maxiter, tr)
        info.traces.append(tr)
        with np.errstate(divide="ignore", invalid="ignore"):
            similarity = float(np.abs(np.vdot(X0, X)) / np.linalg.norm(X0) / np.linalg.norm(X))
        info.similarity = similarity
        if similarity < 0.6:
            X, Y = X0, Y0
            info.rolled_back = True
    s = B_norm / A_norm
    return X.reshape(-1) * s, Y.reshape(-1) * s, quality


def admm_v2(measurements, FW, TX, RX, version, *, train_idx, tree="main", params: Params | None = None,
            info: SolveInfo | None = None):
    """Version switch.  ``tree``: 'main' (ADMM_v2.m:22-45), 'main_nuclear' (ADMM_v2_nuclear.m:32)
    or 'ns' (Numerical_Simulation/.../ADMM_v2.m:22-41).  Only the V4-family rows of SURVEY §2.2
    are on the hot path; the others raise NotImplementedError (scope row §8f-4)."""
    B = np.asarray(measurements, dtype=np.float64).reshape(-1)
    if version == 0:
        return infer_min_l2(FW, B, train_idx=train_idx, info=info)
    if tree == "main" and version == 4:
        return infer_low_rank_v4_multi(FW, B, TX, RX, params, train_idx=train_idx, info=info)
    if tree == "main" and version in (1, 2, 3):
        fn = {1: infer_low_rank_v1, 2: infer_low_rank_v2, 3: infer_low_rank_v3}[version]
        return fn(FW, B, TX, RX, params, train_idx=train_idx, info=info)
    if tree == "main_nuclear" and version == 4:
        return infer_low_rank_nuclear(FW, B, TX, RX, params, train_idx=train_idx, info=info)
    if tree == "ns" and version == 3:
        return infer_low_rank_v4(FW, B, TX, RX, params, train_idx=train_idx, info=info)
    raise NotImplementedError(f"version {version} of tree {tree!r} is outside the hot path")
