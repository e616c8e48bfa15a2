"""Host-side mirror of the reference's solver interface for the ADMM hot path.

Same names, argument meaning and outputs as the MATLAB functions they replace:

* ``inferLowRankV4``        — main/src/my_recovery_algorithms/ADMM_v2/inferLowRankV4.m:1-9
* ``inferLowRankV4_multi``  — …/inferLowRankV4_multi.m:5-13
* ``inferLowRank_Nuclear``  — …/inferLowRank_Nuclear.m:5-13
* ``ADMM_v2`` / ``ADMM_v2_nuclear`` — main/src/my_recovery_algorithms/ADMM_v2.m:1,22-45, ADMM_v2_nuclear.m:32,
  Numerical_Simulation/src/my_recovery_algorithms/ADMM_v2/ADMM_v2.m:22-41 (``tree='ns'``)

plus the batched forms the single-instance calls are built on.  All arithmetic runs in the CUDA
library (``lib.py`` -> ``libtwoace.so``); this file only marshals arrays.  The reference draws its
train/test split from MATLAB's global RNG (``randsample``, inferLowRankV4.m:37), which cannot be
reproduced outside MATLAB: pass ``train_idx`` (0-based, drawn order) to pin the split, otherwise it
is drawn from ``rng`` (NumPy ``Generator``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from . import lib as _lib
from .lib import MINL2, NUCLEAR, V1, V2, V3, V4, V4_MULTI, Params, PlOpts

__all__ = ["inferLowRankV4", "inferLowRankV4_multi", "inferLowRank_Nuclear", "ADMM_v2", "ADMM_v2_nuclear",
           "solve_batch", "solve_batch_codebook", "infer_admm_batch", "spectral_init_batch", "BatchResult",
           "draw_train_idx"]


def _params(lambda_, r, mu0, rho, cc_frac, tol_rel, tol_abs, maxiter) -> Params:
    return Params(float(lambda_), int(r), float(mu0), float(rho), float(cc_frac), float(tol_rel),
                  float(tol_abs), int(maxiter))


def draw_train_idx(m: int, cc_frac: float, ntrial: int, rng: np.random.Generator) -> np.ndarray:
    """The randsample(m, floor(m*cc_frac)) draws of inferLowRankV4.m:37 from a NumPy stream."""
    k = int(math.floor(m * cc_frac))
    return np.stack([rng.permutation(m)[:k] for _ in range(ntrial)]).astype(np.int32)


@dataclass
class BatchResult:
    X: np.ndarray            # [nb, n] complex128
    Y: list                  # nb arrays [y_rows_b] complex128
    quality: np.ndarray      # [nb]
    info: np.ndarray         # [nb, 16] (twoace.h TWOACE_INFO_WORDS)
    stage_words: np.ndarray  # [nb, 4T+1, 12]


def train_rows(variant: int, m: int, cc_frac: float) -> int:
    """Entries of one train draw: floor(m*cc_frac) (inferLowRankV4.m:36); ceil(0.95 m) for inferMinL2 (inferMinL2.m:34)."""
    return int(math.ceil(m * 0.95)) if variant == MINL2 else int(math.floor(m * cc_frac))


def _concat_inputs(A_list, B_list, train_idx_list, ntrial, cc_frac, n, variant=V4):
    nb = len(A_list)
    m = np.array([np.shape(a)[0] for a in A_list], dtype=np.int32)
    for a in A_list:
        if np.shape(a)[1] != n:
            raise ValueError(f"A must have tx*rx = {n} columns, got {np.shape(a)[1]}")
    A = np.concatenate([np.asarray(a, dtype=np.complex128).reshape(-1, order="F") for a in A_list]) if nb else \
        np.zeros(0, np.complex128)
    B = np.concatenate([np.asarray(b, dtype=np.float64).reshape(-1) for b in B_list]) if nb else np.zeros(0)
    tr = []
    for b in range(nb):
        t = np.asarray(train_idx_list[b], dtype=np.int32).reshape(ntrial, -1)
        k = train_rows(variant, int(m[b]), cc_frac)
        if t.shape[1] != k:
            raise ValueError(f"instance {b}: train_idx needs {k} entries per draw, got {t.shape[1]}")
        tr.append(t.reshape(-1))
    T = np.concatenate(tr) if nb else np.zeros(0, np.int32)
    return m, np.ascontiguousarray(A), np.ascontiguousarray(B), np.ascontiguousarray(T.astype(np.int32))


def _unpack(nb, n, m, X, Y, q, info, sw, ntrial):
    offs = np.concatenate([[0], np.cumsum(m)])
    info = info.reshape(nb, _lib.INFO_WORDS)
    Ys = [Y[offs[b]:offs[b] + int(info[b, 5])].copy() for b in range(nb)]
    return BatchResult(X.reshape(nb, n), Ys, q, info, sw.reshape(nb, 4 * ntrial + 1, _lib.STAGE_WORDS))


def solve_batch(variant: int, A_list, B_list, tx: int, rx: int, train_idx_list, params: Params | None = None,
                ctx: _lib.Context | None = None) -> BatchResult:
    """Batched inferLowRankV4 / _multi / _Nuclear on dense per-instance A (ragged m)."""
    ctx = ctx or _lib.default_context()
    p = params or Params.default()
    nb, n = len(A_list), tx * rx
    ntrial = 3 if variant == V4_MULTI else 1
    m, A, B, T = _concat_inputs(A_list, B_list, train_idx_list, ntrial, p.cc_frac, n, variant)
    X = np.empty(nb * n, np.complex128)
    Y = np.empty(int(m.sum()), np.complex128)
    q = np.empty(nb, np.float64)
    info = np.empty(nb * _lib.INFO_WORDS, np.float64)
    sw = np.empty(nb * (4 * ntrial + 1) * _lib.STAGE_WORDS, np.float64)
    ctx.solve_batch_raw(variant, _lib.MEM_HOST, nb, tx, rx, m, A, B, T, p, X, Y, q, info, sw)
    return _unpack(nb, n, m, X, Y, q, info, sw, ntrial)


def solve_batch_codebook(variant: int, rows_list, row_scale: float, B_list, tx: int, rx: int, train_idx_list,
                         params: Params | None = None, ctx: _lib.Context | None = None) -> BatchResult:
    """Same with A_b = row_scale * cb[rows_b] for the codebook registered by ``ctx.set_codebook``
    (row selection of channel_recovery_ADMM_v2_simulation_A2only.m:137-138)."""
    ctx = ctx or _lib.default_context()
    p = params or Params.default()
    nb, n = len(rows_list), tx * rx
    ntrial = 3 if variant == V4_MULTI else 1
    m = np.array([len(r) for r in rows_list], dtype=np.int32)
    rows = np.ascontiguousarray(np.concatenate([np.asarray(r, dtype=np.int32) for r in rows_list]))
    B = np.ascontiguousarray(np.concatenate([np.asarray(b, dtype=np.float64).reshape(-1) for b in B_list]))
    tr = []
    for b in range(nb):
        t = np.asarray(train_idx_list[b], dtype=np.int32).reshape(ntrial, -1)
        k = train_rows(variant, int(m[b]), p.cc_frac)
        if t.shape[1] != k:
            raise ValueError(f"instance {b}: train_idx needs {k} entries per draw, got {t.shape[1]}")
        tr.append(t.reshape(-1))
    T = np.ascontiguousarray(np.concatenate(tr).astype(np.int32))
    X = np.empty(nb * n, np.complex128)
    Y = np.empty(int(m.sum()), np.complex128)
    q = np.empty(nb, np.float64)
    info = np.empty(nb * _lib.INFO_WORDS, np.float64)
    sw = np.empty(nb * (4 * ntrial + 1) * _lib.STAGE_WORDS, np.float64)
    ctx.solve_batch_codebook_raw(variant, _lib.MEM_HOST, nb, tx, rx, m, rows, row_scale, B, T, p, X, Y, q, info, sw)
    return _unpack(nb, n, m, X, Y, q, info, sw, ntrial)


def infer_admm_batch(A_list, B_list, X0_list, scale_by_row: bool, use_rank_one: bool, tx: int, rx: int,
                     params: Params | None = None, nuclear: bool = False, ctx: _lib.Context | None = None):
    """One InferADMM call per instance (inferLowRankV4.m:260-365).  Returns (X, Y, state, words) lists;
    state[b] = dict(X, Z, N, Y, M) of the final iterate, words[b] = twoace.h stage words."""
    ctx = ctx or _lib.default_context()
    p = params or Params.default()
    nb, n = len(A_list), tx * rx
    r = int(np.shape(X0_list[0])[1]) if np.ndim(X0_list[0]) == 2 else 1
    rout = r if scale_by_row else 1
    m = np.array([np.shape(a)[0] for a in A_list], dtype=np.int32)
    A = np.ascontiguousarray(np.concatenate([np.asarray(a, np.complex128).reshape(-1, order="F") for a in A_list]))
    B = np.ascontiguousarray(np.concatenate([np.asarray(b, np.float64).reshape(-1) for b in B_list]))
    X0 = np.ascontiguousarray(np.concatenate(
        [np.asarray(x, np.complex128).reshape(n, r).reshape(-1, order="F") for x in X0_list]))
    X = np.empty(nb * n * rout, np.complex128)
    Y = np.empty(int(m.sum()) * rout, np.complex128)
    st_sizes = [3 * n * r + 2 * int(mm) * r for mm in m]
    state = np.empty(int(sum(st_sizes)), np.complex128)
    words = np.empty(nb * _lib.STAGE_WORDS, np.float64)
    ctx.infer_admm_batch_raw(_lib.MEM_HOST, nb, tx, rx, m, A, B, r, X0, scale_by_row, use_rank_one, nuclear, p,
                             X, Y, state, words)
    Xs, Ys, Ss = [], [], []
    yo = so = 0
    for b in range(nb):
        mb = int(m[b])
        Xs.append(X[b * n * rout:(b + 1) * n * rout].reshape(n, rout, order="F").copy())
        Ys.append(Y[yo:yo + mb * rout].reshape(mb, rout, order="F").copy())
        yo += mb * rout
        s = state[so:so + st_sizes[b]]
        so += st_sizes[b]
        nr, mr = n * r, mb * r
        Ss.append(dict(X=s[:nr].reshape(n, r, order="F"), Z=s[nr:2 * nr].reshape(n, r, order="F"),
                       N=s[2 * nr:3 * nr].reshape(n, r, order="F"),
                       Y=s[3 * nr:3 * nr + mr].reshape(mb, r, order="F"),
                       M=s[3 * nr + mr:].reshape(mb, r, order="F")))
    return Xs, Ys, Ss, words.reshape(nb, _lib.STAGE_WORDS)


def spectral_init_batch(A_list, B_list, r: int, ctx: _lib.Context | None = None):
    """SpectralInitialize (inferLowRankV4.m:540-553) per instance -> list of n x r arrays."""
    ctx = ctx or _lib.default_context()
    nb = len(A_list)
    n = int(np.shape(A_list[0])[1])
    m = np.array([np.shape(a)[0] for a in A_list], dtype=np.int32)
    A = np.ascontiguousarray(np.concatenate([np.asarray(a, np.complex128).reshape(-1, order="F") for a in A_list]))
    B = np.ascontiguousarray(np.concatenate([np.asarray(b, np.float64).reshape(-1) for b in B_list]))
    Xs = np.empty(nb * n * r, np.complex128)
    ctx.spectral_init_batch_raw(_lib.MEM_HOST, nb, n, m, A, B, r, Xs)
    return [Xs[b * n * r:(b + 1) * n * r].reshape(n, r, order="F").copy() for b in range(nb)]


def phaselift_batch(A_list, y_list, opts: PlOpts | None = None, ctx: _lib.Context | None = None):
    """Batch of MyPhaseLift solves (MyPhaseLift.m:69-107) on dense sensing matrices, ragged in m.
    ``y_list`` holds intensities.  Returns (list of n-vectors, info array [nb, 8])."""
    ctx = ctx or _lib.default_context()
    nb = len(A_list)
    if nb == 0:
        return [], np.zeros((0, _lib.PL_INFO_WORDS))
    n = int(np.shape(A_list[0])[1])
    m = np.array([np.shape(a)[0] for a in A_list], dtype=np.int32)
    A = np.ascontiguousarray(np.concatenate([np.asarray(a, np.complex128).reshape(-1, order="F") for a in A_list]))
    y = np.ascontiguousarray(np.concatenate([np.asarray(b, np.float64).reshape(-1) for b in y_list]))
    if y.size != int(m.sum()):
        raise ValueError("measurement vectors do not match the row counts of the sensing matrices")
    sig = np.empty(nb * n, np.complex128)
    info = np.empty((nb, _lib.PL_INFO_WORDS), np.float64)
    ctx.phaselift_batch_raw(_lib.MEM_HOST, nb, n, m, A, None, 1.0, y, opts or PlOpts.default(), sig, info)
    return [sig[b * n:(b + 1) * n].copy() for b in range(nb)], info


def phaselift_batch_codebook(rows_list, row_scale: float, y_list, n: int, opts: PlOpts | None = None,
                             ctx: _lib.Context | None = None):
    """As phaselift_batch with sensing rows taken from the registered codebook (Context.set_codebook)."""
    ctx = ctx or _lib.default_context()
    nb = len(rows_list)
    if nb == 0:
        return [], np.zeros((0, _lib.PL_INFO_WORDS))
    m = np.array([len(r) for r in rows_list], dtype=np.int32)
    rows = np.ascontiguousarray(np.concatenate([np.asarray(r, np.int32).reshape(-1) for r in rows_list]))
    y = np.ascontiguousarray(np.concatenate([np.asarray(b, np.float64).reshape(-1) for b in y_list]))
    if y.size != int(m.sum()):
        raise ValueError("measurement vectors do not match the row lists")
    sig = np.empty(nb * n, np.complex128)
    info = np.empty((nb, _lib.PL_INFO_WORDS), np.float64)
    ctx.phaselift_batch_raw(_lib.MEM_HOST, nb, n, m, None, rows, row_scale, y, opts or PlOpts.default(), sig, info)
    return [sig[b * n:(b + 1) * n].copy() for b in range(nb)], info


def evaluation_batch(X_est, X_true, tx: int, rx: int, phase_bit: int = 2, ctx: _lib.Context | None = None):
    """Evaluation_H.m:81-115 per instance -> array [nb, 4] = (MSE_H, gain_ana, gain_dig, proj_error).
    X_est / X_true: [nb, tx*rx] complex (vec of the rx x tx channel, column-major)."""
    ctx = ctx or _lib.default_context()
    Xe = np.ascontiguousarray(np.asarray(X_est, np.complex128))
    Xt = np.ascontiguousarray(np.asarray(X_true, np.complex128))
    if Xe.shape != Xt.shape or Xe.ndim != 2 or Xe.shape[1] != tx * rx:
        raise ValueError("X_est and X_true must both be [nb, tx*rx]")
    out = np.empty((Xe.shape[0], _lib.METRIC_WORDS), np.float64)
    ctx.metrics_batch_raw(_lib.MEM_HOST, Xe.shape[0], int(tx), int(rx), Xe, Xt, phase_bit, out)
    return out


def angle_evaluation_batch(X_est, angles_true, tx: int, rx: int, nqt: int | None = None, nqr: int | None = None,
                           searching_area: float = 95.0, wavelength: float = 3e8 / 60.48e9, spacing: float = 3.055e-3,
                           ctx: _lib.Context | None = None):
    """Evaluation_Recovery.m:85-146 per instance -> [nb, 6] = (AoD_Err_to_True, AoA_Err_to_True, AoDA_Err and the same
    three against the quantised true angles), degrees.  angles_true: [nb, 2L] (AoD then AoA)."""
    ctx = ctx or _lib.default_context()
    Xe = np.ascontiguousarray(np.asarray(X_est, np.complex128))
    ang = np.ascontiguousarray(np.asarray(angles_true, np.float64))
    nb, L = Xe.shape[0], ang.shape[1] // 2
    out = np.empty((nb, _lib.ANGLE_WORDS), np.float64)
    ctx.angle_metrics_batch_raw(_lib.MEM_HOST, nb, int(tx), int(rx), L, nqt or 4 * tx, nqr or 4 * rx, searching_area,
                                wavelength, spacing, Xe, ang, out)
    return out


# ----------------------------------------------------------------------------- MATLAB-signature calls
def synth_batch(m, snr_db, row_lo, row_hi, trial_id, sp: "_lib.SynthParams | None" = None,
                ctx: _lib.Context | None = None):
    """Instances of the numerical-simulation workload built on the GPU from the registered codebook
    (twoace_synth_batch: Generate_Channel.m:76-139, A2only.m:137, Generate_Measurement.m:84-101, inferLowRankV4.m:36-37).
    Returns per-instance lists: rows, train_idx [ntrain, k], B, vecH and the angles [nb, 2L]."""
    ctx = ctx or _lib.default_context()
    sp = sp or _lib.SynthParams.default()
    m = np.ascontiguousarray(np.asarray(m, dtype=np.int32))
    nb, n = len(m), sp.nt * sp.nr
    bc = lambda v, dt: np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=dt), (nb,)))
    snr, lo, hi, tid = bc(snr_db, np.float64), bc(row_lo, np.int32), bc(row_hi, np.int32), bc(trial_id, np.int64)
    mtr = np.floor(m * sp.cc_frac).astype(np.int64)
    rows = np.empty(int(m.sum()), np.int32)
    train = np.empty(int((mtr * sp.ntrain).sum()), np.int32)
    B = np.empty(int(m.sum()), np.float64)
    vecH = np.empty(nb * n, np.complex128)
    ang = np.empty(nb * 2 * sp.L, np.float64)
    ctx.synth_batch_raw(_lib.MEM_HOST, nb, sp, m, snr, lo, hi, tid, rows, train, B, vecH, ang)
    bo = np.concatenate([[0], np.cumsum(m)])
    to = np.concatenate([[0], np.cumsum(mtr * sp.ntrain)])
    return dict(rows=[rows[bo[b]:bo[b + 1]] for b in range(nb)],
                train_idx=[train[to[b]:to[b + 1]].reshape(sp.ntrain, -1) for b in range(nb)],
                B=[B[bo[b]:bo[b + 1]] for b in range(nb)], vecH=vecH.reshape(nb, n), angles=ang.reshape(nb, 2 * sp.L))


def MyPhaseLift(measurements, measurementMat, *, opts: PlOpts | None = None, ctx=None):
    """recoveredSig = MyPhaseLift(measurements, measurementMat)   (MyPhaseLift.m:69)."""
    sig, _ = phaselift_batch([np.asarray(measurementMat, np.complex128)], [np.asarray(measurements).reshape(-1)],
                             opts, ctx)
    return sig[0]


def _single(variant, A, B, tx, rx, p: Params, train_idx, rng, ctx):
    A = np.asarray(A, dtype=np.complex128)
    m = A.shape[0]
    ntrial = 3 if variant == V4_MULTI else 1
    if train_idx is None:
        train_idx = draw_train_idx(m, p.cc_frac, ntrial, rng or np.random.default_rng())
    res = solve_batch(variant, [A], [B], int(tx), int(rx), [train_idx], p, ctx)
    return res.X[0], res.Y[0], float(res.quality[0])


def inferLowRankV4(A, B, tx, rx, lambda_=0.0, r=20, mu0=1e-3, rho=1.03, cc_frac=0.95, tol_rel=1e-4,
                   tol_abs=1e-8, maxiter=500, *, train_idx=None, rng=None, ctx=None):
    """[X, Y, quality] = inferLowRankV4(A, B, tx, rx, lambda, r, mu0, rho, cc_frac, tol_rel, tol_abs, maxiter)."""
    return _single(V4, A, B, tx, rx, _params(lambda_, r, mu0, rho, cc_frac, tol_rel, tol_abs, maxiter),
                   train_idx, rng, ctx)


def inferLowRankV4_multi(A, B, tx, rx, lambda_=0.0, r=20, mu0=1e-3, rho=1.03, cc_frac=0.95, tol_rel=1e-4,
                         tol_abs=1e-8, maxiter=500, *, train_idx=None, rng=None, ctx=None):
    """[X, Y, quality] = inferLowRankV4_multi(...): three random restarts, best held-out quality kept."""
    return _single(V4_MULTI, A, B, tx, rx, _params(lambda_, r, mu0, rho, cc_frac, tol_rel, tol_abs, maxiter),
                   train_idx, rng, ctx)


def inferLowRank_Nuclear(A, B, tx, rx, lambda_=0.0, r=20, mu0=1e-3, rho=1.03, cc_frac=0.95, tol_rel=1e-4,
                         tol_abs=1e-8, maxiter=500, *, train_idx=None, rng=None, ctx=None):
    """[X, Y, quality] = inferLowRank_Nuclear(...): singular-value soft-threshold ArgMinZ."""
    return _single(NUCLEAR, A, B, tx, rx, _params(lambda_, r, mu0, rho, cc_frac, tol_rel, tol_abs, maxiter),
                   train_idx, rng, ctx)


def inferLowRankV3(A, B, tx, rx, lambda_=0.0, r=20, mu0=1e-3, rho=1.03, cc_frac=0.95, tol_rel=1e-4,
                   tol_abs=1e-8, maxiter=500, *, train_idx=None, rng=None, ctx=None):
    """[X, Y, quality] = inferLowRankV3(...): V4 without the rank-one rerun (inferLowRankV3.m:1)."""
    return _single(V3, A, B, tx, rx, _params(lambda_, r, mu0, rho, cc_frac, tol_rel, tol_abs, maxiter),
                   train_idx, rng, ctx)


def inferLowRankV2(A, B, tx, rx, lambda_=0.0, r=20, tol_rel=1e-4, tol_abs=1e-8, maxiter=500, *, train_idx=None,
                   rng=None, ctx=None):
    """[X, Y, quality] = inferLowRankV2(A, B, tx, rx, lambda, r, tol_rel, tol_abs, maxiter)  (inferLowRankV2.m:1;
    mu0 = 1e-3, rho = 1.03 and the 95 % split are hard-coded there)."""
    return _single(V2, A, B, tx, rx, _params(lambda_, r, 1e-3, 1.03, 0.95, tol_rel, tol_abs, maxiter),
                   train_idx, rng, ctx)


def inferLowRank(A, B, tx, rx, lambda_=0.0, r=20, tol_rel=1e-4, tol_abs=1e-8, maxiter=500, *, train_idx=None,
                 rng=None, ctx=None):
    """[X, Y, quality] = inferLowRank(A, B, tx, rx, lambda, r, tol_rel, tol_abs, maxiter)  (inferLowRank.m:1)."""
    return _single(V1, A, B, tx, rx, _params(lambda_, r, 1e-3, 1.03, 0.95, tol_rel, tol_abs, maxiter),
                   train_idx, rng, ctx)


def inferMinL2(A, B, lambda_=0.0, r=20, tol_rel=1e-4, tol_abs=1e-8, maxiter=500, *, train_idx=None, rng=None, ctx=None):
    """[X, Y, quality] = inferMinL2(A, B, lambda, r, tol_rel, tol_abs, maxiter)  (ADMM_v2/inferMinL2.m:1; ADMM_v2.m:23).
    ``train_idx``: the randsample(m, ceil(m*0.95)) draw of :34, 0-based (drawn from ``rng`` when absent)."""
    A = np.asarray(A, dtype=np.complex128)
    m, n = A.shape
    tx = next((t for t in (16, 32, 8, 4) if n % t == 0), None)      # only n = tx * rx matters to this solver
    if tx is None:
        raise ValueError("inferMinL2 through this library needs n divisible by 4")
    p = _params(lambda_, r, 1e-3, 1.03, 0.95, tol_rel, tol_abs, maxiter)
    if train_idx is None:
        train_idx = (rng or np.random.default_rng()).permutation(m)[:train_rows(MINL2, m, 0.95)].astype(np.int32)
    res = solve_batch(MINL2, [A], [B], tx, n // tx, [train_idx], p, ctx)
    return res.X[0], res.Y[0], float(res.quality[0])


def ADMM_v2(measurements, FW, TX, RX, version, *, tree="main", train_idx=None, rng=None, ctx=None):
    """[X, Y, converged] = ADMM_v2(measurements, FW, TX, RX, version).  As in the reference the third
    output is really ``quality`` (ADMM_v2.m:31-32 vs inferLowRankV4.m:1).  Version map: main/.../ADMM_v2.m:22-31
    (1 inferLowRank, 2 inferLowRankV2, 3 inferLowRankV3, 4 inferLowRankV4_multi) and
    Numerical_Simulation/.../ADMM_v2.m:22-29 (1 inferLowRank, 2 inferLowRankV3, 3 inferLowRankV4).  Version 0
    (inferMinL2) is served by ``inferMinL2``; the trailing else-loop (main :33-43, NS :30-41) calls inferLowRankV2 with
    lambda = rz + 2 != 0, the eigen-path of ArgMinX that no caller of the reference reaches -- not built."""
    B = np.asarray(measurements, dtype=np.float64).reshape(-1)
    kw = dict(train_idx=train_idx, rng=rng, ctx=ctx)
    if version == 0:
        return inferMinL2(FW, B, train_idx=train_idx, rng=rng, ctx=ctx)
    if tree == "main":
        table = {1: inferLowRank, 2: inferLowRankV2, 3: inferLowRankV3, 4: inferLowRankV4_multi}
    elif tree == "ns":
        table = {1: inferLowRank, 2: inferLowRankV3, 3: inferLowRankV4}
    else:
        raise ValueError(f"unknown tree {tree!r} (main | ns)")
    if version in table:
        return table[version](FW, B, TX, RX, **kw)
    raise NotImplementedError(f"ADMM_v2 version {version} (tree {tree!r}): the lambda != 0 retry loop is dead code in the "
                              "reference and outside this build")


def ADMM_v2_nuclear(measurements, FW, TX, RX, version, *, train_idx=None, rng=None, ctx=None):
    """main/src/my_recovery_algorithms/ADMM_v2_nuclear.m:22-32: versions 1-3 as ADMM_v2, 4 -> inferLowRank_Nuclear."""
    B = np.asarray(measurements, dtype=np.float64).reshape(-1)
    kw = dict(train_idx=train_idx, rng=rng, ctx=ctx)
    if version == 0:
        return inferMinL2(FW, B, train_idx=train_idx, rng=rng, ctx=ctx)
    table = {1: inferLowRank, 2: inferLowRankV2, 3: inferLowRankV3, 4: inferLowRank_Nuclear}
    if version in table:
        return table[version](FW, B, TX, RX, **kw)
    raise NotImplementedError(f"ADMM_v2_nuclear version {version}: the lambda != 0 retry loop is outside this build")
