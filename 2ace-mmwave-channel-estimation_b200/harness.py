"""Synthetic instance generator and metrics for the ADMM CSI-recovery path (host side, NumPy).

Restates, for input synthesis only (SURVEY.md §8d "Config 1"):

* Eq. 23 sparse multipath channel — ``Numerical_Simulation/src/generate_channel/Generate_Channel.m:64-142``
  (L>1 => no Rician tail, ``:98-106``; ``vecH = vec(H_Matrix)`` with ``H_Matrix`` Nr x Nt, ``:139``)
* measurement model ``|FW*vecH + noise|`` with signal power 1 —
  ``Numerical_Simulation/src/generate_measurement/Generate_Measurement.m:84-101``
* 2-bit random beams — ``Numerical_Simulation/src/Generate_random_beam.m:31-34``
* NMSE — ``Numerical_Simulation/src/evaluate_plot_results/Evaluation_H.m:81-89``

MATLAB's randn/randperm/randsample streams cannot be reproduced outside MATLAB, so the
generator uses a documented NumPy stream: one ``SeedSequence`` child per trial, draw order
``[AoD, AoA, gains, row subset, noise, train_idx (x3)]``.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")

# first entry of the 40 hard-coded seeds of channel_recovery_ADMM_v2_simulation_A2only.m:103
BASE_SEED = 58659179

WAVELENGTH = 3e8 / 60.48e9   # A2only.m:38
ANT_SPACING = 3.055e-3       # A2only.m:39
SEARCHING_AREA = 95.0        # A2only.m:52

_ROOTS = np.array([1, 1j, -1, -1j], dtype=np.complex128)


def unpack_codes(packed: np.ndarray, shape) -> np.ndarray:
    """2-bit phase codes (4 per byte, little-endian in the byte) -> uint8 array of ``shape``."""
    p = np.asarray(packed, dtype=np.uint8)
    k = np.stack([(p >> s) & 3 for s in (0, 2, 4, 6)], axis=1).reshape(-1)
    return k.reshape(tuple(int(x) for x in shape))


def load_codebook_codes(name: str = "random_probe_cb_16x16") -> np.ndarray:
    """Phase codes k (uint8, exp(1j*k*pi/2)) of a shipped codebook (fixture built from
    ``codebook/codebook_mat/<name>.mat`` by ``tests/golden/make_codebook_fixture.py``)."""
    z = np.load(os.path.join(_DATA, name + ".u2.npz"))
    return unpack_codes(z["codes"], z["shape"])


def load_codebook(name: str = "random_probe_cb_16x16") -> np.ndarray:
    """Complex128 codebook ``cb`` with exact unit-modulus 4-phase entries."""
    return _ROOTS[load_codebook_codes(name)]


def load_codebook_mat(path: str, var: str = "cb") -> np.ndarray:
    """Read a reference codebook file (codebook/codebook_mat/*.mat, MATLAB v5, variable ``cb``: rows x Nt*Nr
    complex, or the 3-D directional layout which is flattened to rows x n) the way the entry points receive it
    (cb_amp .* exp(1j*cb_angle), ...simulation_A2only.m:120).  Entries within 1e-12 of a 4th root of unity are
    snapped to it, so that twoace_set_codebook finds exact 2-bit phase codes."""
    from scipy.io import loadmat
    cb = np.asarray(loadmat(path)[var], dtype=np.complex128)
    if cb.ndim == 3:
        cb = cb.reshape(-1, cb.shape[-1])
    k = np.round(np.angle(cb) / (np.pi / 2)).astype(np.int64) % 4
    snapped = _ROOTS[k] * np.abs(cb)
    close = np.abs(cb - snapped) <= 1e-12 * np.maximum(np.abs(cb), 1e-300)
    return np.where(close, snapped, cb)


def random_beam_codes(rng: np.random.Generator, rows: int, n: int, phase_bit: int = 2) -> np.ndarray:
    """Random Np-phase beam codes (Generate_random_beam.m:31-34, Np = Phase_Bit^2)."""
    return rng.integers(0, phase_bit ** 2, size=(rows, n), dtype=np.uint8)


def generate_channel(rng: np.random.Generator, Nt: int, Nr: int, L: int = 3,
                     searching_area: float = SEARCHING_AREA, lam: float = WAVELENGTH,
                     d: float = ANT_SPACING):
    """Eq. 23 channel (Generate_Channel.m:76-139). Returns (H [Nr x Nt], vecH [Nt*Nr], AoD, AoA)."""
    half = searching_area / 2
    aod = rng.uniform(-half, half, L)
    aoa = rng.uniform(-half, half, L)
    g = (rng.standard_normal(L) + 1j * rng.standard_normal(L)) / math.sqrt(2)
    g = g / np.linalg.norm(g)
    kt = np.arange(Nt)[:, None]
    kr = np.arange(Nr)[:, None]
    ATx = np.exp(-1j * 2 * np.pi / lam * d * np.sin(np.deg2rad(aod))[None, :] * kt) / math.sqrt(Nt)
    ARx = np.exp(-1j * 2 * np.pi / lam * d * np.sin(np.deg2rad(aoa))[None, :] * kr) / math.sqrt(Nr)
    H = math.sqrt(Nt * Nr) * (ARx * g[None, :]) @ ATx.conj().T
    return H, H.reshape(-1, order="F"), aod, aoa


@dataclass
class Instance:
    rows: np.ndarray        # int32 [m] codebook row ids (bit-exact bookkeeping)
    A: np.ndarray           # complex128 [m, n] sensing matrix handed to the solver
    B: np.ndarray           # float64 [m] RSS amplitudes |y|
    train_idx: np.ndarray   # int32 [3, floor(m*cc_frac)] the three randsample draws (0-based)
    vecH: np.ndarray        # complex128 [n] ground truth
    snr_db: float


def make_instance(seed: np.random.SeedSequence, cb: np.ndarray, M: int, snr_db: float,
                  Nt: int = 16, Nr: int = 16, L: int = 3, cc_frac: float = 0.95,
                  row_range=None) -> Instance:
    """One (trial, M, SNR) instance of SURVEY.md §8(d) config 1/2/4.

    ``cb`` [rows x n] unit-modulus codebook; sensing rows are ``cb[rows]/sqrt(n)`` (unit-norm rows so
    that signal power is 1, Generate_Measurement.m:84).  ``row_range`` = (lo, hi) restricts the draw to
    a resolution stage of the multires codebook (…simulation_multiresolution.m:137-143).
    """
    rng = np.random.default_rng(seed)
    n = Nt * Nr
    _, vecH, _, _ = generate_channel(rng, Nt, Nr, L)
    lo, hi = (0, cb.shape[0]) if row_range is None else row_range
    rows = (lo + rng.permutation(hi - lo)[:M]).astype(np.int32)
    A = cb[rows, :] / math.sqrt(n)
    noise_power = 10.0 ** (-snr_db / 10.0)
    w = math.sqrt(noise_power / 2) * (rng.standard_normal(M) + 1j * rng.standard_normal(M))
    B = np.abs(A @ vecH + w)
    k = int(math.floor(M * cc_frac))
    train = np.stack([rng.permutation(M)[:k] for _ in range(3)]).astype(np.int32)
    return Instance(rows, A, B, train, vecH, snr_db)


def make_batch(n_trials: int, cb: np.ndarray, M, snr_db, base_seed: int = BASE_SEED,
               first_trial: int = 0, **kw):
    """``n_trials`` instances; trial t uses child t of ``SeedSequence(base_seed)`` regardless of how
    trials are sharded over ranks.  ``M``/``snr_db`` scalar or per-trial sequences."""
    kids = np.random.SeedSequence(base_seed).spawn(first_trial + n_trials)[first_trial:]
    Ms = np.broadcast_to(np.asarray(M), (n_trials,))
    snrs = np.broadcast_to(np.asarray(snr_db, dtype=np.float64), (n_trials,))
    return [make_instance(kids[t], cb, int(Ms[t]), float(snrs[t]), **kw) for t in range(n_trials)]


# ----------------------------------------------------------------------------------- metrics
def nmse(x: np.ndarray, x_gt: np.ndarray) -> float:
    """Scale/phase-invariant MSE_H of Evaluation_H.m:87-89."""
    x = np.asarray(x).reshape(-1)
    x_gt = np.asarray(x_gt).reshape(-1)
    xx = np.vdot(x, x)
    if not np.isfinite(xx) or xx == 0:
        return float("nan")
    return float(np.linalg.norm(x_gt - (np.vdot(x, x_gt) / xx) * x) ** 2 / np.linalg.norm(x_gt) ** 2)


def nmse_db(values) -> float:
    """10*log10(mean MSE_H) (Plot_result.m:154)."""
    return float(10 * np.log10(np.nanmean(np.asarray(values, dtype=np.float64))))


def aligned_rel_err(x: np.ndarray, x_ref: np.ndarray) -> float:
    """||x*e^{j phi} - x_ref|| / ||x_ref|| after the global-phase alignment of Evaluation_H.m:81-82."""
    x = np.asarray(x).reshape(-1)
    x_ref = np.asarray(x_ref).reshape(-1)
    if not (np.all(np.isfinite(x)) and np.all(np.isfinite(x_ref))):
        both_nan = np.array_equal(np.isnan(x), np.isnan(x_ref))
        return 0.0 if both_nan and np.isnan(x).all() else float("inf")
    ph = np.exp(1j * np.angle(np.vdot(x, x_ref)))
    den = np.linalg.norm(x_ref)
    return float(np.linalg.norm(x * ph - x_ref) / (den if den > 0 else 1.0))


# ----------------------------------------------------------------------------------- parity bundles
_BUNDLE_PARAMS = ("lam", "r", "mu0", "rho", "cc_frac", "tol_rel", "tol_abs", "maxiter")


def export_bundle(path: str, A_list, B_list, tx, rx, solver, train_idx_list=None, params=None) -> None:
    """Write a MATLAB v5 .mat bundle that tools/matlab/run_bundle.m replays through the UNMODIFIED reference
    (the only way to pin parity against the reference needs a MATLAB licence; this is the file
    format both sides share).  ``solver``: reference function name(s) ('inferLowRankV4', 'inferLowRankV4_multi',
    'inferLowRank_Nuclear', 'inferLowRankV3', 'inferLowRankV2', 'inferLowRank', 'MyPhaseLift'), one per instance
    or one for all.  ``train_idx_list``: per instance the 0-based randsample draws [T, floor(m*cc_frac)]
    (stored 1-based, drawn order); ``params``: object with the fields of lib.Params (defaults if None).
    For 'MyPhaseLift' B holds the intensities."""
    from scipy.io import savemat
    nb = len(A_list)
    solvers = [solver] * nb if isinstance(solver, str) else list(solver)
    txs = np.broadcast_to(np.asarray(tx, dtype=np.float64), (nb,))
    rxs = np.broadcast_to(np.asarray(rx, dtype=np.float64), (nb,))
    pv = [0.0, 20.0, 1e-3, 1.03, 0.95, 1e-4, 1e-8, 500.0]
    if params is not None:
        pv = [float(getattr(params, k)) for k in _BUNDLE_PARAMS]
    A = np.empty((nb, 1), dtype=object)
    B = np.empty((nb, 1), dtype=object)
    tr = np.empty((nb, 1), dtype=object)
    sv = np.empty((nb, 1), dtype=object)
    for b in range(nb):
        A[b, 0] = np.asarray(A_list[b], dtype=np.complex128)
        B[b, 0] = np.asarray(B_list[b], dtype=np.float64).reshape(-1, 1)
        draws = [] if train_idx_list is None else np.atleast_2d(np.asarray(train_idx_list[b]))
        cell = np.empty((1, len(draws)), dtype=object)
        for t, d in enumerate(draws):
            cell[0, t] = (np.asarray(d, dtype=np.float64) + 1.0).reshape(-1, 1)
        tr[b, 0] = cell
        sv[b, 0] = solvers[b]
    savemat(path, {"A": A, "B": B, "train_idx": tr, "solver": sv, "tx": txs.reshape(-1, 1),
                   "rx": rxs.reshape(-1, 1), "params": np.tile(np.asarray(pv), (nb, 1))}, format="5")


def import_bundle(path: str) -> dict:
    """Read a bundle written by export_bundle back (0-based draws) -- also what tools/pin_parity.py uses."""
    from scipy.io import loadmat
    S = loadmat(path, squeeze_me=False)
    nb = S["A"].shape[0]
    out = {"A": [], "B": [], "train_idx": [], "solver": [], "tx": S["tx"].reshape(-1).astype(int),
           "rx": S["rx"].reshape(-1).astype(int), "params": S["params"]}
    for b in range(nb):
        out["A"].append(np.asarray(S["A"][b, 0], dtype=np.complex128))
        out["B"].append(np.asarray(S["B"][b, 0], dtype=np.float64).reshape(-1))
        cell = S["train_idx"][b, 0]
        draws = [np.asarray(cell[0, t]).reshape(-1).astype(np.int64) - 1 for t in range(cell.shape[1])] \
            if cell.size else []
        out["train_idx"].append(np.array(draws, dtype=np.int32) if draws else np.zeros((0, 0), np.int32))
        out["solver"].append(str(np.asarray(S["solver"][b, 0]).reshape(-1)[0]))
    return out


def import_reference_results(path: str) -> dict:
    """Read the result file saved by tools/matlab/run_bundle.m: X (list of n-vectors), Y, quality."""
    from scipy.io import loadmat
    S = loadmat(path, squeeze_me=False)
    nb = S["X"].shape[0]
    return {"X": [np.asarray(S["X"][b, 0], dtype=np.complex128).reshape(-1) for b in range(nb)],
            "Y": [np.asarray(S["Y"][b, 0], dtype=np.complex128).reshape(-1) for b in range(nb)],
            "quality": np.asarray(S["quality"], dtype=np.float64).reshape(-1),
            "matlab_version": str(np.asarray(S.get("matlab_version", [""])).reshape(-1)[0])}
