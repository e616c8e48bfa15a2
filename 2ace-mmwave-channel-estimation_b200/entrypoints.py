"""Host-side mirror of the MATLAB-engine entry points that reach the ADMM solver:

    [H_amp, H_angle] = channel_recovery_ADMM_v2_simulation_X(tx_ant_num, rx_ant_num, cb_amp, cb_angle,
                                                              rss_final, seed_id)

for X in {A2only, A2nuclear, multiresolution} (main/channel_recovery_ADMM_v2_simulation_*.m; called from
main/main.py:231,308,427 through the MATLAB engine).  They are thin wrappers: dBm -> amplitude, row
selection, the 8-value M sweep, output packing; the sweep is submitted to the CUDA library as ONE ragged
batch in codebook mode (A is never materialised per instance).

Not reproduced: MATLAB's RNG stream (randperm / randsample).  The row subsets and train
splits are drawn from a NumPy Generator seeded with the same integer seeds, or passed explicitly.
The `phaselift` entry point (MyPhaseLift through Recover_Channel.m:33-36) and `directional` (PLOMP / PLGAMP: stage I on
the GPU, the SVD reduction and the sparse step II on the host, see twostage.py) are mirrored as well.
"""
from __future__ import annotations

import numpy as np

from . import lib as _lib
from . import solvers as _sv

# channel_recovery_ADMM_v2_simulation_A2only.m:103 (40 hard-coded seeds, indexed by seed_id, 1-based)
A2ONLY_SEEDS = [58659179, 42737934, 36326041, 89830260, 90710947, 96474890, 33424536, 67991541, 42149446,
                38961924, 54659060, 32629256, 33087755, 27433950, 9404442, 20146383, 84040563, 75325961,
                47726929, 13999319, 5597853, 74801351, 37024073, 75534492, 99245881, 19650488, 5314224,
                98859252, 60803022, 76056701, 14112116, 64027813, 73073690, 6288587, 42217659, 45632040,
                7495955, 31960297, 92863244, 93081516]
A2NUCLEAR_SEEDS = [1024, 2048, 4096, 8192]          # …_A2nuclear.m:103
RSS_FCT = 1e5 / 3                                    # …_A2only.m:132
MULTIRES_THRESH = (96, 256)                          # …_multiresolution.m:111
MULTIRES_SEPARATION = (1984, 3968, 3968)             # …_multiresolution.m:112


def matlab_round(x):
    """MATLAB round(): half away from zero (NumPy rounds half to even)."""
    x = np.asarray(x, dtype=np.float64)
    return np.sign(x) * np.floor(np.abs(x) + 0.5)


def measurement_counts(tx_ant_num: int, rx_ant_num: int) -> np.ndarray:
    """M = round(linspace(2, sqrt(4*Nt*Nr), 8)).^2   (…_A2only.m:106-118; main.py:67)."""
    if not any(a in (4, 8, 16, 32, 36) for a in (tx_ant_num, rx_ant_num)):
        raise ValueError("Number of antenna on Tx and Rx must be 4/8/16/32!")
    return (matlab_round(np.linspace(2, np.sqrt(4 * tx_ant_num * rx_ant_num), 8)) ** 2).astype(np.int64)


def rss_dbm_to_amplitude(rss_dbm) -> np.ndarray:
    """sqrt(db2pow(rss) / 1000) * rss_fct   (…_A2only.m:139)."""
    return np.sqrt(10.0 ** (np.asarray(rss_dbm, dtype=np.float64) / 10.0) / 1000.0) * RSS_FCT


def multires_row_range(M: int):
    """0-based half-open codebook row range of the resolution stage used for M probes
    (…_multiresolution.m:137-143): coarse (4 antenna groups), medium (8 groups), full (16 free)."""
    s0, s1, s2 = MULTIRES_SEPARATION
    if M <= MULTIRES_THRESH[0]:
        return 0, s0
    if M <= MULTIRES_THRESH[1]:
        return s0, s0 + s1
    return s0 + s1, s0 + s1 + s2


def _run(variant, tx, rx, cb_amp, cb_angle, rss_final, rng, rows=None, train_idx=None, row_range_fn=None,
         params=None, ctx=None):
    tx, rx = int(tx), int(rx)
    n = tx * rx
    cb = np.asarray(cb_amp, dtype=np.float64) * np.exp(1j * np.asarray(cb_angle, dtype=np.float64))  # :120
    if cb.shape[1] != n:
        raise ValueError(f"codebook has {cb.shape[1]} columns, expected tx*rx = {n}")
    rss = np.asarray(rss_final, dtype=np.float64).reshape(-1)
    Ms = measurement_counts(tx, rx)
    p = params or _lib.Params.default()
    T = 3 if variant == _lib.V4_MULTI else 1
    if rows is None:
        rows = []
        for M in Ms:
            lo, hi = (0, len(rss)) if row_range_fn is None else row_range_fn(int(M))
            if hi > cb.shape[0] or int(M) > hi - lo:
                raise ValueError(f"M = {M} probes requested from rows [{lo},{hi}) of a {cb.shape[0]}-row codebook")
            rows.append((lo + rng.permutation(hi - lo)[:int(M)]).astype(np.int32))          # randperm, :137
    if train_idx is None:
        train_idx = [_sv.draw_train_idx(len(r), p.cc_frac, T, rng) for r in rows]
    B = [rss_dbm_to_amplitude(rss[r]) for r in rows]                                          # :139
    ctx = ctx or _lib.default_context()
    ctx.set_codebook(cb)
    res = _sv.solve_batch_codebook(variant, rows, 1.0, B, tx, rx, train_idx, p, ctx)
    H_out = np.zeros((len(Ms), 1, n), dtype=np.complex128)                                    # Method.Number = 1
    H_out[:, 0, :] = res.X / RSS_FCT                                                          # :170
    H_out[np.isnan(H_out)] = 0                                                                # :176
    return np.abs(H_out), np.angle(H_out), dict(M=Ms, rows=rows, train_idx=train_idx, result=res)


def channel_recovery_ADMM_v2_simulation_A2only(tx_ant_num, rx_ant_num, cb_amp, cb_angle, rss_final, seed_id, *,
                                               rows=None, train_idx=None, params=None, ctx=None, details=False):
    """A2only: inferLowRankV4_multi on randperm rows of the random codebook (…_A2only.m:9-178)."""
    rng = np.random.default_rng(A2ONLY_SEEDS[int(seed_id) - 1])
    amp, ang, info = _run(_lib.V4_MULTI, tx_ant_num, rx_ant_num, cb_amp, cb_angle, rss_final, rng, rows, train_idx,
                          None, params, ctx)
    return (amp, ang, info) if details else (amp, ang)


def channel_recovery_ADMM_v2_simulation_A2nuclear(tx_ant_num, rx_ant_num, cb_amp, cb_angle, rss_final, seed_id, *,
                                                  rows=None, train_idx=None, params=None, ctx=None, details=False):
    """A2nuclear: inferLowRank_Nuclear via Recover_Channel_nuclear (…_A2nuclear.m:103-104,158).  The reference
    picks one of four seeds with an UNSEEDED randi(4) and ignores seed_id (not reproducible even in MATLAB);
    here seed_id selects it deterministically."""
    rng = np.random.default_rng(A2NUCLEAR_SEEDS[(int(seed_id) - 1) % 4])
    amp, ang, info = _run(_lib.NUCLEAR, tx_ant_num, rx_ant_num, cb_amp, cb_angle, rss_final, rng, rows, train_idx,
                          None, params, ctx)
    return (amp, ang, info) if details else (amp, ang)


def channel_recovery_ADMM_v2_simulation_multiresolution(tx_ant_num, rx_ant_num, cb_amp, cb_angle, rss_final,
                                                        seed_id, *, rows=None, train_idx=None, params=None,
                                                        ctx=None, details=False):
    """multiresolution: as A2only, rows drawn from the resolution stage keyed on M (…_multiresolution.m:111-143).
    Like the reference, only defined for 16 antennas (thresh / res_separation exist only in that branch)."""
    if int(tx_ant_num) != 16 and int(rx_ant_num) != 16:
        raise ValueError("multiresolution is only defined for 16 antennas (thresh is undefined otherwise)")
    rng = np.random.default_rng(A2ONLY_SEEDS[int(seed_id) - 1])
    amp, ang, info = _run(_lib.V4_MULTI, tx_ant_num, rx_ant_num, cb_amp, cb_angle, rss_final, rng, rows, train_idx,
                          multires_row_range, params, ctx)
    return (amp, ang, info) if details else (amp, ang)


PHASELIFT_SEED = 4096                                # …_phaselift.m:127 (rng(4096); seed_id is ignored there)


def channel_recovery_ADMM_v2_simulation_phaselift(tx_ant_num, rx_ant_num, cb_amp, cb_angle, rss_final, seed_id=1, *,
                                                  rows=None, opts=None, ctx=None, details=False):
    """phaselift: MyPhaseLift on randperm rows of the codebook (…_phaselift.m:9-178 -> Recover_Channel.m:33-36):
    H = MyPhaseLift((rss_train ./ 2e5).^2 .* 1e10, beams) ./ sqrt(1e10) .* 2e5, then H ./ rss_fct.
    The reference seeds rng(4096) regardless of seed_id."""
    tx, rx = int(tx_ant_num), int(rx_ant_num)
    n = tx * rx
    cb = np.asarray(cb_amp, dtype=np.float64) * np.exp(1j * np.asarray(cb_angle, dtype=np.float64))
    if cb.shape[1] != n:
        raise ValueError(f"codebook has {cb.shape[1]} columns, expected tx*rx = {n}")
    rss = np.asarray(rss_final, dtype=np.float64).reshape(-1)
    Ms = measurement_counts(tx, rx)
    if rows is None:
        rng = np.random.default_rng(PHASELIFT_SEED)
        rows = []
        for M in Ms:
            if int(M) > len(rss) or len(rss) > cb.shape[0]:
                raise ValueError(f"M = {M} probes requested from {len(rss)} RSS entries / {cb.shape[0]} codebook rows")
            rows.append(rng.permutation(len(rss))[:int(M)].astype(np.int32))                  # randperm, :132
    y = [(rss_dbm_to_amplitude(rss[r]) / 2e5) ** 2 * 1e10 for r in rows]                      # Recover_Channel.m:35
    ctx = ctx or _lib.default_context()
    ctx.set_codebook(cb)
    sig, info = _sv.phaselift_batch_codebook(rows, 1.0, y, n, opts, ctx)
    H_out = np.zeros((len(Ms), 1, n), dtype=np.complex128)
    H_out[:, 0, :] = np.asarray(sig) / np.sqrt(1e10) * 2e5 / RSS_FCT                          # Recover_Channel.m:35, :170
    H_out[np.isnan(H_out)] = 0
    amp, ang = np.abs(H_out), np.angle(H_out)
    return (amp, ang, dict(M=Ms, rows=rows, sig=sig, info=info)) if details else (amp, ang)


# ------------------------------------------------------------------------------------------------------------------
DIRECTIONAL_SPACING = 2.9e-3                         # …_directional.m:36 (ULA.d; the other entry points use 3.055e-3)
DIRECTIONAL_L = 3                                    # …_directional.m:52


def sparse_dictionary(Nt: int, Nr: int, NQt: int, NQr: int, spacing: float = DIRECTIONAL_SPACING,
                      wavelength: float = 3e8 / 60.48e9) -> np.ndarray:
    """AD of Sparse_Channel_Formulation.m:83-152 for Searching_Area = 180 (…_directional.m:46): both ends of the area
    map to the first / last grid point, so the dictionary holds every (AoD u, AoA v) pair, u outer:
    AD(:, u * NQr + v) = kron(conj(A_Tx(:, u)), A_Rx(:, v)).  It does not depend on the drawn channel H
    (…_directional.m:148-149 generate one only to obtain this matrix)."""
    kph = 2 * np.pi * spacing / wavelength
    aod_v = kph * np.linspace(-1, 1, NQt + 1)[:-1]
    aoa_v = kph * np.linspace(-1, 1, NQr + 1)[:-1]
    A_Tx = np.exp(-1j * aod_v[None, :] * np.arange(Nt)[:, None]) / np.sqrt(Nt)
    A_Rx = np.exp(-1j * aoa_v[None, :] * np.arange(Nr)[:, None]) / np.sqrt(Nr)
    # kron(conj(a_t), a_r)[kt * Nr + kr] = conj(a_t[kt]) a_r[kr]
    AD = (A_Tx.conj()[:, None, :, None] * A_Rx[None, :, None, :]).reshape(Nt * Nr, NQt * NQr)
    return AD


def directional_indexing(M_cur: int) -> np.ndarray:
    """0-based beam indices along one side of the 32 x 32 directional codebook (…_directional.m:137-141)."""
    k = M_cur if M_cur <= 32 else 32
    return (matlab_round(np.linspace(1, 32, k)) - 1).astype(np.int64)


def channel_recovery_ADMM_v2_simulation_directional(tx_ant_num, rx_ant_num, cb_amp, cb_angle, rss_final, seed_id=1, *,
                                                    opts=None, ctx=None, details=False):
    """directional: PLOMP and PLGAMP (two-stage recovery) on a sub-grid of the 32 x 32 directional codebook
    (main/channel_recovery_ADMM_v2_simulation_directional.m:9-175 -> Recover_Channel.m:37-43 ->
    My_TwoStage_Recovery.m).  cb_amp / cb_angle: 32 x 32 x (tx*rx); rss_final: 32 x 32 dBm.  Output [8, 2, n]:
    method 0 = PLOMP, 1 = PLGAMP.  Stage I (PhaseLift on the mCS x mCS programme) runs on the GPU; the SVD reduction and
    the sparse step II are host code, as they are CPU code in the reference (see twostage.py for what step II is)."""
    from . import twostage as _ts
    tx, rx = int(tx_ant_num), int(rx_ant_num)
    n = tx * rx
    cb = np.asarray(cb_amp, dtype=np.float64) * np.exp(1j * np.asarray(cb_angle, dtype=np.float64))        # :128
    if cb.ndim != 3 or cb.shape[:2] != (32, 32) or cb.shape[2] != n:
        raise ValueError(f"directional codebook must be 32 x 32 x {n}, got {cb.shape}")
    rss = np.asarray(rss_final, dtype=np.float64)
    if rss.shape != (32, 32):
        raise ValueError("rss_final must be 32 x 32 for the directional codebook")
    if not any(a in (8, 16, 17, 32, 36) for a in (tx, rx)):
        raise ValueError("Number of antenna on Tx and Rx must be 4/8/16/32!")
    a_t, a_r = (tx - 1, rx - 1) if 17 in (tx, rx) else (tx, rx)
    Mt = matlab_round(np.linspace(2, np.sqrt(4 * a_t * a_r), 8)).astype(np.int64)                            # :107-125
    AD = sparse_dictionary(tx, rx, 4 * tx, 4 * rx)                                                          # :40-41,149
    H_out = np.zeros((len(Mt), 2, n), dtype=np.complex128)
    info = []
    for i, M_cur in enumerate(Mt):
        idx = directional_indexing(int(M_cur))
        k = len(idx)
        cb_train = cb[np.ix_(idx, idx)]                                                                     # :142
        beams = cb_train.reshape(k * k, n, order="F")                                                       # :143
        rss_train = rss[np.ix_(idx, idx)].reshape(k * k, order="F")                                         # :144-145
        measurements = rss_dbm_to_amplitude(rss_train)                                                      # :146
        y = (measurements / 2e5) ** 2 * 1e10                                                                # Recover_Channel.m:40
        plomp, plgamp, d = _ts.my_two_stage_recovery(y, beams @ AD, DIRECTIONAL_L, opts=opts, ctx=ctx, details=True)
        H_out[i, 0, :] = AD @ plomp / np.sqrt(1e10) * 2e5 / RSS_FCT                                         # :41, :160
        H_out[i, 1, :] = AD @ plgamp / np.sqrt(1e10) * 2e5 / RSS_FCT
        info.append(dict(M=int(M_cur), probes=k * k, mCS=d["mCS"], pl_info=d["info"]))
    H_out[np.isnan(H_out)] = 0                                                                              # :166
    amp, ang = np.abs(H_out), np.angle(H_out)
    return (amp, ang, dict(M=Mt, stages=info)) if details else (amp, ang)
