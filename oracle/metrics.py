"""CPU oracle of the evaluation metrics -- TEST INFRASTRUCTURE ONLY.

Restates Numerical_Simulation/src/evaluate_plot_results/Evaluation_H.m:81-115 and
generate_sensing_matrix/Quantize_PS.m.  PARITY UNPINNED (no reference outputs exist).  One deliberate convention:
MATLAB's svd() returns singular vectors with an implementation-defined common phase (u, v) -> e^{i phi}(u, v);
gain_dig and proj_error do not depend on it, the phase-quantised gain_ana does.  Both this oracle and the CUDA
kernel (csrc/metrics.cuh) fix it canonically: the largest-modulus entry of v (first on ties) is real positive.
"""
from __future__ import annotations

import numpy as np


def quantize_ps(fw: np.ndarray, phase_bit: int) -> np.ndarray:
    """Quantize_PS.m: nearest of phi_qua = -pi : 2*pi/2^bits : pi (first minimum), modulus 1/sqrt(rows)."""
    fw = np.asarray(fw, dtype=np.complex128).reshape(-1)
    nps = 2 ** phase_bit
    phi = -np.pi + (2 * np.pi / nps) * np.arange(nps + 1)
    ind = np.argmin(np.abs(np.angle(fw)[:, None] - phi[None, :]), axis=1)
    return np.exp(1j * phi[ind]) / np.sqrt(fw.size)


def _leading(H: np.ndarray):
    U, S, Vh = np.linalg.svd(H)
    u, v = U[:, 0], Vh[0].conj()
    j = int(np.argmax(np.abs(v)))
    ph = np.conj(v[j]) / abs(v[j]) if abs(v[j]) > 0 else 1.0
    return S[0], u * ph, v * ph


def evaluation_h(recovered: np.ndarray, vecH: np.ndarray, Nt: int, Nr: int, phase_bit: int = 2):
    """(MSE_H, gain_ana, gain_dig, proj_error) of Evaluation_H.m:81-115."""
    x = np.asarray(recovered, dtype=np.complex128).reshape(-1)
    xg = np.asarray(vecH, dtype=np.complex128).reshape(-1)
    if not np.all(np.isfinite(x)) or np.vdot(x, x) == 0:
        return (float("nan"),) * 4
    mse = np.linalg.norm(xg - (np.vdot(x, xg) / np.vdot(x, x)) * x) ** 2 / np.linalg.norm(xg) ** 2      # :87-89
    He = x.reshape(Nr, Nt, order="F")                                                                    # :93
    Ht = xg.reshape(Nr, Nt, order="F")
    se, ue, ve = _leading(He)                                                                            # :94-96
    w, f = quantize_ps(ue, phase_bit), quantize_ps(ve, phase_bit)                                        # :97-98
    gain_ana = abs(np.vdot(w, Ht @ f))                                                                   # :102
    gain_dig = abs(np.vdot(ue / np.linalg.norm(ue), Ht @ (ve / np.linalg.norm(ve))))                     # :103
    st, ut, vt = _leading(Ht)                                                                            # :106
    Xg = (st * np.outer(ut, vt.conj())).reshape(-1, order="F")
    X = (se * np.outer(ue, ve.conj())).reshape(-1, order="F")
    proj = np.linalg.norm(Xg - (np.vdot(X, Xg) / np.vdot(X, X)) * X) / np.linalg.norm(Xg)                # :115
    return float(mse), float(gain_ana), float(gain_dig), float(proj)


# ------------------------------------------------------------------------------------------------------------------
# AoD / AoA estimation error -- Numerical_Simulation/src/evaluate_plot_results/Evaluation_Recovery.m:85-146 on the
# virtual-angle dictionary of generate_channel/Sparse_Channel_Formulation.m:74-152.
def angle_grid(Nt, Nr, NQt, NQr, searching_area, wavelength=3e8 / 60.48e9, spacing=3.055e-3):
    """Sparse_Channel_Formulation.m:83-93,120-135: virtual-angle grids and the (0-based, inclusive) index ranges that
    cover the searching area."""
    kph = 2 * np.pi * spacing / wavelength
    part_t = np.linspace(-1, 1, NQt + 1)[:-1]
    part_r = np.linspace(-1, 1, NQr + 1)[:-1]
    aod_v, aoa_v = kph * part_t, kph * part_r
    rng_v = kph * np.sin(np.deg2rad(np.array([-searching_area / 2, searching_area / 2])))
    u = [int(np.argmin(np.abs(aod_v - x))) for x in rng_v]
    v = [int(np.argmin(np.abs(aoa_v - x))) for x in rng_v]
    return kph, aod_v, aoa_v, (u[0], u[1]), (v[0], v[1])


def angular_spectrum(x_est, Nt, Nr, NQt, NQr, searching_area, **kw):
    """z_leakage_reduced of the ESTIMATED channel (Sparse_Channel_Formulation.m:96-103,137-152: z = vec(A_Rx' H A_Tx)
    restricted to the searching area, AoD index outer, AoA index inner): the `recoveredSig` that
    Evaluation_Recovery.m expects, formed from an H-domain estimate such as the ADMM solvers return."""
    kph, aod_v, aoa_v, (u0, u1), (v0, v1) = angle_grid(Nt, Nr, NQt, NQr, searching_area, **kw)
    A_Tx = np.exp(-1j * aod_v[None, :] * np.arange(Nt)[:, None]) / np.sqrt(Nt)
    A_Rx = np.exp(-1j * aoa_v[None, :] * np.arange(Nr)[:, None]) / np.sqrt(Nr)
    H = np.asarray(x_est, dtype=np.complex128).reshape(Nr, Nt, order="F")
    Z = A_Rx.conj().T @ H @ A_Tx                      # NQr x NQt
    return Z[v0:v1 + 1, u0:u1 + 1].reshape(-1, order="F")


def evaluation_angles(x_est, aod_true, aoa_true, Nt, Nr, NQt=None, NQr=None, searching_area=95.0, **kw):
    """Evaluation_Recovery.m:85-146.  Returns (AoD_Err_to_True, AoA_Err_to_True, AoDA_Err, AoD_Err_to_True_Quantized,
    AoA_Err_to_True_Quantized, AoDA_Err_Quantized) in degrees.  Quirk kept: :133 `AoA_True(order_True) =
    AoA_True(order_True)` is a no-op, so the true AoAs stay in path order while the true AoDs are sorted."""
    NQt = 4 * Nt if NQt is None else NQt              # Vs_M_par.m:80-81
    NQr = 4 * Nr if NQr is None else NQr
    aod_true = np.asarray(aod_true, dtype=np.float64)
    aoa_true = np.asarray(aoa_true, dtype=np.float64)
    L = aod_true.size
    kph, aod_v, aoa_v, (u0, u1), (v0, v1) = angle_grid(Nt, Nr, NQt, NQr, searching_area, **kw)
    sig = angular_spectrum(x_est, Nt, Nr, NQt, NQr, searching_area, **kw)
    if not np.all(np.isfinite(sig)):
        return (np.nan,) * 6
    n_aoa = v1 - v0 + 1
    ind = np.argsort(-np.abs(sig), kind="stable")[:L]          # :86 sort(...,'descend') (stable; complex: by modulus)
    ind = np.sort(ind)                                         # :88
    aod_est = np.rad2deg(np.arcsin(aod_v[u0 + ind // n_aoa] / kph))          # :104-113
    aoa_est = np.rad2deg(np.arcsin(aoa_v[v0 + ind % n_aoa] / kph))
    pos_d = [int(np.argmin(np.abs(aod_v - kph * np.sin(np.deg2rad(a))))) for a in aod_true]   # Quan_Pos_Err_h_array
    pos_a = [int(np.argmin(np.abs(aoa_v - kph * np.sin(np.deg2rad(a))))) for a in aoa_true]
    aod_tq = np.rad2deg(np.arcsin(aod_v[pos_d] / kph))                       # :115-123
    aoa_tq = np.rad2deg(np.arcsin(aoa_v[pos_a] / kph))
    order_true = np.argsort(-aod_true, kind="stable")                        # :132
    aod_t = aod_true[order_true]
    aoa_t = aoa_true                                                         # :133 (no-op in the reference)
    aod_tq, aoa_tq = aod_tq[order_true], aoa_tq[order_true]                  # :134-135
    order_est = np.argsort(-aod_est, kind="stable")                          # :136-137
    aod_est, aoa_est = aod_est[order_est], aoa_est[order_est]
    e_dq, e_aq = np.mean(np.abs(aod_est - aod_tq)), np.mean(np.abs(aoa_est - aoa_tq))       # :140-141
    e_d, e_a = np.mean(np.abs(aod_est - aod_t)), np.mean(np.abs(aoa_est - aoa_t))           # :142-143
    if Nt == 1:
        e = e_a
    elif Nr == 1:
        e = e_d
    else:
        e = 0.5 * (e_d + e_a)                                                # :144-151
    return float(e_d), float(e_a), float(e), float(e_dq), float(e_aq), float(0.5 * (e_dq + e_aq))
