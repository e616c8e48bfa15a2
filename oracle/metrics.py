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
