"""Two-stage recovery (PLOMP / PLGAMP) behind the `directional` entry point -- host side of
main/src/my_recovery_algorithms/My_TwoStage_Recovery.m:75-152, the way the reference runs it in MATLAB:

* the SVD reduction of the lifted sensing matrix to mCS dimensions (:77-101) and the sparse step II stay on the CPU
  (the reference's OMP / EM-BG-GAMP are third-party CPU code);
* stage I, the trace-regularised least squares over mCS x mCS PSD matrices (:115-129, solver_TraceLS with the operator
  P = U sqrt(S)), is MyPhaseLift's programme at dimension mCS and runs on the GPU through twoace_phaselift_batch.

Step II.  `OMP(A, y, 1e-12)` (:133) is not vendored in the reference tree (no OMP.m anywhere under main/ or
Numerical_Simulation/); it is restated here as textbook orthogonal matching pursuit with a residual-energy stop.
EM-BG-GAMP (:139-150) is a large third-party package and outside this build: PLGAMP takes the reference's own
fallback branch (:146-149, `catch ... OMP`), so both outputs coincide.
"""
from __future__ import annotations

import math

import numpy as np

from . import lib as _lib
from . import solvers as _sv


def svd_reduction(measurementMat, s: int):
    """My_TwoStage_Recovery.m:77-101: economy SVD, mCS from the 80 %-energy and 1.75 mCS log(mCS) >= m rules.
    Returns (P [m, mCS], C [mCS, n], mCS)."""
    A = np.asarray(measurementMat, dtype=np.complex128)
    m, n = A.shape
    mcs_required = int(_matlab_round(1.75 * s * math.log(n / s)))
    U, dS, Vh = np.linalg.svd(A, full_matrices=False)
    mcs = min(mcs_required, dS.size) - 1
    var_ratio = dS[:mcs].sum() / dS.sum()
    while var_ratio < 0.80 and mcs < dS.size:
        mcs += 1
        var_ratio = dS[:mcs].sum() / dS.sum()
    while _matlab_round(1.75 * mcs * math.log(mcs)) < m and mcs < dS.size:
        mcs += 1
    sq = np.sqrt(dS[:mcs])
    return U[:, :mcs] * sq[None, :], sq[:, None] * Vh[:mcs, :], mcs


def _matlab_round(x: float) -> float:
    return math.copysign(math.floor(abs(x) + 0.5), x)


def omp(A, y, tol: float = 1e-12, max_atoms: int | None = None):
    """Orthogonal matching pursuit: greedy support by the largest correlation with the residual, least squares on the
    support, stop when ||residual||^2 <= tol or the support is full (rank of A)."""
    A = np.asarray(A, dtype=np.complex128)
    y = np.asarray(y, dtype=np.complex128).reshape(-1)
    m, n = A.shape
    kmax = min(m, n) if max_atoms is None else min(max_atoms, m, n)
    x = np.zeros(n, dtype=np.complex128)
    support: list[int] = []
    r = y.copy()
    coef = np.zeros(0, dtype=np.complex128)
    while len(support) < kmax and np.vdot(r, r).real > tol:
        corr = np.abs(A.conj().T @ r)
        corr[support] = -1.0
        support.append(int(np.argmax(corr)))
        coef, *_ = np.linalg.lstsq(A[:, support], y, rcond=None)
        r = y - A[:, support] @ coef
    x[support] = coef
    return x


def my_two_stage_recovery(measurements, measurementMat_Input, s: int, *, opts: "_lib.PlOpts | None" = None, ctx=None,
                          details: bool = False):
    """[recoveredSig_PLOMP, recoveredSig_PLGAMP] = My_TwoStage_Recovery(measurements, measurementMat_Input, s, ...).
    `measurements` are intensities."""
    P, Cm, mcs = svd_reduction(measurementMat_Input, s)
    # (mCS > 256 runs in the row space of P and needs at most 256 measurements: twoace_phaselift_batch raises otherwise)
    sig, info = _sv.phaselift_batch([P], [np.asarray(measurements, dtype=np.float64).reshape(-1)], opts, ctx)   # :115-129
    int_soln = sig[0]
    plomp = omp(Cm, int_soln, 1e-12)                                                                            # :133
    plgamp = plomp.copy()                                                                                        # :146-149
    if details:
        return plomp, plgamp, dict(mCS=mcs, P=P, C=Cm, intSoln=int_soln, info=info)
    return plomp, plgamp
