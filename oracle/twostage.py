"""CPU oracle of the two-stage recovery behind the `directional` entry point -- TEST INFRASTRUCTURE ONLY.

Restates main/src/my_recovery_algorithms/My_TwoStage_Recovery.m:75-152 with stage I solved by oracle/phaselift.py
(solver_TraceLS on the mCS x mCS programme, :115-129).  PARITY UNPINNED.  Step II: `OMP` is not vendored in the reference
(restated as textbook OMP), EM-BG-GAMP is replaced by the reference's own fallback branch (:146-149).
"""
from __future__ import annotations

import math

import numpy as np

from . import phaselift as opl


def _mround(x):
    return math.copysign(math.floor(abs(x) + 0.5), x)


def svd_reduction(A, s):
    """:77-101."""
    A = np.asarray(A, dtype=np.complex128)
    m, n = A.shape
    U, dS, Vh = np.linalg.svd(A, full_matrices=False)
    mcs = min(int(_mround(1.75 * s * math.log(n / s))), dS.size) - 1
    while dS[:mcs].sum() / dS.sum() < 0.80 and mcs < dS.size:
        mcs += 1
    while _mround(1.75 * mcs * math.log(mcs)) < m and mcs < dS.size:
        mcs += 1
    sq = np.sqrt(dS[:mcs])
    return U[:, :mcs] * sq[None, :], sq[:, None] * Vh[:mcs, :], mcs


def omp(A, y, tol=1e-12):
    A = np.asarray(A, dtype=np.complex128)
    y = np.asarray(y, dtype=np.complex128).reshape(-1)
    m, n = A.shape
    x = np.zeros(n, dtype=np.complex128)
    sup, r, coef = [], y.copy(), np.zeros(0, complex)
    while len(sup) < min(m, n) and np.vdot(r, r).real > tol:
        c = np.abs(A.conj().T @ r)
        c[sup] = -1.0
        sup.append(int(np.argmax(c)))
        coef = np.linalg.lstsq(A[:, sup], y, rcond=None)[0]
        r = y - A[:, sup] @ coef
    x[sup] = coef
    return x


def my_two_stage_recovery(measurements, A, s, opts=None):
    P, C, mcs = svd_reduction(A, s)
    int_soln = opl.my_phase_lift(np.asarray(measurements, dtype=np.float64).reshape(-1), P, opts)
    plomp = omp(C, int_soln, 1e-12)
    return plomp, plomp.copy(), dict(mCS=mcs, intSoln=int_soln, P=P, C=C)
