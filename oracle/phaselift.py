"""CPU oracle of the reference's PhaseLift path -- TEST INFRASTRUCTURE ONLY.

NumPy complex128 restatement of
    main/src/my_recovery_algorithms/MyPhaseLift.m:69-107
      -> sparsepr/src/initializeLinopPR.m:50-66           (lifted operator and adjoint)
      -> third/TFOCS/solver_TraceLS.m:23-42               (smooth_quad o (A, -b) + prox_trace)
      -> third/TFOCS/tfocs_AT.m:20-94                     (Auslender-Teboulle accelerated loop)
      -> third/TFOCS/private/tfocs_initialize.m:9-40,418-426,462-478,521,550,582-598
      -> third/TFOCS/private/tfocs_backtrack.m:4-46, tfocs_iterate.m:8-33,326-362, tfocs_cleanup.m:28-52
      -> third/TFOCS/prox_trace.m:62-168, smooth_quad.m:149-155
(TFOCS v1.3/1.4 is vendored in the reference tree under main/3rd_software_component/sparsepr/third/TFOCS.)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; the product path never does.

PARITY UNPINNED: the reference ships no golden vectors for this path and neither MATLAB nor Octave is
available, so this restatement cannot be checked against reference outputs.  Anchors used instead
(tests/test_oracle_phaselift.py): the optimality conditions of the convex programme the reference states
(MyPhaseLift.m:78-80), exact recovery of a rank-one lifted signal from noiseless intensities, and the
invariance of the iteration under the row-space reduction the CUDA path uses.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

EPS = float(np.finfo(np.float64).eps)


@dataclass
class TfocsOpts:
    """MyPhaseLift.m:83-92 on top of the tfocs_initialize.m:9-40 defaults."""
    maxIts: int = 4000            # MyPhaseLift.m:83
    tol: float = 1e-10            # MyPhaseLift.m:84
    restart: int = 200            # MyPhaseLift.m:85
    lam: float = 5e-2             # MyPhaseLift.m:92
    alpha: float = 0.9            # tfocs_initialize.m:22
    beta: float = 0.5             # tfocs_initialize.m:21
    L0: float = 1.0               # tfocs_initialize.m:23,186-195 (Lexact = Inf)
    cntr_reset: int = 50          # tfocs_initialize.m:31,201-207 (tol >= 1e-12)
    backtrack_tol: float = 1e-10  # tfocs_initialize.m:425


@dataclass
class TfocsTrace:
    niter: int = 0
    status: str = ""
    n_prox: int = 0               # prox_trace evaluations (= eigendecompositions)
    n_backtracks: int = 0
    L: float = 0.0
    rank: int = 0                 # rank of the last prox output
    L_hist: list = field(default_factory=list)
    rank_hist: list = field(default_factory=list)


def lifted_forward(A: np.ndarray, X: np.ndarray) -> np.ndarray:
    """initializeLinopPR.m:61  y = diag(PhMat*X*PhMat') (complex, as MATLAB keeps it)."""
    return np.einsum("ij,ij->i", A @ X, A.conj())


def lifted_adjoint(A: np.ndarray, v: np.ndarray) -> np.ndarray:
    """initializeLinopPR.m:65  PhMat' * diag(v) * PhMat."""
    return A.conj().T @ (v[:, None] * A)


def tfocs_dot(x: np.ndarray, y: np.ndarray) -> float:
    """@double/tfocs_dot.m: real part of the inner product (complex branch), plain product when real."""
    return float(np.real(np.vdot(x, y)))


def prox_trace(lam: float, X: np.ndarray, t: float):
    """prox_trace.m:69-158 (dense branch): eig of the Hermitian part, shrink by lam*t, keep the positive part.
    Returns (value, X_new, rank)."""
    tau = lam * t                                            # :77
    Xh = (X + X.conj().T) / 2                                # :92
    D, V = np.linalg.eigh(Xh)
    s = D - tau                                              # :140
    tt = s > 0                                               # :141
    s = s[tt]
    if s.size == 0:                                          # :144-145
        return 0.0, np.zeros_like(X), 0
    Vt = V[:, tt]
    Xn = Vt @ (s[:, None] * Vt.conj().T)                     # :147
    Xn = (Xn + Xn.conj().T) / 2                              # :149
    return lam * float(np.sum(s)), Xn, int(s.size)           # :152


def trace_value(lam: float, X: np.ndarray) -> float:
    """prox_trace.m:160-161 (no projection requested)."""
    return lam * float(np.real(np.trace(X + X.conj().T))) / 2


def solver_trace_ls(A: np.ndarray, b: np.ndarray, opts: TfocsOpts | None = None, trace: TfocsTrace | None = None,
                    x0: np.ndarray | None = None) -> np.ndarray:
    """solver_TraceLS.m:42 = tfocs_AT(smooth_quad, {A,-b}, prox_trace(lambda), x0, opts) with the options of
    MyPhaseLift.m:83-95.  Follows tfocs_AT.m:20-94 statement by statement; comments cite the helper scripts."""
    o = opts or TfocsOpts()
    tr = trace if trace is not None else TfocsTrace()
    A = np.asarray(A, dtype=np.complex128)
    b = np.asarray(b)
    n = A.shape[1]

    def smooth(Ax):
        # smooth_quad_simple o (. - b): tfocs_initialize.m:342, smooth_quad.m:149-155
        g = Ax - b
        return 0.5 * tfocs_dot(g, g), g

    # ---- tfocs_initialize.m:418-478, 521, 582-592 (x0 = zeros(n): MyPhaseLift.m:95)
    L = o.L0
    theta = np.inf
    if x0 is None or not np.any(x0):
        x = np.zeros((n, n), dtype=np.complex128)
        A_x = np.zeros(A.shape[0], dtype=np.complex128)      # :476 (zero_x0, non-square operator)
    else:
        x = np.array(x0, dtype=np.complex128)
        A_x = lifted_forward(A, x)                           # :474
    C_x = trace_value(o.lam, x)                              # :464
    f_x, g_Ax = smooth(A_x)                                  # :521
    g_x = None
    y, z = x, x
    A_y, A_z = A_x, A_x
    f_y, g_Ay, g_y = f_x, g_Ax, g_x
    cntr_Ay = cntr_Ax = 0
    backtrack_simple = True
    backtrack_steps = 0
    restart_iter = 0
    n_iter = 0
    status = ""
    xy_sq = 0.0

    while True:                                              # tfocs_AT.m:20
        x_old, A_x_old, z_old, A_z_old = x, A_x, z, A_z      # :22-25
        L_old = L                                            # :28
        L = L * o.alpha                                      # :29
        theta_old = theta                                    # :30
        while True:                                          # :31 backtracking loop
            # :34, tfocs_initialize.m:550
            theta = 2.0 / (1.0 + np.sqrt(1.0 + 4.0 * (L / L_old) / theta_old ** 2)) if np.isfinite(theta_old) else 1.0
            if theta < 1:                                    # :37-50
                y = (1 - theta) * x_old + theta * z_old
                if cntr_Ay >= o.cntr_reset:
                    A_y = lifted_forward(A, y)
                    cntr_Ay = 0
                else:
                    cntr_Ay += 1
                    A_y = (1 - theta) * A_x_old + theta * A_z_old
                f_y, g_Ay, g_y = np.inf, None, None
            if g_y is None:                                  # :53-56
                if g_Ay is None:
                    f_y, g_Ay = smooth(A_y)
                g_y = lifted_adjoint(A, g_Ay)
            step = 1.0 / (theta * L)                         # :59
            C_z, z, rk = prox_trace(o.lam, z_old - step * g_y, step)   # :60
            tr.n_prox += 1
            A_z = lifted_forward(A, z)                       # :61
            if theta == 1:                                   # :64-78
                x, A_x, C_x = z, A_z, C_z
            else:
                x = (1 - theta) * x_old + theta * z
                if cntr_Ax >= o.cntr_reset:
                    cntr_Ax = 0
                    A_x = lifted_forward(A, x)
                else:
                    cntr_Ax += 1
                    A_x = (1 - theta) * A_x_old + theta * A_z
                C_x = np.inf
            f_x, g_Ax, g_x = np.inf, None, None              # :79
            # ---- tfocs_backtrack.m:4-46
            if o.beta >= 1:
                break
            xy = x - y
            xy_sq = tfocs_dot(xy, xy)
            if xy_sq == 0:
                break
            if xy_sq / tfocs_dot(x, x) < EPS:
                cntr_Ax = np.inf
            if backtrack_simple:
                if np.isinf(f_x):
                    f_x, _ = smooth(A_x)
                q_x = f_y + tfocs_dot(xy, g_y) + 0.5 * L * xy_sq
                localL = L + 2 * max(f_x - q_x, 0.0) / xy_sq
                backtrack_simple = abs(f_y - f_x) >= o.backtrack_tol * max(abs(f_x), abs(f_y))
            else:
                if g_Ax is None:
                    f_x, g_Ax = smooth(A_x)
                localL = 2 * tfocs_dot(A_x - A_y, g_Ax - g_Ay) / xy_sq
            backtrack_steps += 1
            if localL <= L:                                  # :37 (Lexact = Inf)
                break
            if not np.isinf(localL):
                L = localL
            else:
                localL = L
            L = max(localL, L / o.beta)                      # :43
            tr.n_backtracks += 1
        # ---- tfocs_iterate.m:8-33
        n_iter += 1
        norm_x = np.sqrt(tfocs_dot(x, x))
        dx = x - x_old
        norm_dx = np.sqrt(tfocs_dot(dx, dx))
        if np.isnan(f_y):
            status = "NaN found -- aborting"
        elif norm_dx == 0:
            if n_iter > 1:
                status = "Step size tolerance reached (||dx||=0)"
        elif norm_dx < o.tol * max(norm_x, 1.0):
            status = "Step size tolerance reached"
        elif n_iter == o.maxIts:
            status = "Iteration limit reached"
        elif backtrack_steps > 0 and xy_sq == 0:
            status = "Unexpectedly small stepsize"
        tr.L_hist.append(L)
        tr.rank_hist.append(rk)
        if status:
            break
        # ---- tfocs_iterate.m:331-362 (restart > 0: fixed-period restart only)
        backtrack_steps = 0
        if n_iter - restart_iter == abs(round(o.restart)):
            restart_iter = n_iter
            backtrack_simple = True
            theta = np.inf
            y, A_y, f_y, g_Ay, g_y = x, A_x, f_x, g_Ax, g_x
            z, A_z = x, A_x
    # ---- tfocs_cleanup.m:28-52: the x sequence is returned (v_is_y is never set without stopFcn)
    tr.niter, tr.status, tr.L, tr.rank = n_iter, status, L, rk
    return x


def my_phase_lift(measurements: np.ndarray, A: np.ndarray, opts: TfocsOpts | None = None,
                  trace: TfocsTrace | None = None, return_matrix: bool = False):
    """MyPhaseLift.m:69-107.  `measurements` are intensities (|y|^2; the squaring is caller-side,
    Recover_Channel.m:35).  Returns sqrt(largest eigenvalue) * leading eigenvector (global phase arbitrary)."""
    X = solver_trace_ls(A, np.asarray(measurements).reshape(-1), opts, trace)
    D, V = np.linalg.eigh((X + X.conj().T) / 2)             # :106 (X is Hermitian already)
    sig = np.sqrt(D[-1] + 0j) * V[:, -1]                     # :107
    return (sig, X) if return_matrix else sig
