"""CPU checks of the PhaseLift oracle (oracle/phaselift.py).  The reference ships no golden vectors for this
path (parity unpinned); these tests anchor the restatement on the mathematics the reference states:
the convex programme of MyPhaseLift.m:78-80, its optimality conditions, and exact recovery."""
import numpy as np
import pytest

from oracle import phaselift as pl


def _gauss(rng, m, n):
    return (rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n))) / np.sqrt(2)


def _aligned_err(a, b):
    ph = np.vdot(a, b)
    ph = ph / abs(ph) if abs(ph) > 0 else 1.0
    return np.linalg.norm(a * ph - b) / np.linalg.norm(b)


def test_operator_adjoint_identity():
    # initializeLinopPR.m:61,65: <A(X), v> == <X, A*(v)>
    rng = np.random.default_rng(0)
    A = _gauss(rng, 12, 7)
    X = _gauss(rng, 7, 7)
    v = rng.standard_normal(12) + 1j * rng.standard_normal(12)
    lhs = np.vdot(pl.lifted_forward(A, X), v)
    rhs = np.vdot(X, pl.lifted_adjoint(A, v))
    assert abs(lhs - rhs) < 1e-12 * abs(lhs)


def test_prox_trace_is_the_minimiser():
    # prox_trace.m:62-158: argmin_{X >= 0} lam*t*trace(X) + 0.5 ||X - W||_F^2
    rng = np.random.default_rng(1)
    W = _gauss(rng, 9, 9)
    W = (W + W.conj().T) / 2
    lam, t = 0.3, 0.7
    val, X, rk = pl.prox_trace(lam, W, t)
    ev = np.linalg.eigvalsh(X)
    assert ev.min() > -1e-13 and rk == int(np.sum(np.linalg.eigvalsh(W) > lam * t))
    assert abs(val - lam * np.real(np.trace(X))) < 1e-13
    obj = lambda Z: lam * t * np.real(np.trace(Z)) + 0.5 * np.linalg.norm(Z - W) ** 2
    base = obj(X)
    for _ in range(50):
        P = _gauss(rng, 9, 3)
        Z = X + 1e-3 * (P @ P.conj().T) * rng.uniform(-1, 1)
        w, V = np.linalg.eigh((Z + Z.conj().T) / 2)
        Z = (V * np.maximum(w, 0)) @ V.conj().T           # feasible neighbour
        assert obj(Z) >= base - 1e-12


def test_noiseless_recovery_and_kkt():
    # MyPhaseLift.m:78-80 on noiseless intensities with m = 8n Gaussian rows: the lifted solution is
    # (nearly) rank one and its leading eigenvector recovers x up to a global phase and the trace bias.
    rng = np.random.default_rng(2)
    n, m = 8, 64
    A = _gauss(rng, m, n)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    y = np.abs(A @ x) ** 2
    tr = pl.TfocsTrace()
    sig, X = pl.my_phase_lift(y, A, pl.TfocsOpts(maxIts=3000), tr, return_matrix=True)
    assert tr.status.startswith("Step size tolerance")
    assert _aligned_err(sig, x) < 2e-2
    # optimality: X >= 0, G = A*(A(X) - y) + lam I >= 0, <X, G> = 0
    lam = pl.TfocsOpts().lam
    G = pl.lifted_adjoint(A, pl.lifted_forward(A, X) - y) + lam * np.eye(n)
    G = (G + G.conj().T) / 2
    scale = np.linalg.norm(G)
    assert np.linalg.eigvalsh(X).min() > -1e-10
    assert np.linalg.eigvalsh(G).min() > -1e-5 * scale
    assert abs(np.real(np.vdot(X, G))) < 1e-5 * scale * np.linalg.norm(X)


def test_restart_and_iteration_limit_bookkeeping():
    rng = np.random.default_rng(3)
    n, m = 6, 30
    A = _gauss(rng, m, n)
    y = np.abs(A @ (rng.standard_normal(n) + 1j * rng.standard_normal(n))) ** 2
    tr = pl.TfocsTrace()
    pl.solver_trace_ls(A, y, pl.TfocsOpts(maxIts=7), tr)
    assert tr.niter == 7 and tr.status == "Iteration limit reached"
    assert tr.n_prox == 7 + tr.n_backtracks and len(tr.L_hist) == 7
    # L decays by alpha between backtracking events (tfocs_AT.m:29)
    for a, b in zip(tr.L_hist[:-1], tr.L_hist[1:]):
        assert b <= a / pl.TfocsOpts().beta + 1e-12 and (b >= 0.9 * a - 1e-12 or b < a)


def test_zero_measurements_give_zero():
    rng = np.random.default_rng(4)
    A = _gauss(rng, 10, 5)
    tr = pl.TfocsTrace()
    sig = pl.my_phase_lift(np.zeros(10), A, pl.TfocsOpts(maxIts=20), tr)
    assert np.all(sig == 0) and tr.niter == 2 and tr.status.startswith("Step size tolerance reached (||dx||=0)")


@pytest.mark.parametrize("m,n", [(10, 24), (20, 32)])
def test_row_space_reduction_is_exact(m, n):
    """The CUDA path iterates on Xr = Qh X Qh' with operator L (A A' = L L', A = L Qh) when m < n.
    With x0 = 0 this is the same iteration: every iterate of the reference lies in range(A') (x) range(A')."""
    rng = np.random.default_rng(5)
    A = _gauss(rng, m, n)
    y = np.abs(A @ (rng.standard_normal(n) + 1j * rng.standard_normal(n))) ** 2
    o = pl.TfocsOpts(maxIts=40)
    t1, t2 = pl.TfocsTrace(), pl.TfocsTrace()
    X = pl.solver_trace_ls(A, y, o, t1)
    Lc = np.linalg.cholesky(A @ A.conj().T)
    Xr = pl.solver_trace_ls(Lc, y, o, t2)
    Qh = np.linalg.solve(Lc, A)                               # orthonormal rows
    assert np.linalg.norm(Qh @ Qh.conj().T - np.eye(m)) < 1e-12
    Xb = Qh.conj().T @ Xr @ Qh
    assert np.linalg.norm(Xb - X) < 1e-10 * np.linalg.norm(X)
    assert t1.niter == t2.niter and t1.n_backtracks == t2.n_backtracks and t1.rank == t2.rank
    assert abs(t1.L - t2.L) < 1e-6 * t1.L      # localL carries the cancellation f_x - q_x (tfocs_backtrack.m:25-26)
