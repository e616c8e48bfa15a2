"""GPU parity of the PhaseLift path (C ABI twoace_phaselift_batch) against oracle/phaselift.py.

Tolerances (FP64 both sides, result compared after global-phase alignment, Evaluation_H.m:81-82):
  * bounded iteration counts (1 ... 40 TFOCS iterations, identical inputs): relative error <= 1e-9 and the
    integer bookkeeping (iterations, prox evaluations, backtracking steps, rank, status) bit-exact, L to 1e-6
    (L carries the cancellation f_x - q_x of tfocs_backtrack.m:25-26);
  * converged solves (MyPhaseLift.m defaults): relative error <= 1e-4 (BASELINE north-star bar; measured 1e-8).
    The iteration at which TFOCS stops is decided by rounding noise in the reference itself: after
    |f_y - f_x| < 1e-10 max(|f_x|,|f_y|) the backtracking estimate localL = 2<A_x - A_y, g_Ax - g_Ay>/|x - y|^2
    is a 0/0 of rounding residues, L jumps by orders of magnitude and the step-size test fires
    (tfocs_backtrack.m:27-32, tfocs_iterate.m:23).  The programme is convex, so the result is not affected
    beyond the stopping tolerance; iteration counts of converged runs are therefore not compared.
"""
import numpy as np
import pytest

from oracle import phaselift as opl

pytestmark = pytest.mark.gpu


def _gauss(rng, m, n):
    return (rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n))) / np.sqrt(2)


def _err(a, b):
    from twoace_b200 import harness as hz
    return hz.aligned_rel_err(a, b)


def _problem(seed, m, n):
    rng = np.random.default_rng(seed)
    A = _gauss(rng, m, n)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    return A, np.abs(A @ x) ** 2, x


def _both(A, y, gpu_ctx, **kw):
    import twoace_b200 as tw
    tr = opl.TfocsTrace()
    okw = {k: v for k, v in kw.items() if k != "reduce"}
    ref = opl.my_phase_lift(y, A, opl.TfocsOpts(**okw), tr)
    sig, info = tw.phaselift_batch([A], [y], tw.PlOpts.default(**kw), gpu_ctx)
    return ref, tr, sig[0], info[0]


def _check_counts(tr, info):
    assert int(info[0]) == tr.niter and int(info[1]) == tr.n_prox and int(info[2]) == tr.n_backtracks
    assert int(info[4]) == tr.rank
    assert abs(info[5] - tr.L) <= 1e-6 * tr.L
    code = {"Step size tolerance reached": 1, "Iteration limit reached": 2,
            "Step size tolerance reached (||dx||=0)": 3}[tr.status]
    assert int(info[3]) == code


@pytest.mark.parametrize("its", [1, 2, 5, 30])
@pytest.mark.parametrize("shape,reduce", [((96, 16), 1), ((24, 40), 1), ((24, 40), 0), ((40, 40), 1), ((70, 33), 1)])
def test_bounded_iterations_parity(gpu_ctx, its, shape, reduce):
    m, n = shape
    A, y, _ = _problem(11 + m, m, n)
    ref, tr, sig, info = _both(A, y, gpu_ctx, maxIts=its, reduce=reduce)
    assert _err(sig, ref) < 1e-9
    _check_counts(tr, info)
    assert int(info[6]) == (m if (reduce and m < n) else n)


def test_codebook_instance_parity(codebook, gpu_ctx):
    """16x16 codebook instance of BASELINE config 3 (M = 128 -> 128-dimensional reduced iteration), 40 iterations;
    dense, unreduced and codebook-row entry points agree."""
    import twoace_b200 as tw
    from twoace_b200 import harness as hz
    ins = hz.make_batch(1, codebook, 128, 20.0)[0]
    y = (ins.B / 2.0) ** 2                      # Recover_Channel.m:35 scaling, caller-side
    ref, tr, sig, info = _both(ins.A, y, gpu_ctx, maxIts=40)
    assert _err(sig, ref) < 1e-8
    _check_counts(tr, info)
    assert int(info[6]) == 128
    sig_full, info_full = tw.phaselift_batch([ins.A], [y], tw.PlOpts.default(maxIts=40, reduce=0), gpu_ctx)
    assert _err(sig_full[0], ref) < 1e-8 and int(info_full[0][6]) == 256
    gpu_ctx.set_codebook(codebook)
    sig_cb, info_cb = tw.phaselift_batch_codebook([ins.rows], 1.0 / 16.0, [y], 256, tw.PlOpts.default(maxIts=40),
                                                  gpu_ctx)
    assert _err(sig_cb[0], ref) < 1e-8
    assert np.array_equal(info_cb[0][:5], info[:5])


@pytest.mark.parametrize("shape", [(64, 8), (60, 12)])
def test_converged_solve_parity(gpu_ctx, shape):
    m, n = shape
    A, y, x = _problem(3 + n, m, n)
    ref, tr, sig, info = _both(A, y, gpu_ctx)
    assert tr.status.startswith("Step size tolerance") and int(info[3]) in (1, 3)
    assert _err(sig, ref) < 1e-4
    if m >= 5 * n:
        assert _err(sig, x) < 2e-2               # noiseless, well over-sampled: PhaseLift recovers x


def test_converged_codebook_solve_parity(codebook, gpu_ctx):
    """BASELINE config 3 instance (16x16, M = 128, 20 dB) with the MyPhaseLift.m defaults: both sides stop on the
    step-size test after ~1000-1400 iterations (at different iterations, see the module docstring) at the same
    minimiser.  Oracle self-sensitivity to a 1e-14 perturbation of y: 8e-8."""
    from twoace_b200 import harness as hz
    ins = hz.make_batch(1, codebook, 128, 20.0)[0]
    ref, tr, sig, info = _both(ins.A, (ins.B / 2.0) ** 2, gpu_ctx)
    assert tr.status == "Step size tolerance reached" and int(info[3]) == 1
    assert _err(sig, ref) < 1e-4
    assert int(info[4]) == tr.rank


def test_iteration_limited_underdetermined_solve(gpu_ctx):
    """m < n Gaussian rows do not converge within maxIts = 4000; the reference's own result then moves by 7e-4
    under a 1e-14 perturbation of y (backtracking decisions on rounding residues).  The GPU must stay within
    100x that self-sensitivity and run the same 4000 iterations."""
    A, y, _ = _problem(33, 20, 30)
    ref, tr, sig, info = _both(A, y, gpu_ctx)
    rng = np.random.default_rng(0)
    ref2 = opl.my_phase_lift(y * (1 + 1e-14 * rng.standard_normal(y.size)), A)
    self_sens = _err(ref2, ref)
    assert tr.niter == 4000 and int(info[0]) == 4000 and int(info[3]) == 2
    assert _err(sig, ref) < max(1e-4, 100 * self_sens)


def test_ragged_batch_matches_single_solves_bitwise(gpu_ctx):
    import twoace_b200 as tw
    probs = [_problem(100 + i, m, n=24) for i, m in enumerate([10, 24, 31, 17, 60, 12])]
    o = tw.PlOpts.default(maxIts=25)
    sig_b, info_b = tw.phaselift_batch([p[0] for p in probs], [p[1] for p in probs], o, gpu_ctx)
    for i, p in enumerate(probs):
        s1, i1 = tw.phaselift_batch([p[0]], [p[1]], o, gpu_ctx)
        assert np.array_equal(s1[0], sig_b[i]) and np.array_equal(i1[0][:9], info_b[i][:9])
        ref = opl.my_phase_lift(p[1], p[0], opl.TfocsOpts(maxIts=25))
        assert _err(sig_b[i], ref) < 1e-9


def test_rank_deficient_rows_fall_back_to_full_dimension(gpu_ctx):
    A, y, _ = _problem(7, 12, 20)
    A = np.vstack([A, A[:3]])                   # duplicated rows: A A' singular, no Cholesky factor
    y = np.concatenate([y, y[:3]])
    ref, tr, sig, info = _both(A, y, gpu_ctx, maxIts=20)
    assert int(info[6]) == 20
    assert _err(sig, ref) < 1e-7                # singular A A': the problem itself is ill-conditioned
    _check_counts(tr, info)


@pytest.mark.parametrize("shape,its", [((48, 300), 1), ((48, 300), 5), ((48, 300), 30), ((200, 640), 12),
                                       ((120, 1024), 3)])
def test_wide_problem_runs_in_the_row_space(gpu_ctx, shape, its):
    """n > 256 (up to the 32x32 array, n = 1024): the GPU iterates in the m-dimensional row space of A, the oracle on
    the full n x n matrix (initializeLinopPR.m:61-77); bounded iteration counts, same bars as the n <= 256 cases."""
    m, n = shape
    A, y, _ = _problem(500 + m, m, n)
    ref, tr, sig, info = _both(A, y, gpu_ctx, maxIts=its)
    assert sig.shape == (n,)
    assert _err(sig, ref) < 1e-9
    _check_counts(tr, info)
    assert int(info[6]) == m


def test_wide_problem_limits(gpu_ctx):
    import twoace_b200 as tw
    A, y, _ = _problem(77, 20, 300)
    with pytest.raises(tw.TwoaceError, match="row space"):
        tw.phaselift_batch([A], [y], tw.PlOpts.default(maxIts=3, reduce=0), gpu_ctx)
    A2, y2, _ = _problem(78, 257, 300)
    with pytest.raises(tw.TwoaceError, match="row space"):
        tw.phaselift_batch([A2], [y2], tw.PlOpts.default(maxIts=3), gpu_ctx)
    Ad = np.vstack([A, A[:2]])                  # dependent rows: no Cholesky factor, and d = n = 300 is not built
    with pytest.raises(tw.TwoaceError, match="linearly dependent"):
        tw.phaselift_batch([Ad], [np.concatenate([y, y[:2]])], tw.PlOpts.default(maxIts=3), gpu_ctx)
    # a ragged wide batch equals the single solves bitwise
    probs = [_problem(300 + i, mm, 400) for i, mm in enumerate([12, 90, 33])]
    o = tw.PlOpts.default(maxIts=10)
    sig_b, info_b = tw.phaselift_batch([q[0] for q in probs], [q[1] for q in probs], o, gpu_ctx)
    for i, q in enumerate(probs):
        s1, i1 = tw.phaselift_batch([q[0]], [q[1]], o, gpu_ctx)
        assert np.array_equal(s1[0], sig_b[i]) and np.array_equal(i1[0][:9], info_b[i][:9])


def test_zero_measurements(gpu_ctx):
    A, _, _ = _problem(9, 10, 5)
    ref, tr, sig, info = _both(A, np.zeros(10), gpu_ctx, maxIts=20)
    assert np.all(sig == 0) and np.all(ref == 0)
    assert int(info[0]) == tr.niter == 2 and int(info[3]) == 3


def test_matlab_signature_and_errors(gpu_ctx):
    import twoace_b200 as tw
    A, y, _ = _problem(21, 30, 6)
    sig = tw.MyPhaseLift(y, A, opts=tw.PlOpts.default(maxIts=10), ctx=gpu_ctx)
    ref = opl.my_phase_lift(y, A, opl.TfocsOpts(maxIts=10))
    assert sig.shape == (6,) and _err(sig, ref) < 1e-9
    with pytest.raises(tw.TwoaceError):
        tw.phaselift_batch([np.zeros((4, 2001), complex)], [np.zeros(4)], None, gpu_ctx)     # the largescale range
    with pytest.raises(tw.TwoaceError):
        tw.phaselift_batch([A], [y], tw.PlOpts.default(lam=0.0), gpu_ctx)
    with pytest.raises(ValueError):
        tw.phaselift_batch([A], [y[:-1]], None, gpu_ctx)
