"""GPU parity tests: the CUDA path (through the C ABI) against the NumPy oracle on identical seeded
inputs and identical index sets.

Tolerances (floating point, FP64 on both sides):
  * per-stage iterate state after 1/2/10/100 iterations: relative Frobenius error <= 1e-9
    (measured ~1e-14; the bar is loose only to stay robust across cuBLAS-free summation orders);
  * full solves: per-instance relative error of the recovered CSI after global-phase alignment
    (Evaluation_H.m:81-82) <= 1e-4 for >= 95 % of instances, quality to 1e-6,
    integer bookkeeping (rank-one flag, roll-back flag, best trial, iteration counts, best
    iteration / column) bit-exact on the instances that meet the bar;
  * NMSE aggregated as 10*log10(mean) within 0.05 dB.

Determinism notes (see DESIGN.md "Parity regimes") — two places where the REFERENCE's own result is
decided by rounding noise, so that no independent implementation (MATLAB on another BLAS included) can
reproduce it per instance:
  (1) tolerances forced to 0 on an under-determined instance (M=64 < 256 unknowns per column): the
      objective reaches ~1e-16 after ~150 iterations and the best-iterate / best-column / mu decisions
      (:325, :333-334, :358) then compare rounding noise;
  (2) the column orthonormalisation [Vx,~] = eig(X'*X); X = X*Vx (:242-243) when X'*X is numerically
      rank deficient (rank-one re-runs, nuclear variant): columns X*v for null-space v are pure
      cancellation noise, which the per-column rescale of :282-284 then blows up to O(1) start points
      of the parallel refinement.
Measured on B200 (tools/gpu_diag3.py, gpurun_out/diag3_*.log): stage A keeps a 1e-14 input perturbation at
1e-14; the parallel-refinement stage of the ORACLE ITSELF turns it into 1e-12 ... 1e-2, and the GPU's
deviation from the oracle is never larger than that self-sensitivity.  Per-instance parity of full solves
is therefore asserted on the "reference-determined" instances — those where the oracle reproduces its own
CSI to 1e-6 when the RSS input is perturbed by 1e-14 relative — and everywhere else the GPU error is bounded
by 100x the oracle's self-sensitivity; iterate-state parity per stage (identical start point) and NMSE
statistics are asserted on everything.
"""
import numpy as np
import pytest

from oracle import admm

pytestmark = pytest.mark.gpu

TX = RX = 16


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


def _stage_case(codebook, M, trial=0):
    from twoace_b200 import harness as hz
    ins = hz.make_batch(trial + 1, codebook, M, 20.0)[trial]
    A, B, _, _ = admm._preprocess(ins.A, ins.B, 1e-8)
    tr = ins.train_idx[0]
    return A[tr], B[tr]


@pytest.mark.parametrize("M", [36, 64, 121, 225, 256])
def test_spectral_init_parity(codebook, gpu_ctx, M):
    from twoace_b200 import solvers as sv
    At, Bt = _stage_case(codebook, M)
    r = min(20, At.shape[0])
    Xo = admm.spectral_initialize(At, Bt, r)
    Xg = sv.spectral_init_batch([At], [Bt], r, gpu_ctx)[0]
    # the top-r eigenvectors are defined up to phase: compare the gauge-invariant X X'
    assert rel(Xg @ Xg.conj().T, Xo @ Xo.conj().T) < 1e-10


@pytest.mark.parametrize("M", [361, 1024])
@pytest.mark.parametrize("solver", ["tridiag", "jacobi"])
def test_spectral_init_n_by_n_branch(codebook, gpu_ctx, M, solver):
    """m_train > n = 256: eig of the n x n matrix As'As.  Both eigensolvers: the leading-eigenpair route
    (csrc/tridiag_eig.cuh, the default for 96 < d <= 512) and the full block-Jacobi decomposition."""
    from twoace_b200 import solvers as sv
    At, Bt = _stage_case(codebook, M)
    Xo = admm.spectral_initialize(At, Bt, 20)
    gpu_ctx.set_option("spectral_jacobi", 1 if solver == "jacobi" else 0)
    try:
        Xg = sv.spectral_init_batch([At], [Bt], 20, gpu_ctx)[0]
    finally:
        gpu_ctx.set_option("spectral_jacobi", 0)
    assert rel(Xg @ Xg.conj().T, Xo @ Xo.conj().T) < 1e-9


@pytest.fixture(params=["general", "fast_cs2", "fast_cs4", "simt_cs2", "simt_cs4"])
def kernel_path(request, gpu_ctx):
    """Run a test on the general kernel and on the shared-memory cluster kernel: cluster size 2 and 4, with the
    A-products on the tensor cores (tcgen05 int8, the default) and as FP64 SIMT products ("simt_*")."""
    gpu_ctx.set_option("fast", 0 if request.param == "general" else 1)
    gpu_ctx.set_option("fast_cs", 4 if request.param.endswith("cs4") else 2)
    gpu_ctx.set_option("tensor", 0 if request.param.startswith("simt") else 1)
    yield request.param
    gpu_ctx.set_option("fast", 1)
    gpu_ctx.set_option("fast_cs", 2)
    gpu_ctx.set_option("tensor", 1)


def _stage_check(gpu_ctx, At, Bt, X0, sbr, r1, nuc, iters):
    """One InferADMM call on the GPU against the oracle from the same start point: iterate state, outputs and the
    integer bookkeeping.  Returns the stage words."""
    import twoace_b200 as tw
    from twoace_b200 import solvers as sv
    snap = {iters: None}
    tro = admm.StageTrace()
    zfn = admm.argmin_z_nuclear if nuc else admm.argmin_z
    Xo, Yo, _ = admm.infer_admm(At, Bt, X0, sbr, r1, TX, RX, 0.0, 1e-3, 1.03, 0.0, 0.0, iters, None, None, zfn,
                                tro, snap)
    p = tw.Params.default(maxiter=iters, tol_rel=0.0, tol_abs=0.0)
    Xg, Yg, Sg, W = sv.infer_admm_batch([At], [Bt], [X0], sbr, r1, TX, RX, p, nuclear=nuc, ctx=gpu_ctx)
    s = snap[iters]
    tol = 1e-9
    if (nuc and iters >= 100) or (iters >= 60 and At.shape[0] > 256):
        # the nuclear iteration is expansive while tau = 1/mu still zeroes Z (x1.26 per iteration measured), and so is
        # the iteration from a raw spectral start at large m (column-wise projection, r = 1: 1e-15 -> 1e-8 ... 1e-3
        # after 60 iterations in the oracle itself, profiles/r02_big_check_M361_529_1024.log):
        # bound the deviation by the oracle's own response to a 1e-15 relative perturbation of X0
        snap2 = {iters: None}
        rng = np.random.default_rng(1)
        admm.infer_admm(At, Bt, X0 * (1 + 1e-15 * rng.standard_normal(X0.shape)), sbr, r1, TX, RX, 0.0, 1e-3, 1.03,
                        0.0, 0.0, iters, None, None, zfn, None, snap2)
        tol = max(tol, 100 * rel(snap2[iters]["X"], s["X"]))
    assert rel(Sg[0]["X"], s["X"]) < tol
    assert rel(Sg[0]["Z"], s["Z"]) < tol or np.linalg.norm(s["Z"]) < 1e-12
    assert rel(Sg[0]["Y"], s["Y"]) < tol
    # M and N can be pure rounding noise (exactly-satisfied constraints): compare absolutely
    assert np.linalg.norm(Sg[0]["M"] - s["M"]) < tol * max(1.0, np.linalg.norm(s["M"]))
    assert np.linalg.norm(Sg[0]["N"] - s["N"]) < tol * max(1.0, np.linalg.norm(s["N"]))
    assert rel(Xg[0], Xo) < tol and rel(Yg[0], Yo) < tol
    assert int(W[0][2]) == iters
    if tol == 1e-9:
        assert abs(W[0][0] - tro.mu) <= 1e-12 * tro.mu
        assert int(W[0][3]) == tro.opt_iter and int(W[0][4]) == tro.opt_col     # bit-exact bookkeeping
        assert int(W[0][5]) == tro.n_mu_bumps
    return W[0]


@pytest.mark.parametrize("sbr,r,r1,nuc", [(True, 20, False, False), (False, 20, False, False),
                                          (True, 1, True, False), (True, 20, True, False),
                                          (True, 20, False, True), (False, 20, False, True), (True, 1, False, True),
                                          (False, 7, False, False)])
@pytest.mark.parametrize("iters", [1, 2, 10, 100])
def test_stage_state_parity(codebook, gpu_ctx, kernel_path, sbr, r, r1, nuc, iters):
    At, Bt = _stage_case(codebook, 64)
    fast0, tc0 = gpu_ctx.fast_launch_count, gpu_ctx.tensor_launch_count
    X0 = admm.spectral_initialize(At, Bt, 20)[:, :r]
    _stage_check(gpu_ctx, At, Bt, X0, sbr, r1, nuc, iters)
    eligible = kernel_path != "general" and r in (1, 20)
    assert (gpu_ctx.fast_launch_count - fast0 == 1) == eligible      # the intended kernel really ran
    assert (gpu_ctx.tensor_launch_count - tc0 == 1) == (kernel_path.startswith("fast") and r == 20)


@pytest.mark.parametrize("M,snr", [(32, 0.0), (32, 30.0), (128, 0.0), (128, 30.0), (256, 0.0), (256, 30.0)])
@pytest.mark.parametrize("iters", [1, 2, 10, 100, 320])
@pytest.mark.parametrize("tensor", [1, 0])
def test_stage_state_parity_headline_grid(codebook, gpu_ctx, M, snr, iters, tensor):
    """The operating points of the headline benchmark (inferLowRank_Nuclear, M in {32,128,256}, SNR 0/30 dB) on the
    kernel each of them takes by default -- cluster size 2 (m = 30) and cluster size 4 (m = 121, 243), with the
    tensor-core and with the FP64 SIMT products -- up to an iteration count at which the singular-value
    threshold 1/mu no longer zeroes Z, so that the r x r eigensolve of inferLowRank_Nuclear.m:411-439 is live."""
    from twoace_b200 import harness as hz
    ins = hz.make_batch(1, codebook, M, snr)[0]
    A, B, _, _ = admm._preprocess(ins.A, ins.B, 1e-8)
    tr = ins.train_idx[0]
    At, Bt = A[tr], B[tr]
    X0 = admm.spectral_initialize(At, Bt, 20)
    gpu_ctx.set_option("tensor", tensor)
    try:
        fast0, tc0 = gpu_ctx.fast_launch_count, gpu_ctx.tensor_launch_count
        W = _stage_check(gpu_ctx, At, Bt, X0, True, False, True, iters)
        assert gpu_ctx.fast_launch_count - fast0 == 1
        assert gpu_ctx.tensor_launch_count - tc0 == tensor
        if iters >= 320:
            assert W[8] > 0      # Jacobi sweeps were executed: the SVT was live
    finally:
        gpu_ctx.set_option("tensor", 1)


@pytest.mark.parametrize("M", [36, 121, 225, 361])
def test_stage_parity_other_shapes(codebook, gpu_ctx, kernel_path, M):
    """M=36: K-split small-m path; M=361: explicit-inverse (non-Woodbury) branch."""
    import twoace_b200 as tw
    from twoace_b200 import solvers as sv
    At, Bt = _stage_case(codebook, M)
    X0 = admm.spectral_initialize(At, Bt, 20)
    for iters in (1, 25):
        snap = {iters: None}
        admm.infer_admm(At, Bt, X0, True, False, TX, RX, 0.0, 1e-3, 1.03, 0.0, 0.0, iters, None, None,
                        admm.argmin_z, None, snap)
        p = tw.Params.default(maxiter=iters, tol_rel=0.0, tol_abs=0.0)
        _, _, Sg, _ = sv.infer_admm_batch([At], [Bt], [X0], True, False, TX, RX, p, ctx=gpu_ctx)
        assert rel(Sg[0]["X"], snap[iters]["X"]) < 1e-9
        assert rel(Sg[0]["Y"], snap[iters]["Y"]) < 1e-9


@pytest.mark.parametrize("M", [361, 529, 1024])
@pytest.mark.parametrize("sbr,r,r1", [(True, 20, False), (False, 20, False), (True, 20, True), (True, 1, False),
                                      (True, 1, True), (False, 1, False)])
@pytest.mark.parametrize("iters", [1, 10, 60])
def test_stage_state_parity_large_m(codebook, gpu_ctx, M, sbr, r, r1, iters):
    """256 < m <= 1024 rows -- the upper half of the reference's M sweep (A2only.m:106-118), where the solver recovers
    the channel -- on the chunked cluster kernel (r = 20: big_stage_kernel) and the r = 1 refinement kernel
    (big1_stage_kernel), against the oracle from the same start point."""
    At, Bt = _stage_case(codebook, M)
    assert At.shape[0] > 256
    fast0 = gpu_ctx.fast_launch_count
    X0 = admm.spectral_initialize(At, Bt, 20)[:, :r]
    _stage_check(gpu_ctx, At, Bt, X0, sbr, r1, False, iters)
    assert gpu_ctx.fast_launch_count - fast0 == 1      # not the general kernel


def test_stage_large_m_convergence_mode(codebook, gpu_ctx):
    """Default tolerances at m > 256: iteration count, convergence flag and result against the oracle (r = 20 and 1)."""
    import twoace_b200 as tw
    from twoace_b200 import solvers as sv
    At, Bt = _stage_case(codebook, 529)
    for r in (20, 1):
        X0 = admm.spectral_initialize(At, Bt, 20)[:, :r]
        tro = admm.StageTrace()
        Xo, Yo, conv = admm.infer_admm(At, Bt, X0, True, False, TX, RX, 0.0, 1e-3, 1.03, 1e-4, 1e-8, 500, None, None,
                                       admm.argmin_z, tro)
        fast0 = gpu_ctx.fast_launch_count
        Xg, Yg, _, W = sv.infer_admm_batch([At], [Bt], [X0], True, False, TX, RX, tw.Params.default(), ctx=gpu_ctx)
        assert gpu_ctx.fast_launch_count - fast0 == 1
        assert int(W[0][2]) == tro.iters and bool(W[0][6]) == conv
        assert rel(Xg[0], Xo) < 1e-9 and rel(Yg[0], Yo) < 1e-9


def test_convergence_test_mode_matches_iteration_count(codebook, gpu_ctx, kernel_path):
    import twoace_b200 as tw
    from twoace_b200 import solvers as sv
    At, Bt = _stage_case(codebook, 64)
    X0 = admm.spectral_initialize(At, Bt, 20)
    tro = admm.StageTrace()
    Xo, Yo, conv = admm.infer_admm(At, Bt, X0, True, False, TX, RX, 0.0, 1e-3, 1.03, 1e-4, 1e-8, 500, None, None,
                                   admm.argmin_z, tro)
    Xg, Yg, _, W = sv.infer_admm_batch([At], [Bt], [X0], True, False, TX, RX, tw.Params.default(), ctx=gpu_ctx)
    assert int(W[0][2]) == tro.iters and bool(W[0][6]) == conv
    assert rel(Xg[0], Xo) < 1e-9


def _solve_both(variant, insts, p_gpu, p_or, ctx):
    """GPU batch solve + oracle solve + a second oracle solve on inputs perturbed at rounding level
    (B * (1 + 1e-14 N(0,1))): the reference's own sensitivity, instance by instance."""
    import twoace_b200 as tw
    from twoace_b200 import harness as hz, solvers as sv
    T = 3 if variant == tw.V4_MULTI else 1
    res = sv.solve_batch(variant, [i.A for i in insts], [i.B for i in insts], TX, RX,
                         [i.train_idx[:T] for i in insts], p_gpu, ctx)
    fn = {tw.V4: admm.infer_low_rank_v4, tw.V4_MULTI: admm.infer_low_rank_v4_multi,
          tw.NUCLEAR: admm.infer_low_rank_nuclear}[variant]
    out = []
    rng = np.random.default_rng(12345)
    for ins in insts:
        info = admm.SolveInfo()
        tri = ins.train_idx[:3] if variant == tw.V4_MULTI else ins.train_idx[0]
        Xo, Yo, qo = fn(ins.A, ins.B, TX, RX, p_or, train_idx=tri, info=info)
        Xp, _, _ = fn(ins.A, ins.B * (1 + 1e-14 * rng.standard_normal(ins.B.shape)), TX, RX, p_or, train_idx=tri)
        out.append((Xo, Yo, qo, info, hz.aligned_rel_err(Xp, Xo)))
    return res, out


def _check_full(res, out, insts, frac=0.95, min_determined=1):
    """BASELINE.md §4 bar on the instances the reference itself determines: an instance counts when the
    oracle reproduces its own CSI to 1e-6 under a 1e-14 relative perturbation of the RSS input."""
    from twoace_b200 import harness as hz
    errs = np.array([hz.aligned_rel_err(res.X[b], out[b][0]) for b in range(len(insts))])
    selfs = np.array([out[b][4] for b in range(len(insts))])
    det = selfs <= 1e-6
    print(f"\n  reference-determined instances: {det.sum()}/{len(insts)}; gpu-vs-oracle {errs}; oracle self-sensitivity {selfs}")
    assert det.sum() >= min_determined, f"test set has only {det.sum()} reference-determined instances"
    ok = errs <= 1e-4
    if det.any():
        assert ok[det].mean() >= frac, f"only {ok[det].mean():.2%} of the determined instances within 1e-4: {errs} {selfs}"
    # where the reference is noise-decided (self-sensitivity > 1e-6) both numbers are single samples of a chaotically
    # amplified quantity: the GPU's deviation from the oracle is held to 100x the oracle's own response to a 1e-14
    # perturbation of its input; the instances still enter the NMSE statistics below
    nd = ~det
    assert np.all(errs[nd] <= np.maximum(1e-4, 100.0 * selfs[nd])), f"noise-decided instances beyond 100x the oracle's self-sensitivity: {errs} {selfs}"
    assert np.all(np.isfinite(res.X) | np.isnan(res.X).all(axis=1, keepdims=True))
    for b in np.nonzero(ok & det)[0]:
        Xo, Yo, qo, info, _ = out[b]
        assert abs(res.quality[b] - qo) < 1e-6 or (np.isnan(qo) and np.isnan(res.quality[b]))
        assert int(res.info[b, 2]) == int(info.used_rank_one)
        assert int(res.info[b, 3]) == int(info.rolled_back)
        assert int(res.info[b, 4]) == info.best_trial
        assert res.Y[b].shape == Yo.shape
    idx = np.nonzero(det)[0]
    if len(idx):
        nm_g = hz.nmse_db([hz.nmse(res.X[b], insts[b].vecH) for b in idx])
        nm_o = hz.nmse_db([hz.nmse(out[b][0], insts[b].vecH) for b in idx])
        assert abs(nm_g - nm_o) <= 0.05, (nm_g, nm_o)
    nm_g = hz.nmse_db([hz.nmse(res.X[b], insts[b].vecH) for b in range(len(insts))])
    nm_o = hz.nmse_db([hz.nmse(out[b][0], insts[b].vecH) for b in range(len(insts))])
    assert abs(nm_g - nm_o) <= 1.0, (nm_g, nm_o)
    return errs


@pytest.mark.parametrize("variant_name,M,snr", [("V4", 64, 20.0), ("V4_MULTI", 64, 20.0), ("V4", 225, 20.0),
                                                 ("NUCLEAR", 64, 20.0)])
def test_full_solve_parity_default_tolerances(codebook, gpu_ctx, kernel_path, variant_name, M, snr):
    """The reference's own operating mode (no caller passes more than 4 arguments)."""
    import twoace_b200 as tw
    from twoace_b200 import harness as hz
    variant = getattr(tw, variant_name)
    n_inst = 12 if M == 64 else 8
    insts = hz.make_batch(n_inst, codebook, M, snr)
    res, out = _solve_both(variant, insts, tw.Params.default(), admm.Params(), gpu_ctx)
    _check_full(res, out, insts, min_determined=0 if variant_name == "NUCLEAR" else 1)


@pytest.mark.parametrize("variant_name,M,n_inst", [("V4", 361, 6), ("V4", 529, 6), ("V4_MULTI", 529, 4), ("V4", 1024, 3)])
def test_full_solve_parity_where_recovery_works(codebook, gpu_ctx, variant_name, M, n_inst):
    """M >= 361 at 20 dB is the regime in which the reference recovers the channel (NMSE about -15 dB at M=529), so
    the parity statistic discriminates here: per-instance error on the reference-determined instances, NMSE of
    both sides against the true channel, bit-exact flags.  Runs on the large-m cluster kernels."""
    import twoace_b200 as tw
    from twoace_b200 import harness as hz
    variant = getattr(tw, variant_name)
    insts = hz.make_batch(n_inst, codebook, M, 20.0)
    fast0 = gpu_ctx.fast_launch_count
    res, out = _solve_both(variant, insts, tw.Params.default(), admm.Params(), gpu_ctx)
    assert gpu_ctx.fast_launch_count > fast0
    # measured (B200, r02): even here the oracle's own CSI moves by 1e-5 ... 2e-1 under a 1e-14 perturbation of the RSS
    # input (the refinement stages do not converge within 500 iterations), so no instance is reference-determined;
    # the bar is the 100x-self-sensitivity bound per instance plus the NMSE of both sides against the true channel
    errs = _check_full(res, out, insts, min_determined=0)
    nm_g = hz.nmse_db([hz.nmse(res.X[b], insts[b].vecH) for b in range(n_inst)])
    nm_o = hz.nmse_db([hz.nmse(out[b][0], insts[b].vecH) for b in range(n_inst)])
    print(f"  M={M} {variant_name}: NMSE gpu {nm_g:.3f} dB, oracle {nm_o:.3f} dB, max err {errs.max():.2e}")
    if M >= 529:
        assert nm_o < -8.0 and nm_g < -8.0       # the channel really is recovered on both sides


def test_known_answer_nuclear_and_fixed_iterations(gpu_ctx):
    """Gaussian A / exact rank-2 channel (the authors' recipe): well conditioned for every variant;
    checks the nuclear variant and the forced-iteration mode end to end against the oracle."""
    import twoace_b200 as tw
    from twoace_b200 import harness as hz, solvers as sv
    rng = np.random.default_rng(11)
    T = 256
    A = (rng.standard_normal((T, 64)) + 1j * rng.standard_normal((T, 64))) / 8
    Z = (rng.standard_normal((8, 2)) + 1j * rng.standard_normal((8, 2))) @ \
        (rng.standard_normal((2, 8)) + 1j * rng.standard_normal((2, 8)))
    Xgt = Z.reshape(-1, order="F")
    B = np.abs(A @ Xgt) * (1 + 0.01 * rng.standard_normal(T))
    tr = rng.permutation(T)[:243].astype(np.int32)
    for variant, fn, p, po in [
            (tw.NUCLEAR, admm.infer_low_rank_nuclear, tw.Params.default(maxiter=150), admm.Params(maxiter=150)),
            (tw.V4, admm.infer_low_rank_v4, tw.Params.default(maxiter=80).fixed_iters(),
             admm.Params(maxiter=80).fixed_iters())]:
        res = sv.solve_batch(variant, [A], [B], 8, 8, [tr[None]], p, gpu_ctx)
        info = admm.SolveInfo()
        Xo, Yo, qo = fn(A, B, 8, 8, po, train_idx=tr, info=info)
        Xp, _, _ = fn(A, B * (1 + 1e-14 * rng.standard_normal(T)), 8, 8, po, train_idx=tr)
        e_self = hz.aligned_rel_err(Xp, Xo)
        assert hz.aligned_rel_err(res.X[0], Xo) <= max(1e-6, 100 * e_self)
        if e_self < 1e-8:
            assert abs(res.quality[0] - qo) < 1e-6
            assert [int(w[2]) for w in res.stage_words[0] if w[2] > 0] == [t.iters for t in info.traces]


def test_nmse_statistics_fixed_iterations_config1(codebook, gpu_ctx):
    """BASELINE config 1 (M=64, SNR 20 dB, 500 forced iterations): per-instance CSI is noise-determined
    (see module docstring) but the reported NMSE must agree within 0.05 dB."""
    import twoace_b200 as tw
    from twoace_b200 import harness as hz
    insts = hz.make_batch(16, codebook, 64, 20.0)
    res, out = _solve_both(tw.V4, insts, tw.Params.default().fixed_iters(), admm.Params().fixed_iters(), gpu_ctx)
    selfs = np.array([o[4] for o in out])
    assert np.median(selfs) > 1e-3     # documents (1): the reference does not reproduce itself here
    nm_g = hz.nmse_db([hz.nmse(res.X[b], insts[b].vecH) for b in range(len(insts))])
    nm_o = hz.nmse_db([hz.nmse(out[b][0], insts[b].vecH) for b in range(len(insts))])
    assert abs(nm_g - nm_o) <= 0.05, (nm_g, nm_o)
    # every stage ran exactly maxiter iterations
    for b in range(len(insts)):
        assert all(int(w[2]) in (0, 500) for w in res.stage_words[b])


def test_known_answer_recipe_on_gpu(gpu_ctx):
    """ADMM_v2.m:13-19,47-48 through the MATLAB-signature mirror."""
    import twoace_b200 as tw
    rng = np.random.default_rng(5)
    T = 256
    A = rng.standard_normal((T, 64)) + 1j * rng.standard_normal((T, 64))
    Z = (rng.standard_normal((8, 2)) + 1j * rng.standard_normal((8, 2))) @ \
        (rng.standard_normal((2, 8)) + 1j * rng.standard_normal((2, 8)))
    Xgt = Z.reshape(-1, order="F")
    B = np.abs(A @ Xgt)
    tr = rng.permutation(T)[:243]
    X, Y, q = tw.inferLowRankV4(A, B, 8, 8, train_idx=tr[None], ctx=gpu_ctx)
    ratio = X / Xgt
    assert np.linalg.norm(ratio - ratio.mean()) / np.linalg.norm(ratio) < 1e-4
    assert q > 0.99
    Xo, Yo, qo = admm.infer_low_rank_v4(A, B, 8, 8, train_idx=tr)
    from twoace_b200 import harness as hz
    assert hz.aligned_rel_err(X, Xo) < 1e-6
    X3, _, q3 = tw.ADMM_v2(B, A, 8, 8, 3, tree="ns", train_idx=tr[None], ctx=gpu_ctx)
    np.testing.assert_array_equal(X3, X)


def test_codebook_mode_equals_dense_mode(codebook, gpu_ctx):
    import twoace_b200 as tw
    from twoace_b200 import harness as hz, solvers as sv
    insts = hz.make_batch(4, codebook, 64, 20.0)
    p = tw.Params.default()
    d = sv.solve_batch(tw.V4, [i.A for i in insts], [i.B for i in insts], TX, RX, [i.train_idx[:1] for i in insts],
                       p, gpu_ctx)
    gpu_ctx.set_codebook(codebook)
    c = sv.solve_batch_codebook(tw.V4, [i.rows for i in insts], 1.0 / 16, [i.B for i in insts], TX, RX,
                                [i.train_idx[:1] for i in insts], p, gpu_ctx)
    for b in range(4):
        assert hz.aligned_rel_err(c.X[b], d.X[b]) < 1e-8
        assert abs(c.quality[b] - d.quality[b]) < 1e-9
    np.testing.assert_array_equal(c.info[:, 2:7], d.info[:, 2:7])


def test_ragged_batch_equals_separate_solves(codebook, gpu_ctx):
    import twoace_b200 as tw
    from twoace_b200 import harness as hz, solvers as sv
    Ms = [36, 121, 64, 9]
    insts = [hz.make_batch(1, codebook, M, 20.0, base_seed=100 + M)[0] for M in Ms]
    p = tw.Params.default(maxiter=80)
    for fast in (0, 1):     # every instance takes the same kernel whatever its batch mates are
        gpu_ctx.set_option("fast", fast)
        both = sv.solve_batch(tw.V4, [i.A for i in insts], [i.B for i in insts], TX, RX,
                              [i.train_idx[:1] for i in insts], p, gpu_ctx)
        for b, ins in enumerate(insts):
            one = sv.solve_batch(tw.V4, [ins.A], [ins.B], TX, RX, [ins.train_idx[:1]], p, gpu_ctx)
            np.testing.assert_array_equal(one.X[0], both.X[b])   # same kernels, same order: bit-exact
            np.testing.assert_array_equal(one.info[0], both.info[b])
    gpu_ctx.set_option("fast", 1)


def test_degenerate_small_m_does_not_crash(codebook, gpu_ctx):
    """M=4 (m_train=3 < r=4): zero spectral column -> Inf/NaN per-column rescale; NaN columns
    never win (MATLAB min skips NaN) and the library must return without error."""
    import twoace_b200 as tw
    from twoace_b200 import harness as hz, solvers as sv
    ins = hz.make_batch(1, codebook, 4, 20.0)[0]
    res = sv.solve_batch(tw.V4_MULTI, [ins.A], [ins.B], TX, RX, [ins.train_idx], tw.Params.default(maxiter=50), gpu_ctx)
    assert res.X.shape == (1, 256)
    assert np.all(np.isfinite(res.X)) or np.all(np.isnan(res.X[0]))


def test_error_reporting(codebook, gpu_ctx):
    import twoace_b200 as tw
    from twoace_b200 import harness as hz, solvers as sv
    ins = hz.make_batch(1, codebook, 36, 20.0)[0]
    bad = ins.train_idx[:1].copy()
    bad[0, 0] = bad[0, 1]                                   # duplicate index
    with pytest.raises(tw.TwoaceError, match="unique"):
        sv.solve_batch(tw.V4, [ins.A], [ins.B], TX, RX, [bad], None, gpu_ctx)
    with pytest.raises(tw.TwoaceError, match="lambda"):
        sv.solve_batch(tw.V4, [ins.A], [ins.B], TX, RX, [ins.train_idx[:1]], tw.Params.default(lam=0.1), gpu_ctx)
    # the context stays usable after an error
    r = sv.solve_batch(tw.V4, [ins.A], [ins.B], TX, RX, [ins.train_idx[:1]], tw.Params.default(maxiter=5), gpu_ctx)
    assert r.X.shape == (1, 256)


def test_general_complex_A_takes_general_kernel(gpu_ctx):
    """A non-quantised sensing matrix (e.g. the 8-phase 'Random_Phase_State' alphabet or beams*AD,
    SURVEY.md §7.2) must not be forced onto the 2-bit path."""
    import twoace_b200 as tw
    from twoace_b200 import solvers as sv
    rng = np.random.default_rng(3)
    A = np.exp(1j * np.pi / 8 * rng.integers(0, 16, (60, 256))) / 16
    B = np.abs(A @ (rng.standard_normal(256) + 1j * rng.standard_normal(256)))
    B /= np.linalg.norm(B)
    X0 = (rng.standard_normal((256, 20)) + 1j * rng.standard_normal((256, 20))) / 16
    f0 = gpu_ctx.fast_launch_count
    p = tw.Params.default(maxiter=20, tol_rel=0.0, tol_abs=0.0)
    Xg, Yg, Sg, W = sv.infer_admm_batch([A], [B], [X0], True, False, TX, RX, p, ctx=gpu_ctx)
    assert gpu_ctx.fast_launch_count == f0
    snap = {20: None}
    admm.infer_admm(A, B, X0, True, False, TX, RX, 0.0, 1e-3, 1.03, 0.0, 0.0, 20, None, None, admm.argmin_z, None, snap)
    assert rel(Sg[0]["X"], snap[20]["X"]) < 1e-9


def test_concurrent_launch_groups_do_not_change_results(codebook, gpu_ctx):
    """A batch that mixes every kernel group of a stage launch (M = 4 general kernel, 36 and 121 cluster kernels of
    both shapes, 361 large-M kernels): with the groups launched concurrently on side streams and tasks handed out by
    the device queue (default) the results are bitwise those of the serial launch order (option overlap = 0)."""
    import twoace_b200 as tw
    from twoace_b200 import harness as hz, solvers as sv
    insts = []
    for M in (4, 36, 121, 361, 36, 4, 121):
        insts += hz.make_batch(2, codebook, M, 20.0)
    p = tw.Params.default(maxiter=40).fixed_iters()
    args = ([i.A for i in insts], [i.B for i in insts], TX, RX, [i.train_idx[:3] for i in insts], p, gpu_ctx)
    a = sv.solve_batch(tw.V4_MULTI, *args)
    gpu_ctx.set_option("overlap", 0)
    try:
        b = sv.solve_batch(tw.V4_MULTI, *args)
    finally:
        gpu_ctx.set_option("overlap", 1)
    assert np.array_equal(a.X, b.X, equal_nan=True) and np.array_equal(a.quality, b.quality, equal_nan=True)
    assert np.array_equal(a.info, b.info, equal_nan=True)


def test_dense_batch_kernel_choice_is_per_instance(codebook, gpu_ctx):
    """Dense mode: a quantised and a non-quantised sensing matrix in ONE batch.  The quantised instance still takes the
    cluster kernel (the 2-bit decision is per instance), and both results are bitwise those of separate calls."""
    import twoace_b200 as tw
    from twoace_b200 import harness as hz, solvers as sv
    rng = np.random.default_rng(5)
    ins = hz.make_batch(1, codebook, 64, 20.0)[0]
    A8 = np.exp(1j * np.pi / 8 * rng.integers(0, 16, (64, 256))) / 16
    B8 = np.abs(A8 @ ins.vecH)
    tr = ins.train_idx[:1]
    p = tw.Params.default(maxiter=40).fixed_iters()
    f0 = gpu_ctx.fast_launch_count
    both = sv.solve_batch(tw.V4, [ins.A, A8], [ins.B, B8], TX, RX, [tr, tr], p, gpu_ctx)
    assert gpu_ctx.fast_launch_count > f0                      # the quantised instance ran on the cluster kernel
    one = sv.solve_batch(tw.V4, [ins.A], [ins.B], TX, RX, [tr], p, gpu_ctx)
    two = sv.solve_batch(tw.V4, [A8], [B8], TX, RX, [tr], p, gpu_ctx)
    assert np.array_equal(both.X[0], one.X[0]) and np.array_equal(both.X[1], two.X[0])
    assert both.quality[0] == one.quality[0] and both.quality[1] == two.quality[0]


def _synthetic_case(tx, rx, m, seed, L=3):
    """2-bit random beams (Generate_random_beam.m:31-34) on a tx x rx array, sparse multipath channel."""
    from twoace_b200 import harness as hz
    rng = np.random.default_rng(seed)
    n = tx * rx
    _, vecH, _, _ = hz.generate_channel(rng, tx, rx, L)
    A = np.exp(1j * (np.pi / 2) * rng.integers(0, 4, (m, n))) / np.sqrt(n)
    B = np.abs(A @ vecH)
    return A, B, vecH


@pytest.mark.parametrize("tx,rx,m", [(4, 4, 24), (8, 8, 40), (8, 4, 30), (32, 32, 96)])
@pytest.mark.parametrize("nuc", [False, True])
def test_stage_parity_other_antenna_counts(gpu_ctx, tx, rx, m, nuc):
    """Rank-shaping profiles other than the 16-antenna one (inferLowRankV4.m:416-443: one stage for 4 and 8
    antennas, four stages [3 4 6 12] for 32) and n = 1024 (BASELINE config 5 shape) on the general kernel."""
    import twoace_b200 as tw
    from twoace_b200 import solvers as sv
    A, B, _ = _synthetic_case(tx, rx, m, 17 + tx)
    A, B, _, _ = admm._preprocess(A, B, 1e-8)
    n = tx * rx
    r = min(20, m, n)
    X0 = admm.spectral_initialize(A, B, r)
    zfn = admm.argmin_z_nuclear if nuc else admm.argmin_z
    for iters in (1, 8):
        snap = {iters: None}
        tro = admm.StageTrace()
        admm.infer_admm(A, B, X0, True, False, tx, rx, 0.0, 1e-3, 1.03, 0.0, 0.0, iters, None, None, zfn, tro, snap)
        p = tw.Params.default(maxiter=iters, tol_rel=0.0, tol_abs=0.0)
        _, _, Sg, W = sv.infer_admm_batch([A], [B], [X0], True, False, tx, rx, p, nuclear=nuc, ctx=gpu_ctx)
        assert rel(Sg[0]["X"], snap[iters]["X"]) < 1e-9
        assert rel(Sg[0]["Y"], snap[iters]["Y"]) < 1e-9
        assert rel(Sg[0]["Z"], snap[iters]["Z"]) < 1e-9 or np.linalg.norm(snap[iters]["Z"]) < 1e-12
        assert int(W[0][3]) == tro.opt_iter


@pytest.mark.parametrize("tx,rx,m,r", [(32, 32, 200, 20), (32, 32, 200, 1), (8, 32, 70, 20), (4, 4, 24, 16)])
def test_general_kernel_codes_match_dense_products(gpu_ctx, tx, rx, m, r):
    """The general kernel keeps the 2-bit codes of a quantised A in shared memory (products without global reads of
    A, I + A A' by popcount arithmetic); with the option "fast" off it runs the dense products on the same instance.
    Same algorithm, different rounding: the states agree to 1e-11 after 6 iterations, the bookkeeping exactly."""
    import twoace_b200 as tw
    from twoace_b200 import solvers as sv
    A, B, _ = _synthetic_case(tx, rx, m, 40 + tx + r)
    A, B, _, _ = admm._preprocess(A, B, 1e-8)
    n = tx * rx
    X0 = admm.spectral_initialize(A, B, r)
    p = tw.Params.default(maxiter=6, tol_rel=0.0, tol_abs=0.0)
    _, _, Sc, Wc = sv.infer_admm_batch([A], [B], [X0], r > 1, False, tx, rx, p, ctx=gpu_ctx)
    gpu_ctx.set_option("fast", 0)
    try:
        _, _, Sd, Wd = sv.infer_admm_batch([A], [B], [X0], r > 1, False, tx, rx, p, ctx=gpu_ctx)
    finally:
        gpu_ctx.set_option("fast", 1)
    for k in ("X", "Y", "Z"):
        assert rel(Sc[0][k], Sd[0][k]) < 1e-11 or np.linalg.norm(Sd[0][k]) < 1e-12
    assert np.array_equal(Wc[0][2:7], Wd[0][2:7])       # iterations, best iteration / column, mu bumps, converged


@pytest.mark.parametrize("tx,rx,m", [(8, 8, 48), (4, 4, 40)])
def test_full_solve_other_antenna_counts(gpu_ctx, tx, rx, m):
    """inferLowRankV4 end to end (default tolerances) away from 16 x 16: same flags, quality and CSI."""
    import twoace_b200 as tw
    from twoace_b200 import harness as hz
    from twoace_b200 import solvers as sv
    A, B, vecH = _synthetic_case(tx, rx, m, 5 + tx)
    rng = np.random.default_rng(2)
    tr = sv.draw_train_idx(m, 0.95, 1, rng)
    info = admm.SolveInfo()
    Xo, Yo, qo = admm.infer_low_rank_v4(A, B, tx, rx, admm.Params(), train_idx=tr[0], info=info)
    res = sv.solve_batch(tw.V4, [A], [B], tx, rx, [tr], None, gpu_ctx)
    rng2 = np.random.default_rng(9)
    Xs, _, _ = admm.infer_low_rank_v4(A, B * (1 + 1e-14 * rng2.standard_normal(m)), tx, rx, admm.Params(),
                                      train_idx=tr[0])
    self_sens = hz.aligned_rel_err(Xs, Xo)
    assert hz.aligned_rel_err(res.X[0], Xo) < max(1e-6, 100 * self_sens)
    if self_sens < 1e-6:
        assert abs(res.quality[0] - qo) < 1e-6
        assert int(res.info[0][2]) == int(info.used_rank_one) and int(res.info[0][3]) == int(info.rolled_back)


@pytest.mark.parametrize("tx,rx", [(16, 16), (8, 4), (32, 32)])
def test_evaluation_metrics_parity(gpu_ctx, tx, rx):
    """twoace_metrics_batch against oracle/metrics.py (Evaluation_H.m:81-115): MSE_H, gain_ana, gain_dig,
    proj_error to 1e-9 relative; NaN estimates give NaN."""
    from oracle import metrics as om
    from twoace_b200 import harness as hz
    from twoace_b200 import solvers as sv
    rng = np.random.default_rng(40 + tx)
    n = tx * rx
    Xt, Xe = [], []
    for k in range(12):
        _, vecH, _, _ = hz.generate_channel(rng, tx, rx, 3)
        noise = (rng.standard_normal(n) + 1j * rng.standard_normal(n)) * (0.02 * (k + 1))
        Xt.append(vecH)
        Xe.append((vecH + noise) * np.exp(1j * rng.uniform(0, 6.28)) * rng.uniform(0.5, 2.0))
    Xe[5] = Xe[5].copy()
    Xe[5][0] = np.nan
    Xe[7] = np.zeros(n, complex)
    got = sv.evaluation_batch(np.array(Xe), np.array(Xt), tx, rx, 2, gpu_ctx)
    for k in range(12):
        ref = om.evaluation_h(Xe[k], Xt[k], tx, rx, 2)
        if k in (5, 7):
            assert np.all(np.isnan(got[k])) and all(np.isnan(v) for v in ref)
            continue
        for j in range(4):
            assert abs(got[k, j] - ref[j]) <= 1e-9 * max(abs(ref[j]), 1e-3), (k, j, got[k], ref)


@pytest.mark.parametrize("version", [1, 2, 3])
@pytest.mark.parametrize("case", ["cb16_M64", "cb16_M36", "syn8x8", "syn4x4_bad"])
def test_older_solver_versions_parity(codebook, gpu_ctx, version, case):
    """inferLowRank.m / inferLowRankV2.m / inferLowRankV3.m (main ADMM_v2.m versions 1-3) through ADMM_v2:
    V1 runs the general kernel (its single-stage rank profile), V2 / V3 the shared-memory kernel on 16 x 16.
    Same bar as the V4 full solves: reference-determined instances within 1e-6, quality and Y length equal."""
    import twoace_b200 as tw
    from twoace_b200 import harness as hz
    from twoace_b200 import solvers as sv
    rng = np.random.default_rng(77)
    if case.startswith("cb16"):
        M = int(case.split("M")[1])
        ins = hz.make_batch(2, codebook, M, 20.0)[1]
        A, B, tx, rx = ins.A, ins.B, 16, 16
    elif case == "syn8x8":
        A, B, _ = _synthetic_case(8, 8, 48, 3)
        tx = rx = 8
    else:
        A, B, _ = _synthetic_case(4, 4, 40, 4)
        tx = rx = 4
    m = A.shape[0]
    tr = sv.draw_train_idx(m, 0.95, 1, rng)
    if case == "syn4x4_bad":
        # consistent training rows, held-out rows off by 10x: quality ~ 0.1 <= 0.6 whatever the solver finds
        # (V1 / V2 then skip the refine, V3 refines without the roll-back test)
        test_rows = np.setdiff1d(np.arange(m), tr[0])
        B = B.copy()
        B[test_rows] *= 10.0
    fn = {1: admm.infer_low_rank_v1, 2: admm.infer_low_rank_v2, 3: admm.infer_low_rank_v3}[version]
    info = admm.SolveInfo()
    Xo, Yo, qo = fn(A, B, tx, rx, admm.Params(), train_idx=tr[0], info=info)
    rng2 = np.random.default_rng(9)
    Xs, _, _ = fn(A, B * (1 + 1e-14 * rng2.standard_normal(m)), tx, rx, admm.Params(), train_idx=tr[0])
    self_sens = hz.aligned_rel_err(Xs, Xo)
    f0 = gpu_ctx.fast_launch_count
    Xg, Yg, qg = sv.ADMM_v2(B, A, tx, rx, version, train_idx=tr, ctx=gpu_ctx)
    used_fast = gpu_ctx.fast_launch_count > f0
    assert used_fast == (case.startswith("cb16") and version != 1)
    if case != "syn4x4_bad" or self_sens < 1e-6:
        # (inconsistent data is decided by rounding noise in the reference itself, heavy-tailed: the 'bad' case
        #  only checks the control flow below unless the oracle reproduces itself)
        assert hz.aligned_rel_err(Xg, Xo) < max(1e-6, 100 * self_sens)
    if self_sens < 1e-6:
        assert abs(qg - qo) < 1e-6
    if case == "syn4x4_bad":
        assert not qo > 0.6
        res = sv.solve_batch({1: tw.V1, 2: tw.V2, 3: tw.V3}[version], [A], [B], tx, rx, [tr], None, gpu_ctx)
        assert int(res.info[0][5]) == len(Yo) == (m if version == 3 else len(tr[0]))
        assert int(res.info[0][3]) == 0          # not a roll-back


def test_nuclear_rerun_dedup_is_bitwise_identical(codebook, gpu_ctx):
    """inferLowRank_Nuclear.m:69-70 reruns the train solve with use_rank_one = true, which its ArgMinZ never
    reads: the opt-in "dedup_nuclear_rerun" must return exactly what the literal flow returns."""
    import twoace_b200 as tw
    from twoace_b200 import harness as hz
    from twoace_b200 import solvers as sv
    insts = hz.make_batch(3, codebook, 32, 0.0) + hz.make_batch(2, codebook, 64, 10.0)
    p = tw.Params.default(maxiter=120).fixed_iters()
    args = ([i.A for i in insts], [i.B for i in insts], 16, 16, [i.train_idx[:1] for i in insts], p, gpu_ctx)
    lit = sv.solve_batch(tw.NUCLEAR, *args)
    gpu_ctx.set_option("dedup_nuclear_rerun", 1)
    try:
        ded = sv.solve_batch(tw.NUCLEAR, *args)
    finally:
        gpu_ctx.set_option("dedup_nuclear_rerun", 0)
    assert lit.info[:, 2].sum() > 0                      # the rerun really fired somewhere
    assert np.array_equal(lit.X, ded.X) and np.array_equal(lit.quality, ded.quality)
    for a, b in zip(lit.Y, ded.Y):
        assert np.array_equal(a, b)
    assert np.array_equal(lit.info[:, :15], ded.info[:, :15], equal_nan=True)
    assert ded.info[:, 15].sum() < lit.info[:, 15].sum()  # fewer iterations executed


def test_residual_trace_output(codebook, gpu_ctx):
    """The optional residual trace (twoace_set_trace): res_comb of every executed iteration of every stage, NaN
    elsewhere; its last finite entry of a stage is stage word 7 and the count of finite entries the iteration count.
    Checked against the oracle's per-iteration res_comb for the first stage."""
    import twoace_b200 as tw
    from twoace_b200 import harness as hz, solvers as sv
    insts = hz.make_batch(3, codebook, 64, 20.0)
    p = tw.Params.default()
    nstage, maxiter = 5, int(p.maxiter)
    buf = np.zeros(len(insts) * nstage * maxiter, dtype=np.float64)
    gpu_ctx.set_trace(buf)
    try:
        res = sv.solve_batch(tw.V4, [i.A for i in insts], [i.B for i in insts], TX, RX,
                             [i.train_idx[:1] for i in insts], p, gpu_ctx)
    finally:
        gpu_ctx.set_trace(None)
    tr = buf.reshape(len(insts), nstage, maxiter)
    for b, ins in enumerate(insts):
        for st in range(nstage):
            iters = int(res.stage_words[b][st][2])
            fin = np.isfinite(tr[b, st])
            assert fin.sum() == iters and fin[:iters].all()
            if iters:
                assert tr[b, st, iters - 1] == res.stage_words[b][st][7]
        # first stage (over-parameterised, training rows) against the oracle's own residual history
        A, B, _, _ = admm._preprocess(ins.A, ins.B, 1e-8)
        t = ins.train_idx[0]
        X0 = admm.spectral_initialize(A[t], B[t], 20)
        tro = admm.StageTrace()
        admm.infer_admm(A[t], B[t], X0, True, False, TX, RX, 0.0, 1e-3, 1.03, 1e-4, 1e-8, 500, None, None,
                        admm.argmin_z, tro)
        assert int(res.stage_words[b][0][2]) == tro.iters
        if hasattr(tro, "res_comb") and len(tro.res_comb):
            ref = np.asarray(tro.res_comb)
            assert np.allclose(tr[b, 0, :tro.iters], ref, rtol=1e-6, atol=1e-12)
