/* twoace.h — C ABI of libtwoace: B200-native batched solver for 2ACE's ADMM CSI-recovery hot path.
 *
 * The reference has no FFI for this path; its boundary is the MATLAB function signature
 *   [X,Y,quality] = inferLowRankV4(A,B,tx,rx,lambda,r,mu0,rho,cc_frac,tol_rel,tol_abs,maxiter)
 * (main/src/my_recovery_algorithms/ADMM_v2/inferLowRankV4.m:1-9; same for inferLowRankV4_multi.m:5
 * and inferLowRank_Nuclear.m:5), reached through ADMM_v2.m:22-45 / ADMM_v2_nuclear.m:32 and, from
 * Python, through the MATLAB engine (main/main.py:231,308,427).  A MEX file of the same base name
 * shadows the .m file; the MEX shim (mex/twoace_mex.cpp) and the ctypes binding
 * (2ace-mmwave-channel-estimation_b200/lib.py) both sit on the entry points below.
 *
 * Conventions
 *   - complex arrays are interleaved (re,im) doubles, matrices column-major (MATLAB layout);
 *   - batches are ragged in m: per-instance arrays are concatenated in instance order;
 *   - the MATLAB global RNG is externalised: the randsample() draws of inferLowRankV4.m:37
 *     (inferLowRankV4_multi.m:48: three draws) are the `train_idx` input, 0-based, drawn order,
 *     floor(m*cc_frac) entries per draw;
 *   - `mem` says where ALL data pointers of a call live: TWOACE_MEM_HOST or TWOACE_MEM_DEVICE
 *     (index arrays `m`, `train_idx`, `cb_rows` are always host pointers);
 *   - every function returns 0 on success or a negative TWOACE_E_* code; twoace_last_error()
 *     returns the message.  Numerical failures are not errors: NaN propagates as in the reference.
 */
#ifndef TWOACE_H
#define TWOACE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct twoace_ctx twoace_ctx;

enum {
  TWOACE_OK = 0,
  TWOACE_E_INVALID = -1,      /* bad argument */
  TWOACE_E_UNSUPPORTED = -2,  /* valid in the reference but outside this build (e.g. lambda != 0) */
  TWOACE_E_CUDA = -3,         /* CUDA runtime failure */
  TWOACE_E_NOMEM = -4
};

enum { TWOACE_MEM_HOST = 0, TWOACE_MEM_DEVICE = 1 };

/* Solver variant = which reference file is replaced. */
enum {
  TWOACE_V4 = 0,        /* inferLowRankV4.m        (NS ADMM_v2.m version 3)        */
  TWOACE_V4_MULTI = 1,  /* inferLowRankV4_multi.m  (main ADMM_v2.m version 4)       */
  TWOACE_NUCLEAR = 2,   /* inferLowRank_Nuclear.m  (main ADMM_v2_nuclear.m version 4) */
  TWOACE_V3 = 3,        /* inferLowRankV3.m        (main ADMM_v2.m version 3): V4 without the rank-one rerun */
  TWOACE_V2 = 4,        /* inferLowRankV2.m        (main ADMM_v2.m version 2): refine only if quality > 0.6  */
  TWOACE_V1 = 5,        /* inferLowRank.m          (main ADMM_v2.m version 1): single-stage rank profile     */
  TWOACE_MINL2 = 6      /* inferMinL2.m            (ADMM_v2.m version 0): no low-rank variable, X = pinv(A)(Y - M/mu).
                         * [X,Y,quality] = inferMinL2(A,B,lambda,r,tol_rel,tol_abs,maxiter): mu0, rho and cc_frac of
                         * twoace_params are ignored (1e-3, 1.03, 0.95 are hard-wired, inferMinL2.m:34,269-270), the
                         * train draw has ceil(0.95 m) entries (:34), tx / rx only give n = tx * rx */
};

/* Optional positional arguments of inferLowRankV4.m:2-9 (defaults in twoace_default_params). */
typedef struct {
  double lambda;   /* 0 (only value supported: the lambda != 0 eigen-path is dead in the reference) */
  int32_t r;       /* 20 */
  double mu0;      /* 1e-3 */
  double rho;      /* 1.03 */
  double cc_frac;  /* 0.95 */
  double tol_rel;  /* 1e-4 */
  double tol_abs;  /* 1e-8 */
  int32_t maxiter; /* 500 */
} twoace_params;

#define TWOACE_INFO_WORDS 16
/* info[b*16 + k]: 0 quality, 1 similarity (NaN if not computed), 2 use_rank_one (last trial),
 * 3 rolled_back, 4 best_trial, 5 rows of returned Y, 6 bitmask of trials that re-ran rank-one,
 * 7..9 per-trial quality, 10 max_quality, 11 refine iterations, 12 refine opt_iter,
 * 13 refine final mu, 14 refine mu bumps, 15 total ADMM iterations over all stages. */

#define TWOACE_STAGE_WORDS 16
/* stage bookkeeping words (per InferADMM call): 0 mu, 1 opt_obj, 2 iters, 3 opt_iter (1-based),
 * 4 opt_col (0-based, -1 for scale_by_row), 5 mu bumps, 6 converged, 7 last res_comb,
 * 8 Jacobi sweeps; 9..14 device clock cycles (fast kernel only): eigensolve, X update, whole loop,
 * Y/M update, ArgMinZ, exchange + best-iterate tail; 15 set-up cycles of the stage (codes, S^-1, operand blocks). */

void twoace_default_params(twoace_params* p);

/* Library / build identification (e.g. 100 = round 1). */
int twoace_version(void);

int twoace_create(int device, twoace_ctx** ctx);
/* One context over several GPUs of a node (twoace_create(device_ids, n_dev, &ctx)).  Instances
 * are independent, so every batched entry point (solve_batch, solve_batch_codebook, phaselift_batch, metrics_batch,
 * synth_batch) splits its batch into contiguous slices balanced by row count, one per GPU, each driven by its own
 * host thread; results land in the caller's buffers at the slice offsets and do not depend on the GPU count.
 * A multi-GPU context takes HOST buffers only; codebook and options are replicated on every GPU; the residual trace
 * is not available.  The per-stage test entry points (infer_admm_batch, spectral_init_batch) run on the first GPU. */
int twoace_create_multi(const int* devices, int n_dev, twoace_ctx** ctx);
/* GPUs behind the context (1 for twoace_create). */
int twoace_device_count(const twoace_ctx* ctx);
void twoace_destroy(twoace_ctx* ctx);
const char* twoace_last_error(const twoace_ctx* ctx);
/* CUDA stream the context launches on (cudaStream_t as void*), for event timing by callers. */
void* twoace_stream(twoace_ctx* ctx);
/* Kernels launched by this context since creation (bench.py's gpu_launches). */
int64_t twoace_launch_count(const twoace_ctx* ctx);
/* Block until all work queued by the context has finished. */
int twoace_synchronize(twoace_ctx* ctx);

/* Full solve of nb independent instances — replaces inferLowRankV4 / _multi / _Nuclear
 * (see the variant enum).  A is dense per instance.
 *   m[nb]            rows per instance (host)
 *   A                concat of m_b x n column-major complex matrices (n = tx*rx)
 *   B                concat of m_b real RSS amplitudes
 *   train_idx        host; per instance T draws of floor(m_b*cc_frac) ints (T = 3 for V4_MULTI else 1)
 *   X                out, nb x n complex
 *   Y                out, concat of m_b complex (rows beyond info[5] are zero after a roll-back)
 *   quality          out, nb
 *   info             out or NULL, nb x TWOACE_INFO_WORDS doubles (host or device per `mem`)
 *   stage_words      out or NULL, nb x (4T+1) x TWOACE_STAGE_WORDS doubles; stage order per trial:
 *                    over-param, refinement, rank-one over-param, rank-one refinement; then refine
 */
int twoace_solve_batch(twoace_ctx* ctx, int variant, int mem, int nb, int tx, int rx,
                       const int32_t* m, const double* A, const double* B, const int32_t* train_idx,
                       const twoace_params* params, double* X, double* Y, double* quality,
                       double* info, double* stage_words);

/* Residual trace (the optional fourth output of SURVEY.md section 8b): after twoace_set_trace(ctx, mem, buf, cap) every
 * twoace_solve_batch / twoace_solve_batch_codebook call also writes, per instance and InferADMM stage, the combined
 * residual res_comb (inferLowRankV4.m:345) of every executed iteration:
 *   buf[(b * (4T+1) + stage) * maxiter + (it - 1)],   NaN for iterations / stages that did not run
 * (stage order as in stage_words).  buf holds `capacity` doubles, host or device per `mem`; device buffers need an
 * even nb * (4T+1) * maxiter.  twoace_set_trace(ctx, mem, NULL, 0) switches the trace off again. */
int twoace_set_trace(twoace_ctx* ctx, int mem, double* trace, int64_t capacity);

/* Register a codebook shared by all instances: rows x n complex, column-major (the `cb` variable of
 * codebook/codebook_mat, the .mat files), host or device per `mem`.  Kept on the device until replaced. */
int twoace_set_codebook(twoace_ctx* ctx, int mem, int rows, int n, const double* cb);

/* Same as twoace_solve_batch with A_b = row_scale * cb[cb_rows_b, :] (the row selection of
 * channel_recovery_ADMM_v2_simulation_A2only.m:137-138).  cb_rows: host, concat of m_b row ids. */
int twoace_solve_batch_codebook(twoace_ctx* ctx, int variant, int mem, int nb, int tx, int rx,
                                const int32_t* m, const int32_t* cb_rows, double row_scale,
                                const double* B, const int32_t* train_idx, const twoace_params* params,
                                double* X, double* Y, double* quality, double* info,
                                double* stage_words);

/* One InferADMM call per instance — inferLowRankV4.m:260-365 — exposed for per-stage parity tests.
 *   X0     concat of n x r complex start points (same r for all instances)
 *   X, Y   out: n x rout and m_b x rout per instance, rout = r if scale_by_row else 1
 *   state  out or NULL: per instance final [X Z N (n x r each) | Y M (m_b x r each)] complex
 *   words  out or NULL: nb x TWOACE_STAGE_WORDS
 */
int twoace_infer_admm_batch(twoace_ctx* ctx, int mem, int nb, int tx, int rx, const int32_t* m,
                            const double* A, const double* B, int r, const double* X0,
                            int scale_by_row, int use_rank_one, int nuclear,
                            const twoace_params* params, double* X, double* Y, double* state,
                            double* words);

/* SpectralInitialize (inferLowRankV4.m:540-553) per instance: Xs out, nb x (n x r) complex. */
int twoace_spectral_init_batch(twoace_ctx* ctx, int mem, int nb, int n, const int32_t* m,
                               const double* A, const double* B, int r, double* Xs);

/* ---- PhaseLift ---------------------------------------------------------------
 * Replaces  recoveredSig = MyPhaseLift(measurements, measurementMat)
 * (main/src/my_recovery_algorithms/MyPhaseLift.m:69-107): trace-regularised least squares over PSD matrices,
 *   min 0.5 || y - diag(A X A') ||^2 + lambda trace(X),  X >= 0,
 * solved by TFOCS' Auslender-Teboulle method (solver_TraceLS.m:42, tfocs_AT.m:20-94, prox_trace.m:62-168) from
 * X0 = 0; the result is sqrt(lambda_max) * leading eigenvector of X (global phase arbitrary, as from eig()).
 * `y` are intensities (the squaring and the 2e5 / 1e10 scalings of Recover_Channel.m:35 are caller-side). */
typedef struct {
  int32_t maxIts;        /* 4000   MyPhaseLift.m:83 */
  double tol;            /* 1e-10  MyPhaseLift.m:84 */
  int32_t restart;       /* 200    MyPhaseLift.m:85 */
  double lambda;         /* 5e-2   MyPhaseLift.m:92 */
  double alpha;          /* 0.9    tfocs_initialize.m:22 */
  double beta;           /* 0.5    tfocs_initialize.m:21 */
  double L0;             /* 1      tfocs_initialize.m:23 */
  int32_t cntr_reset;    /* 50     tfocs_initialize.m:31,201-207 */
  double backtrack_tol;  /* 1e-10  tfocs_initialize.m:425 */
  int32_t reduce;        /* 1: iterate in the row space of A when m < n (exact; see csrc/phaselift.cuh) */
} twoace_pl_opts;

void twoace_pl_default_opts(twoace_pl_opts* o);

#define TWOACE_PL_INFO_WORDS 16
/* info[b*16 + k]: 0 TFOCS iterations, 1 prox_trace evaluations (= eigendecompositions), 2 backtracking steps,
 * 3 status (1 step-size tolerance, 2 iteration limit, 3 ||dx|| = 0, 4 NaN, 5 small step, -2 not run), 4 rank of the last
 * prox output, 5 final Lipschitz estimate L, 6 dimension iterated in (m if reduced, else n), 7 lambda_max,
 * 8 Jacobi sweeps over all eigendecompositions, 9..13 device clock cycles: gradient GEMM, warm-start transform,
 * Jacobi, z and A(z), x update and tests; 14..15 reserved. */

/* Batch of independent PhaseLift solves, ragged in m.  Sensing matrices: dense `A` (concatenated m_b x n
 * column-major blocks) or, with A == NULL, rows `cb_rows` of the registered codebook scaled by `row_scale`.
 * sig: nb x n complex out.  n <= 256 with any m <= 4096; 256 < n <= 2000 (up to the 32 x 32 array and beyond, the
 * range below the reference's opts.largescale switch, MyPhaseLift.m:87-89) with reduce = 1 and every m <= 256: the
 * iteration then runs in the m-dimensional row space of A.  Linearly dependent rows have no row-space factor; at
 * n > 256 such an instance gets status -2 and a NaN signal, and a TWOACE_MEM_HOST call returns TWOACE_E_UNSUPPORTED. */
int twoace_phaselift_batch(twoace_ctx* ctx, int mem, int nb, int n, const int32_t* m, const double* A,
                           const int32_t* cb_rows, double row_scale, const double* y,
                           const twoace_pl_opts* opts, double* sig, double* info);

/* ---- Evaluation metrics ---------------------------------------------------------------
 * Per instance, Evaluation_H.m:81-115: out[b*4 + {0,1,2,3}] = MSE_H, gain_ana, gain_dig, proj_error of the
 * recovered channel X_est[b] against the ground truth X_true[b] (vec of the rx x tx channel matrix, column-major,
 * tx, rx <= 32).  Singular pairs are phase-normalised canonically (largest-modulus entry of v real positive);
 * a NaN or all-zero estimate yields NaN metrics. */
#define TWOACE_METRIC_WORDS 4
int twoace_metrics_batch(twoace_ctx* ctx, int mem, int nb, int tx, int rx, const double* X_est,
                         const double* X_true, int phase_bit, double* out);

/* AoD / AoA estimation error per instance (Numerical_Simulation/src/evaluate_plot_results/Evaluation_Recovery.m:85-146)
 * of an H-domain estimate X_est[b] (vec of the nr x nt channel): its angular spectrum z = vec(A_Rx' H A_Tx) on the
 * nqt x nqr virtual-angle dictionary of generate_channel/Sparse_Channel_Formulation.m:83-103, restricted to the
 * searching area (degrees, :120-152), stands in for `recoveredSig`; the L largest entries give the estimated angles.
 * angles_true: nb x 2L (AoD then AoA of the L paths, degrees -- the `angles` output of twoace_synth_batch).
 * out[b*6 + k]: 0 AoD_Err_to_True, 1 AoA_Err_to_True, 2 AoDA_Err, 3..5 the same against the quantised true angles
 * (degrees; NaN for a non-finite estimate).  Reference grid: nqt = 4 nt, nqr = 4 nr (Vs_M_par.m:80-81). */
#define TWOACE_ANGLE_WORDS 6
int twoace_angle_metrics_batch(twoace_ctx* ctx, int mem, int nb, int nt, int nr, int L, int nqt, int nqr,
                               double searching_area, double wavelength, double spacing, const double* X_est,
                               const double* angles_true, double* out);

/* ---- On-device instance synthesis -----------------------------------------------------
 * Builds nb independent (trial, M, SNR) instances of the numerical-simulation workload on the GPU, so that the
 * 100k / 1M-trial configurations need no per-instance host->device input traffic:
 *   channel      Eq. 23 sparse multipath, L paths, AoD/AoA ~ U(-area/2, area/2) degrees, gains CN(0,1) normalised
 *                (Numerical_Simulation/src/generate_channel/Generate_Channel.m:76-139)
 *   probes       m_b rows without replacement from rows [row_lo_b, row_hi_b) of the registered codebook
 *                (randperm, main/channel_recovery_ADMM_v2_simulation_A2only.m:137; ..._multiresolution.m:137-143)
 *   RSS          B = | row_scale * cb[rows, :] vecH + CN(0, 10^(-snr_db/10)) |
 *                (Numerical_Simulation/src/generate_measurement/Generate_Measurement.m:84-101)
 *   train draws  ntrain draws of floor(m_b * cc_frac) measurement ids (randsample, inferLowRankV4.m:36-37)
 * The random stream is Philox4x32-10 keyed by `seed` and counted by (index, stream, trial_id) -- see csrc/synth.cuh for
 * the exact stream layout (restated by oracle/synth.py) -- so an instance depends only on (seed, trial_id): results
 * are independent of batch composition and of how trials are sharded over GPUs.
 * Index outputs (cb_rows: concat of m_b; train_idx: concat of ntrain x floor(m_b * cc_frac)) are written to HOST
 * memory, ready to be passed to twoace_solve_batch_codebook; B (concat of m_b), vecH (nb x n complex) and angles
 * (nb x 2L: AoD then AoA, degrees; may be NULL) follow `mem`.  row_hi_b - row_lo_b <= 8192. */
typedef struct {
  int32_t nt, nr, L;         /* transmit / receive antennas (n = nt * nr), paths (<= 32) */
  double searching_area;     /* degrees: 95 (A2only.m:52) */
  double wavelength;         /* 3e8 / 60.48e9 (A2only.m:38) */
  double spacing;            /* 3.055e-3 (A2only.m:39) */
  double row_scale;          /* 1 / sqrt(n): unit-norm probes, signal power 1 */
  double cc_frac;            /* 0.95 */
  int32_t ntrain;            /* 1, or 3 for inferLowRankV4_multi */
  uint64_t seed;
} twoace_synth_params;
void twoace_synth_default_params(twoace_synth_params* p, int nt, int nr);
int twoace_synth_batch(twoace_ctx* ctx, int mem, int nb, const twoace_synth_params* sp, const int32_t* m,
                       const double* snr_db, const int32_t* row_lo, const int32_t* row_hi, const int64_t* trial_id,
                       int32_t* cb_rows, int32_t* train_idx, double* B, double* vecH, double* angles);

/* Execution options.  "fast" (default 1): run eligible InferADMM launches (16x16, quantised 4-phase A,
 * r in {20,1}, m <= 1024, V4 or nuclear ArgMinZ) on the shared-memory cluster kernels instead of the general one, and let
 * the general kernel keep the 2-bit codes of a quantised A in shared memory (any n divisible by 16); 0 = general
 * kernel with dense products for everything;
 * "fast_cs" (2 or 4, default 2): cluster size of the r = 20 stages; "chunk": instances per internal pass;
 * "dedup_nuclear_rerun" (default 0): inferLowRank_Nuclear.m:69-70 reruns the train solve with use_rank_one = true,
 * a flag its ArgMinZ (:411-419) never reads, so the rerun reproduces the first run bit for bit; 1 = do not
 * re-execute it (identical X, Y, quality, flags; roughly 1.6x fewer iterations when quality < 0.6);
 * "tensor" (default 1): the two sensing-matrix products of the cluster kernel run as exact int8 tensor-core products
 * (tcgen05, csrc/tc_prod.cuh), 0 = FP64 SIMT products; "cache_sinv" (default 1): the stages of one trial share one
 * (I + A A')^-1 instead of inverting it once per stage; "spectral_jacobi" (default 0): 1 = SpectralInitialize uses the
 * full Jacobi eigendecomposition for every size instead of the leading-eigenpair solver (csrc/tridiag_eig.cuh);
 * "overlap" (default 1): the general-kernel group of a mixed InferADMM launch (tiny problems) runs on a second stream
 * beside the cluster-kernel groups. */
int twoace_set_option(twoace_ctx* ctx, const char* key, int value);
/* InferADMM launches that took the shared-memory cluster kernel since context creation. */
int64_t twoace_fast_launch_count(const twoace_ctx* ctx);
/* launches of the cluster kernel whose A-products ran as exact int8 tensor-core (tcgen05) products */
int64_t twoace_tensor_launch_count(const twoace_ctx* ctx);

/* Measurement hooks (bench.py).  With timing on, every InferADMM stage-kernel launch is bracketed by
 * CUDA events on the context's stream; twoace_timing_collect synchronises, returns the summed
 * duration (ms) and launch count since the last collect, and resets the accumulators. */
int twoace_set_timing(twoace_ctx* ctx, int on);
int twoace_timing_collect(twoace_ctx* ctx, double* stage_ms, int64_t* stage_launches);
/* FP64 FMA throughput of this GPU (TFLOP/s, 2 flops per DFMA), measured with a register-resident
 * DFMA loop: the roofline denominator of the FP64-pipe-bound stage kernel (MEASURED_PEAKS.json
 * only carries HBM and bf16 figures). */
int twoace_fp64_peak(twoace_ctx* ctx, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* TWOACE_H */
