/* Plain-C client of the C ABI (include/twoace.h): what a cgo / JNI / MEX binding would do.
 * Built and run by tests/test_abi.py::test_c_client_* (gcc, no C++/torch types on this side of the boundary).
 * Usage: abi_smoke           -> checks the symbols link and the defaults (no GPU needed)
 *        abi_smoke solve     -> runs one inferLowRankV4 solve, one PhaseLift solve and the metrics on cuda:0 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "twoace.h"

static double frand(unsigned* s) { *s = *s * 1664525u + 1013904223u; return ((*s >> 8) & 0xFFFFFF) / 16777216.0; }

int main(int argc, char** argv) {
  twoace_params p;
  twoace_default_params(&p);
  twoace_pl_opts o;
  twoace_pl_default_opts(&o);
  if (p.r != 20 || p.maxiter != 500 || p.mu0 != 1e-3 || o.maxIts != 4000 || o.restart != 200 || o.lambda != 0.05) {
    printf("FAIL defaults\n");
    return 1;
  }
  printf("version %d\n", twoace_version());
  if (argc < 2 || strcmp(argv[1], "solve") != 0) { printf("OK link\n"); return 0; }

  twoace_ctx* ctx = NULL;
  if (twoace_create(0, &ctx) != TWOACE_OK) { printf("FAIL create (no GPU?)\n"); return 2; }
  enum { TX = 4, RX = 4, N = 16, M = 40 };
  static double A[2 * M * N], B[M], Y2[M], X[2 * N], Y[2 * M], h[2 * N], sig[2 * N];
  unsigned s = 7u;
  /* rank-one channel H = a b' (vec, column-major rx x tx): what the rank-shaping ArgMinZ is built for */
  double a[2 * RX], b[2 * TX];
  for (int k = 0; k < RX; ++k) { a[2 * k] = frand(&s) - 0.5; a[2 * k + 1] = frand(&s) - 0.5; }
  for (int k = 0; k < TX; ++k) { b[2 * k] = frand(&s) - 0.5; b[2 * k + 1] = frand(&s) - 0.5; }
  for (int t = 0; t < TX; ++t)
    for (int r = 0; r < RX; ++r) {
      h[2 * (r + RX * t)] = a[2 * r] * b[2 * t] + a[2 * r + 1] * b[2 * t + 1];          /* a_r conj(b_t) */
      h[2 * (r + RX * t) + 1] = a[2 * r + 1] * b[2 * t] - a[2 * r] * b[2 * t + 1];
    }
  for (int k = 0; k < N; ++k)           /* column-major m x n, 2-bit phases, unit-norm rows */
    for (int i = 0; i < M; ++i) {
      const int code = (int)(frand(&s) * 4.0) & 3;
      const double re[4] = {1, 0, -1, 0}, im[4] = {0, 1, 0, -1};
      A[2 * (i + M * k)] = re[code] / 4.0;
      A[2 * (i + M * k) + 1] = im[code] / 4.0;
    }
  for (int i = 0; i < M; ++i) {
    double yr = 0, yi = 0;
    for (int k = 0; k < N; ++k) {
      const double ar = A[2 * (i + M * k)], ai = A[2 * (i + M * k) + 1];
      yr += ar * h[2 * k] - ai * h[2 * k + 1];
      yi += ar * h[2 * k + 1] + ai * h[2 * k];
    }
    B[i] = sqrt(yr * yr + yi * yi);
    Y2[i] = B[i] * B[i];
  }
  int32_t m = M, train[38];
  for (int i = 0; i < 38; ++i) train[i] = i + 1;   /* rows 1..38 train, rows 0 and 39 held out */
  double quality = 0, info[TWOACE_INFO_WORDS], met[TWOACE_METRIC_WORDS], plinfo[TWOACE_PL_INFO_WORDS];
  int rc = twoace_solve_batch(ctx, TWOACE_V4, TWOACE_MEM_HOST, 1, TX, RX, &m, A, B, train, &p, X, Y, &quality, info, NULL);
  if (rc != TWOACE_OK) { printf("FAIL solve: %s\n", twoace_last_error(ctx)); return 3; }
  rc = twoace_metrics_batch(ctx, TWOACE_MEM_HOST, 1, TX, RX, X, h, 2, met);
  if (rc != TWOACE_OK) { printf("FAIL metrics: %s\n", twoace_last_error(ctx)); return 4; }
  printf("V4: quality %.6f MSE_H %.3e launches %lld\n", quality, met[0], (long long)twoace_launch_count(ctx));
  if (!(quality > 0.9) || !(met[0] < 1e-4)) { printf("FAIL accuracy\n"); return 5; }
  o.maxIts = 400;
  rc = twoace_phaselift_batch(ctx, TWOACE_MEM_HOST, 1, N, &m, A, NULL, 1.0, Y2, &o, sig, plinfo);
  if (rc != TWOACE_OK) { printf("FAIL phaselift: %s\n", twoace_last_error(ctx)); return 6; }
  rc = twoace_metrics_batch(ctx, TWOACE_MEM_HOST, 1, TX, RX, sig, h, 2, met);
  printf("PhaseLift: iterations %.0f MSE_H %.3e\n", plinfo[0], met[0]);
  /* error path: bad variant must fail with a message, and the context must stay usable */
  rc = twoace_solve_batch(ctx, 99, TWOACE_MEM_HOST, 1, TX, RX, &m, A, B, train, &p, X, Y, &quality, NULL, NULL);
  if (rc != TWOACE_E_INVALID || strlen(twoace_last_error(ctx)) == 0) { printf("FAIL error path\n"); return 7; }
  /* the simulation pipeline without per-instance host data: register the sensing rows as a codebook, build instances on
   * the GPU (twoace_synth_batch), solve them in codebook mode, evaluate (Evaluation_H.m, Evaluation_Recovery.m) */
  {
    enum { NB = 3, MS = 24 };
    static double cb[2 * M * N];
    for (int i = 0; i < 2 * M * N; ++i) cb[i] = A[i] * 4.0;          /* unit-modulus 4-phase entries, column-major M x N */
    rc = twoace_set_codebook(ctx, TWOACE_MEM_HOST, M, N, cb);
    if (rc != TWOACE_OK) { printf("FAIL set_codebook: %s\n", twoace_last_error(ctx)); return 8; }
    twoace_synth_params sp;
    twoace_synth_default_params(&sp, TX, RX);
    int32_t ms[NB] = {MS, MS, MS}, lo[NB] = {0, 0, 0}, hi[NB] = {M, M, M};
    double snr[NB] = {30.0, 30.0, 30.0};
    int64_t trial[NB] = {1, 2, 3};
    static int32_t rows[NB * MS], tr[NB * MS];
    static double Bs[NB * MS], Hs[2 * NB * N], ang[NB * 6], Xs[2 * NB * N], Ys[2 * NB * MS], qs[NB], mets[NB * TWOACE_METRIC_WORDS],
        angm[NB * TWOACE_ANGLE_WORDS];
    rc = twoace_synth_batch(ctx, TWOACE_MEM_HOST, NB, &sp, ms, snr, lo, hi, trial, rows, tr, Bs, Hs, ang);
    if (rc != TWOACE_OK) { printf("FAIL synth: %s\n", twoace_last_error(ctx)); return 9; }
    for (int i = 0; i < NB * MS; ++i)
      if (rows[i] < 0 || rows[i] >= M || !(Bs[i] >= 0.0)) { printf("FAIL synth output\n"); return 10; }
    rc = twoace_solve_batch_codebook(ctx, TWOACE_V4, TWOACE_MEM_HOST, NB, TX, RX, ms, rows, sp.row_scale, Bs, tr, &p, Xs, Ys, qs,
                                     NULL, NULL);
    if (rc != TWOACE_OK) { printf("FAIL codebook solve: %s\n", twoace_last_error(ctx)); return 11; }
    rc = twoace_metrics_batch(ctx, TWOACE_MEM_HOST, NB, TX, RX, Xs, Hs, 2, mets);
    if (rc == TWOACE_OK)
      rc = twoace_angle_metrics_batch(ctx, TWOACE_MEM_HOST, NB, TX, RX, sp.L, 4 * TX, 4 * RX, sp.searching_area, sp.wavelength,
                                      sp.spacing, Xs, ang, angm);
    if (rc != TWOACE_OK) { printf("FAIL evaluation: %s\n", twoace_last_error(ctx)); return 12; }
    printf("synth + codebook solve: quality %.3f %.3f %.3f, MSE_H %.2e, AoDA error %.2f deg\n", qs[0], qs[1], qs[2], mets[0], angm[2]);
  }
  /* one context over a set of GPUs (here: the one GPU this client assumes) */
  {
    twoace_ctx* mctx = NULL;
    int dev[1] = {0};
    if (twoace_create_multi(dev, 1, &mctx) != TWOACE_OK || twoace_device_count(mctx) != 1) { printf("FAIL create_multi\n"); return 13; }
    twoace_destroy(mctx);
  }
  twoace_destroy(ctx);
  printf("OK solve\n");
  return 0;
}
