// MEX shim: [X, Y, quality] = inferLowRankV4(A, B, tx, rx, lambda, r, mu0, rho, cc_frac, tol_rel, tol_abs, maxiter)
// Same base name on the MATLAB path shadows main/src/my_recovery_algorithms/ADMM_v2/inferLowRankV4.m.
// Build three times with -DTWOACE_VARIANT=0|1|2 and -output inferLowRankV4 | inferLowRankV4_multi |
// inferLowRank_Nuclear:
//   mex -R2018a -DTWOACE_HAVE_MEX -DTWOACE_VARIANT=1 -output inferLowRankV4_multi mex/twoace_mex.cpp -Iinclude -L<dir> -ltwoace
// (without -DTWOACE_HAVE_MEX the file compiles to an empty object: the guard keeps the tree buildable without mex.h)
// Not compiled in this repository's CI: mex.h / libmex are absent (no MATLAB in the image).  The code only
// uses the documented interleaved-complex MEX API and the C ABI of include/twoace.h.
#ifdef TWOACE_HAVE_MEX
#include <cmath>
#include <cstring>
#include <vector>

#include "mex.h"
#include "twoace.h"

#ifndef TWOACE_VARIANT
#define TWOACE_VARIANT TWOACE_V4
#endif

static twoace_ctx* g_ctx = nullptr;
static void at_exit() { if (g_ctx) { twoace_destroy(g_ctx); g_ctx = nullptr; } }

static double scalar_arg(int nrhs, const mxArray* prhs[], int idx, double dflt) {
  return (nrhs > idx && !mxIsEmpty(prhs[idx])) ? mxGetScalar(prhs[idx]) : dflt;   // nargin defaults, :2-9
}

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs < 4 || nrhs > 12) mexErrMsgIdAndTxt("twoace:nargin", "expected 4 to 12 inputs");
  if (nlhs > 3) mexErrMsgIdAndTxt("twoace:nargout", "at most 3 outputs");
  if (!g_ctx) {
    if (twoace_create(0, &g_ctx) != TWOACE_OK) mexErrMsgIdAndTxt("twoace:cuda", "no usable CUDA device");
    mexAtExit(at_exit);
  }
  const mxArray* Am = prhs[0];
  const int32_t m = (int32_t)mxGetM(Am), n = (int32_t)mxGetN(Am);
  const int tx = (int)mxGetScalar(prhs[2]), rx = (int)mxGetScalar(prhs[3]);
  if (tx * rx != n) mexErrMsgIdAndTxt("twoace:size", "size(A,2) must equal tx*rx");
  // A as interleaved complex double (a real A is widened)
  std::vector<double> Abuf((size_t)2 * m * n);
  if (mxIsComplex(Am)) {
    std::memcpy(Abuf.data(), mxGetComplexDoubles(Am), sizeof(double) * 2 * m * n);
  } else {
    const double* ar = mxGetDoubles(Am);
    for (size_t i = 0; i < (size_t)m * n; ++i) { Abuf[2 * i] = ar[i]; Abuf[2 * i + 1] = 0.0; }
  }
  if ((int32_t)mxGetNumberOfElements(prhs[1]) != m) mexErrMsgIdAndTxt("twoace:size", "B must have size(A,1) entries");
  const double* B = mxGetDoubles(prhs[1]);
  twoace_params p;
  twoace_default_params(&p);
  p.lambda = scalar_arg(nrhs, prhs, 4, p.lambda);
  p.r = (int32_t)scalar_arg(nrhs, prhs, 5, p.r);
  p.mu0 = scalar_arg(nrhs, prhs, 6, p.mu0);
  p.rho = scalar_arg(nrhs, prhs, 7, p.rho);
  p.cc_frac = scalar_arg(nrhs, prhs, 8, p.cc_frac);
  p.tol_rel = scalar_arg(nrhs, prhs, 9, p.tol_rel);
  p.tol_abs = scalar_arg(nrhs, prhs, 10, p.tol_abs);
  p.maxiter = (int32_t)scalar_arg(nrhs, prhs, 11, p.maxiter);
  // the randsample draws of inferLowRankV4.m:37 (x3 in _multi.m:48) come from MATLAB itself, in the
  // original order, so the global RNG stream is consumed exactly as by the .m file
  const int T = (TWOACE_VARIANT == TWOACE_V4_MULTI) ? 3 : 1;
  const int k = (int)std::floor(m * p.cc_frac);
  std::vector<int32_t> train((size_t)T * k);
  for (int t = 0; t < T; ++t) {
    mxArray* in[2] = {mxCreateDoubleScalar(m), mxCreateDoubleScalar(k)};
    mxArray* out[1];
    mexCallMATLAB(1, out, 2, in, "randsample");
    const double* v = mxGetDoubles(out[0]);
    for (int i = 0; i < k; ++i) train[(size_t)t * k + i] = (int32_t)v[i] - 1;   // 1-based -> 0-based
    mxDestroyArray(in[0]); mxDestroyArray(in[1]); mxDestroyArray(out[0]);
  }
  plhs[0] = mxCreateDoubleMatrix(n, 1, mxCOMPLEX);
  mxArray* Ym = mxCreateDoubleMatrix(m, 1, mxCOMPLEX);
  double quality = 0.0, info[TWOACE_INFO_WORDS];
  const int rc = twoace_solve_batch(g_ctx, TWOACE_VARIANT, TWOACE_MEM_HOST, 1, tx, rx, &m, Abuf.data(), B,
                                    train.data(), &p, (double*)mxGetComplexDoubles(plhs[0]),
                                    (double*)mxGetComplexDoubles(Ym), &quality, info, nullptr);
  if (rc != TWOACE_OK) mexErrMsgIdAndTxt("twoace:solve", "%s", twoace_last_error(g_ctx));
  if (nlhs > 1) {
    mxSetM(Ym, (mwSize)info[5]);   // rows of Y: m, or m_train after the roll-back of :72-77
    plhs[1] = Ym;
  } else {
    mxDestroyArray(Ym);
  }
  if (nlhs > 2) plhs[2] = mxCreateDoubleScalar(quality);
}
#endif  // TWOACE_HAVE_MEX
