// MEX shim: [X, Y, quality] = inferMinL2(A, B, lambda, r, tol_rel, tol_abs, maxiter)
// Same base name on the MATLAB path shadows main/src/my_recovery_algorithms/ADMM_v2/inferMinL2.m (ADMM_v2.m:23, version 0).
//   mex -R2018a -DTWOACE_HAVE_MEX -output inferMinL2 mex/inferminl2_mex.cpp -Iinclude -L<dir> -ltwoace
// (without -DTWOACE_HAVE_MEX the file compiles to an empty object: the guard keeps the tree buildable without mex.h)
// Not compiled in this repository's CI: mex.h / libmex are absent (no MATLAB in the image).
#ifdef TWOACE_HAVE_MEX
#include <cmath>
#include <cstring>
#include <vector>

#include "mex.h"
#include "twoace.h"

static twoace_ctx* g_ctx = nullptr;
static void at_exit() { if (g_ctx) { twoace_destroy(g_ctx); g_ctx = nullptr; } }

static double scalar_arg(int nrhs, const mxArray* prhs[], int idx, double dflt) {
  return (nrhs > idx && !mxIsEmpty(prhs[idx])) ? mxGetScalar(prhs[idx]) : dflt;   // nargin defaults, inferMinL2.m:2-6
}

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs < 2 || nrhs > 7) mexErrMsgIdAndTxt("twoace:nargin", "expected 2 to 7 inputs");
  if (nlhs > 3) mexErrMsgIdAndTxt("twoace:nargout", "at most 3 outputs");
  if (!g_ctx) {
    if (twoace_create(0, &g_ctx) != TWOACE_OK) mexErrMsgIdAndTxt("twoace:cuda", "no usable CUDA device");
    mexAtExit(at_exit);
  }
  const mxArray* Am = prhs[0];
  const int32_t m = (int32_t)mxGetM(Am), n = (int32_t)mxGetN(Am);
  // the solver only needs n = tx * rx: any factorisation the library accepts (tx a multiple of 4, at most 32)
  int tx = 0;
  for (int t : {16, 32, 8, 4}) if (n % t == 0) { tx = t; break; }
  if (!tx) mexErrMsgIdAndTxt("twoace:size", "size(A,2) must be divisible by 4");
  std::vector<double> Abuf((size_t)2 * m * n);
  if (mxIsComplex(Am)) {
    std::memcpy(Abuf.data(), mxGetComplexDoubles(Am), sizeof(double) * 2 * m * n);
  } else {
    const double* ar = mxGetDoubles(Am);
    for (size_t i = 0; i < (size_t)m * n; ++i) { Abuf[2 * i] = ar[i]; Abuf[2 * i + 1] = 0.0; }
  }
  if ((int32_t)mxGetNumberOfElements(prhs[1]) != m) mexErrMsgIdAndTxt("twoace:size", "B must have size(A,1) entries");
  twoace_params p;
  twoace_default_params(&p);
  p.lambda = scalar_arg(nrhs, prhs, 2, 0.0);
  p.r = (int32_t)scalar_arg(nrhs, prhs, 3, 20);
  p.tol_rel = scalar_arg(nrhs, prhs, 4, 1e-4);
  p.tol_abs = scalar_arg(nrhs, prhs, 5, 1e-8);
  p.maxiter = (int32_t)scalar_arg(nrhs, prhs, 6, 500);
  // train_idx = randsample(m, ceil(m*0.95))  (inferMinL2.m:34), drawn by MATLAB itself so the global RNG stream is
  // consumed exactly as by the .m file
  const int k = (int)std::ceil(m * 0.95);
  std::vector<int32_t> train((size_t)k);
  {
    mxArray* in[2] = {mxCreateDoubleScalar(m), mxCreateDoubleScalar(k)};
    mxArray* out[1];
    mexCallMATLAB(1, out, 2, in, "randsample");
    const double* v = mxGetDoubles(out[0]);
    for (int i = 0; i < k; ++i) train[i] = (int32_t)v[i] - 1;
    mxDestroyArray(in[0]); mxDestroyArray(in[1]); mxDestroyArray(out[0]);
  }
  plhs[0] = mxCreateDoubleMatrix(n, 1, mxCOMPLEX);
  mxArray* Ym = mxCreateDoubleMatrix(m, 1, mxCOMPLEX);
  double quality = 0.0, info[TWOACE_INFO_WORDS];
  const int rc = twoace_solve_batch(g_ctx, TWOACE_MINL2, TWOACE_MEM_HOST, 1, tx, n / tx, &m, Abuf.data(),
                                    mxGetDoubles(prhs[1]), train.data(), &p, (double*)mxGetComplexDoubles(plhs[0]),
                                    (double*)mxGetComplexDoubles(Ym), &quality, info, nullptr);
  if (rc != TWOACE_OK) mexErrMsgIdAndTxt("twoace:solve", "%s", twoace_last_error(g_ctx));
  if (nlhs > 1) {
    mxSetM(Ym, (mwSize)info[5]);   // rows of Y: m after the refinement, ceil(0.95 m) otherwise (:42-58)
    plhs[1] = Ym;
  } else {
    mxDestroyArray(Ym);
  }
  if (nlhs > 2) plhs[2] = mxCreateDoubleScalar(quality);
}
#endif  // TWOACE_HAVE_MEX
