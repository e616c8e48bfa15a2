// MEX shim: recoveredSig = MyPhaseLift(measurements, measurementMat)
// Same base name on the MATLAB path shadows main/src/my_recovery_algorithms/MyPhaseLift.m:69.
//   mex -R2018a -DTWOACE_HAVE_MEX -output MyPhaseLift mex/myphaselift_mex.cpp -Iinclude -L<dir> -ltwoace
// Not compiled in this repository's CI: mex.h / libmex are absent (no MATLAB in the image).  The code only
// uses the documented interleaved-complex MEX API and the C ABI of include/twoace.h.
#ifdef TWOACE_HAVE_MEX
#include <cstring>
#include <vector>

#include "mex.h"
#include "twoace.h"

static twoace_ctx* g_ctx = nullptr;
static void at_exit() { if (g_ctx) { twoace_destroy(g_ctx); g_ctx = nullptr; } }

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs != 2) mexErrMsgIdAndTxt("twoace:nargin", "expected (measurements, measurementMat)");
  if (nlhs > 1) mexErrMsgIdAndTxt("twoace:nargout", "one output");
  if (!g_ctx) {
    if (twoace_create(0, &g_ctx) != TWOACE_OK) mexErrMsgIdAndTxt("twoace:cuda", "no usable CUDA device");
    mexAtExit(at_exit);
  }
  const mxArray* Am = prhs[1];
  const int32_t m = (int32_t)mxGetM(Am), n = (int32_t)mxGetN(Am);            // MyPhaseLift.m:71
  if ((int32_t)mxGetNumberOfElements(prhs[0]) != m) mexErrMsgIdAndTxt("twoace:size", "measurements must have size(A,1) entries");
  if (mxIsComplex(prhs[0])) mexErrMsgIdAndTxt("twoace:type", "measurements (intensities) must be real");
  std::vector<double> Abuf((size_t)2 * m * n);
  if (mxIsComplex(Am)) {
    std::memcpy(Abuf.data(), mxGetComplexDoubles(Am), sizeof(double) * 2 * m * n);
  } else {
    const double* ar = mxGetDoubles(Am);
    for (size_t i = 0; i < (size_t)m * n; ++i) { Abuf[2 * i] = ar[i]; Abuf[2 * i + 1] = 0.0; }
  }
  twoace_pl_opts o;
  twoace_pl_default_opts(&o);                                               // MyPhaseLift.m:83-92
  plhs[0] = mxCreateDoubleMatrix(n, 1, mxCOMPLEX);
  const int rc = twoace_phaselift_batch(g_ctx, TWOACE_MEM_HOST, 1, n, &m, Abuf.data(), nullptr, 1.0,
                                        mxGetDoubles(prhs[0]), &o, (double*)mxGetComplexDoubles(plhs[0]), nullptr);
  if (rc != TWOACE_OK) mexErrMsgIdAndTxt("twoace:solve", "%s", twoace_last_error(g_ctx));
}
#endif  // TWOACE_HAVE_MEX
