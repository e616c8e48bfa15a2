// Host/device shared task descriptors for the twoace kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace twoace {

// A row-subset view of a row-major complex sensing matrix:
//   A_eff[i, k] = (*scale) * base[rows[i] * n + k]      (rows == nullptr -> rows[i] = i)
// Dense inputs are transposed to row-major once per solve; codebook inputs are viewed in place.
struct AView {
  const double2* base;
  const int* rows;
  const double* scale;   // device scalar (1/A_norm, computed on device by the prep kernel)
};

// Solver hyper-parameters (inferLowRankV4.m:2-9) + execution flags.
struct DevParams {
  double mu0, rho, tol_rel, tol_abs;
  int maxiter;
  int need_dual;   // 0 when tol_rel == tol_abs == 0 (res_dual is then dead: :344,:351 can never fire)
};

// One InferADMM call (inferLowRankV4.m:260-365) on one instance.
struct StageTask {
  AView A;
  const double* B;      // B_eff[i] = bscale * B[brows ? brows[i] : i]
  const int* brows;
  const double* bscale; // device scalar (1/B_norm)
  int m, r;             // rows of A_eff, columns of X0
  const int* r_ptr;     // optional device override of r (inferMinL2.m:181-185: rank chosen by the spectral initialisation)
  const double2* X0;    // n x r column-major
  double2* Xout;        // n x (sbr ? r : 1)
  double2* Yout;        // m x (sbr ? r : 1)
  int sbr, rank_one, nuclear;   // nuclear: 0 V4 ArgMinZ, 1 SVT ArgMinZ, 2 no low-rank variable (inferMinL2.m)
  const int* rank_one_ptr;  // optional device override of rank_one (decided by an earlier kernel)
  const int* active;    // device flag (nullptr = always active); inactive tasks return immediately
  int active_expect;    // task runs iff *active == active_expect
  const uint32_t* codes;  // 2-bit phase codes of A.base rows (16 words per row, n = 256) or nullptr
  const double* cscale;   // device scalar: A_eff = (*cscale) * u(code)   (valid when codes != nullptr)
  double* scal;         // [STAGE_SCAL] per-task bookkeeping (may be nullptr)
  double2* state;       // optional dump of final [X Z N (n x r each) | Y M (m x r each)] (may be nullptr)
  double2* sinv;        // optional per-instance store of (I + A A')^-1 (m x m) shared by the stages of one trial
  int sinv_state;       // 0: compute it (into `sinv` when given, else the cluster workspace); 1: reuse `sinv`
  double* trace;        // optional residual trace: trace[it-1] = res_comb of iteration it (:345), maxiter entries
};

enum {
  SC_MU = 0, SC_OPT_OBJ, SC_ITERS, SC_OPT_ITER, SC_OPT_COL, SC_BUMPS, SC_CONVERGED, SC_RES_COMB,
  SC_SWEEPS, STAGE_SCAL = 16
};

}  // namespace twoace
