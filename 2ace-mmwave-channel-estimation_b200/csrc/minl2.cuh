// InferADMM of inferMinL2.m (:227-346; ADMM_v2.m:23, version 0): phase retrieval without the low-rank variable,
//   X = pinv(A) (Y - M/mu),  Y = ArgMinY(A X, B, M, mu),  M += mu (A X - Y),
// with mu = 1e-3, rho = 1.03 hard-wired as in the reference (:269-270) and its own residual thresholds (:316-325).
// One CTA per task, state in a per-CTA global workspace (any n, m, r <= 32; general complex A).  pinv(A) is formed once
// per stage through the smaller Gram matrix -- A'(A A')^-1 for m <= n, (A'A)^-1 A' for m > n -- which equals MATLAB's
// SVD-based pinv for the full-rank sensing matrices of this path (a rank-deficient A gives Inf/NaN, which propagate).
// For m <= n the iteration is stationary after one step (A pinv(A) = I makes A X = Y - M/mu exactly), as in the
// reference; the real work is the m > n half of the measurement sweep.
#pragma once
#include "admm_stage.cuh"

namespace twoace {

struct Minl2Dims { int n, maxm, maxr; size_t ws_stride; };

__host__ __device__ inline size_t minl2_ws_elems(const Minl2Dims& d) {
  const size_t g = (size_t)(d.maxm < d.n ? d.maxm : d.n);
  // Acm | Arm | U (n x m) | Ginv (g x g) | X, AtY (n x r) | AX, Y, M, T, dY (m x r)
  return 3 * (size_t)d.maxm * d.n + g * g + 2 * (size_t)d.n * d.maxr + 5 * (size_t)d.maxm * d.maxr;
}
__host__ __device__ inline size_t minl2_smem_bytes(const Minl2Dims& d) {
  const size_t g = (size_t)(d.maxm < d.n ? d.maxm : d.n);
  return (size_t)QT * RCH * sizeof(cd) + (size_t)NT * RCH * sizeof(cd) + 2 * g * sizeof(cd) +
         (size_t)d.maxm * (sizeof(double) + sizeof(int)) + 16 * NW * sizeof(double) + 2 * SMALL_DMAX * sizeof(double) + 256;
}

__global__ void __launch_bounds__(NT, 2)
minl2_stage_kernel(const StageTask* __restrict__ tasks, int ntasks, DevParams prm, Minl2Dims dm, cd* wsbase) {
  extern __shared__ __align__(16) unsigned char minl2_smem[];
  unsigned char* p = minl2_smem;
  cd* tile = (cd*)p;      p += (size_t)QT * RCH * sizeof(cd);
  cd* ksred = (cd*)p;     p += (size_t)NT * RCH * sizeof(cd);
  const int gmax = dm.maxm < dm.n ? dm.maxm : dm.n;
  cd* colk = (cd*)p;      p += (size_t)gmax * sizeof(cd);
  cd* rowk = (cd*)p;      p += (size_t)gmax * sizeof(cd);
  double* Bs = (double*)p; p += (size_t)dm.maxm * sizeof(double);
  double* red = (double*)p; p += 16 * NW * sizeof(double);
  double* xcol = (double*)p; p += 2 * SMALL_DMAX * sizeof(double);
  int* rows_s = (int*)p;
  const int tid = threadIdx.x, n = dm.n;
  cd* ws = wsbase + (size_t)blockIdx.x * dm.ws_stride;

  for (int t = blockIdx.x; t < ntasks; t += gridDim.x) {
    const StageTask tk = tasks[t];
    if (tk.active != nullptr && *tk.active != tk.active_expect) continue;
    const int m = tk.m, r = tk.r_ptr ? *tk.r_ptr : tk.r;
    const int g = m <= n ? m : n;
    const double asc = *tk.A.scale, bsc = *tk.bscale;
    cd* Acm = ws;                               // A_eff, column-major m x n
    cd* Arm = Acm + (size_t)m * n;              // A_eff, row-major
    cd* U = Arm + (size_t)m * n;                // pinv(A), n x m column-major
    cd* Gi = U + (size_t)m * n;                 // inverse Gram, g x g
    cd* X = Gi + (size_t)g * g;
    cd* AtY = X + (size_t)n * r;
    cd* AX = AtY + (size_t)n * r;
    cd* Y = AX + (size_t)m * r;
    cd* M = Y + (size_t)m * r;
    cd* T = M + (size_t)m * r;
    cd* dY = T + (size_t)m * r;
    __syncthreads();
    for (int i = tid; i < m; i += NT) {
      rows_s[i] = tk.A.rows ? tk.A.rows[i] : i;
      Bs[i] = bsc * tk.B[tk.brows ? tk.brows[i] : i];
    }
    __syncthreads();
    for (size_t idx = tid; idx < (size_t)m * n; idx += NT) {
      const int k = (int)(idx % n), i = (int)(idx / n);
      const cd a = cscale(tk.A.base[(size_t)rows_s[i] * n + k], asc);
      Arm[idx] = a;
      Acm[i + (size_t)m * k] = a;
    }
    double nb2;
    {
      double v[1] = {0.0};
      for (int i = tid; i < m; i += NT) v[0] += Bs[i] * Bs[i];
      block_sum<1>(v, red);
      nb2 = v[0];
    }
    const double normB = sqrt(nb2);
    // ---- U = pinv(A)  (:236)
    if (m <= n) {
      gemm_tpo(m, n, m, [&](int k, int i) { return Acm[i + (size_t)m * k]; },
               [&](int k, int j) { return cconj(Arm[(size_t)j * n + k]); },
               [&](int i, int j, cd v) { Gi[i + (size_t)m * j] = v; }, tile, ksred);            // A A'
      spd_inverse(Gi, m, colk, rowk);
      gemm_tpo(n, m, m, [&](int i, int k) { return cconj(Arm[(size_t)i * n + k]); },
               [&](int i, int j) { return Gi[i + (size_t)m * j]; },
               [&](int k, int j, cd v) { U[k + (size_t)n * j] = v; }, tile, ksred);             // A' (A A')^-1
    } else {
      gemm_tpo(n, m, n, [&](int i, int k) { return cconj(Arm[(size_t)i * n + k]); },
               [&](int i, int l) { return Arm[(size_t)i * n + l]; },
               [&](int k, int l, cd v) { Gi[k + (size_t)n * l] = v; }, tile, ksred);            // A'A
      spd_inverse(Gi, n, colk, rowk);
      gemm_tpo(n, n, m, [&](int l, int k) { return Gi[k + (size_t)n * l]; },
               [&](int l, int i) { return cconj(Arm[(size_t)i * n + l]); },
               [&](int k, int i, cd v) { U[k + (size_t)n * i] = v; }, tile, ksred);             // (A'A)^-1 A'
    }
    auto prod_ax = [&]() {      // AX = A X
      gemm_tpo(m, n, r, [&](int k, int i) { return Acm[i + (size_t)m * k]; },
               [&](int k, int c) { return X[k + (size_t)n * c]; },
               [&](int i, int c, cd v) { AX[i + (size_t)m * c] = v; }, tile, ksred);
    };
    // ---- X = X0 rescaled so that |A X| matches |B| (:248-258), Y = normalize_rows(A X, B) (:262), M = 0
    for (size_t idx = tid; idx < (size_t)n * r; idx += NT) X[idx] = tk.X0[idx];
    for (size_t idx = tid; idx < (size_t)m * r; idx += NT) M[idx] = cmk(0.0, 0.0);
    __syncthreads();
    prod_ax();
    if (tk.sbr) {
      double v[1] = {0.0};
      for (size_t idx = tid; idx < (size_t)m * r; idx += NT) v[0] += cabs2(AX[idx]);
      block_sum<1>(v, red);
      const double s = normB / sqrt(v[0]);
      for (size_t idx = tid; idx < (size_t)n * r; idx += NT) X[idx] = cscale(X[idx], s);
      for (size_t idx = tid; idx < (size_t)m * r; idx += NT) AX[idx] = cscale(AX[idx], s);
    } else {
      for (int c = 0; c < r; ++c) {
        double v[1] = {0.0};
        for (int i = tid; i < m; i += NT) v[0] += cabs2(AX[i + (size_t)m * c]);
        block_sum<1>(v, red);
        const double s = normB / sqrt(v[0]);
        for (int k = tid; k < n; k += NT) X[k + (size_t)n * c] = cscale(X[k + (size_t)n * c], s);
        for (int i = tid; i < m; i += NT) AX[i + (size_t)m * c] = cscale(AX[i + (size_t)m * c], s);
      }
    }
    __syncthreads();
    const double isr = 1.0 / sqrt((double)r);
    for (int i = tid; i < m; i += NT) {          // normalize_rows (:383-410)
      if (tk.sbr) {
        double d2 = 0.0;
        for (int c = 0; c < r; ++c) d2 += cabs2(AX[i + (size_t)m * c]);
        double D = sqrt(d2);
        const bool z = D == 0.0;
        if (z) D = 1.0;
        for (int c = 0; c < r; ++c) Y[i + (size_t)m * c] = cscale(z ? cmk(isr, 0.0) : AX[i + (size_t)m * c], Bs[i] / D);
      } else {
        for (int c = 0; c < r; ++c) {
          cd v = AX[i + (size_t)m * c];
          double D = sqrt(cabs2(v));
          if (D == 0.0) { v = cmk(1.0, 0.0); D = 1.0; }
          Y[i + (size_t)m * c] = cscale(v, Bs[i] / D);
        }
      }
    }
    __syncthreads();
    if (prm.need_dual)       // AtY = A' Y (:263)
      gemm_tpo(n, m, r, [&](int i, int k) { return cconj(Arm[(size_t)i * n + k]); },
               [&](int i, int c) { return Y[i + (size_t)m * c]; },
               [&](int k, int c, cd v) { AtY[k + (size_t)n * c] = v; }, tile, ksred);

    double mu = 0.001, opt_obj = INFINITY, last_res = INFINITY, res_comb = 0.0;
    const double rho = 1.03;
    int iters = 0, opt_iter = -1, opt_col = -1, bumps = 0, converged = 0, have_opt = 0;
    for (int it = 1; it <= prm.maxiter; ++it) {
      const double imu = 1.0 / mu, i1mu = 1.0 / (1.0 + mu);
      // ---- X = U (Y - M/mu)  (:282, :348-351)
      for (size_t idx = tid; idx < (size_t)m * r; idx += NT) {
        const cd y = Y[idx], mm = M[idx];
        T[idx] = cmk(fma(-mm.x, imu, y.x), fma(-mm.y, imu, y.y));
      }
      __syncthreads();
      gemm_tpo(n, m, r, [&](int i, int k) { return U[k + (size_t)n * i]; },
               [&](int i, int c) { return T[i + (size_t)m * c]; },
               [&](int k, int c, cd v) { X[k + (size_t)n * c] = v; }, tile, ksred);
      prod_ax();                                                                             // :284
      // ---- Y update (:286, :359-381), M update (:291-292), objective (:296-312)
      double pYd2 = 0.0, pJM2 = 0.0, pY2 = 0.0, pAX2 = 0.0, obj2 = 0.0;
      for (int i = tid; i < m; i += NT) {
        double D = 1.0;
        bool z = false;
        double a2row = 0.0;
        if (tk.sbr) {
          double d2 = 0.0;
          for (int c = 0; c < r; ++c) {
            const cd ax = AX[i + (size_t)m * c], mm = M[i + (size_t)m * c];
            d2 += cabs2(cmk(fma(mm.x, imu, ax.x), fma(mm.y, imu, ax.y)));
            a2row += cabs2(ax);
          }
          D = sqrt(d2);
          z = D == 0.0;
          if (z) D = 1.0;
        }
        for (int c = 0; c < r; ++c) {
          const size_t q = i + (size_t)m * c;
          const cd ax = AX[q], mm = M[q], yo = Y[q];
          cd cc = cmk(fma(mm.x, imu, ax.x), fma(mm.y, imu, ax.y));
          double Dc = D;
          if (tk.sbr) {
            if (z) cc = cmk(isr, 0.0);
          } else {
            Dc = sqrt(cabs2(cc));
            if (Dc == 0.0) { cc = cmk(1.0, 0.0); Dc = 1.0; }
          }
          const cd yn = cscale(cc, (Bs[i] / Dc + mu) * i1mu);
          const cd jm = cmk(ax.x - yn.x, ax.y - yn.y), dy = cmk(yn.x - yo.x, yn.y - yo.y);
          Y[q] = yn;
          M[q] = cmk(fma(mu, jm.x, mm.x), fma(mu, jm.y, mm.y));
          dY[q] = dy;
          pYd2 += cabs2(dy); pJM2 += cabs2(jm); pY2 += cabs2(yn);
          pAX2 += cabs2(ax);
        }
        if (tk.sbr) { const double dd = sqrt(a2row) - Bs[i]; obj2 += dd * dd; }
      }
      {
        double v[5] = {pYd2, pJM2, pY2, pAX2, obj2};
        block_sum<5>(v, red);
        pYd2 = v[0]; pJM2 = v[1]; pY2 = v[2]; pAX2 = v[3]; obj2 = v[4];
      }
      double obj; int jbest = -1;
      if (tk.sbr) {
        obj = sqrt(obj2);
      } else {
        for (int c = 0; c < r; ++c) {
          double v[1] = {0.0};
          for (int i = tid; i < m; i += NT) { const double dd = sqrt(cabs2(AX[i + (size_t)m * c])) - Bs[i]; v[0] += dd * dd; }
          block_sum<1>(v, red);
          if (tid == 0) xcol[c] = sqrt(v[0]);
        }
        __syncthreads();
        obj = NAN;
        for (int c = 0; c < r; ++c) { const double oc = xcol[c]; if (oc == oc && (jbest < 0 || oc < obj)) { obj = oc; jbest = c; } }
      }
      if (obj < opt_obj) {
        opt_obj = obj; opt_iter = it; opt_col = jbest; have_opt = 1;
        if (tk.sbr) {
          if (tk.Xout) for (size_t idx = tid; idx < (size_t)n * r; idx += NT) tk.Xout[idx] = X[idx];
          if (tk.Yout) for (size_t idx = tid; idx < (size_t)m * r; idx += NT) tk.Yout[idx] = Y[idx];
        } else {
          if (tk.Xout) for (int k = tid; k < n; k += NT) tk.Xout[k] = X[k + (size_t)n * jbest];
          if (tk.Yout) for (int i = tid; i < m; i += NT) tk.Yout[i] = Y[i + (size_t)m * jbest];
        }
      }
      // ---- A'(Y - Y0) (:288, only feeds res_dual), residuals and stopping rule (:316-330)
      double nAtYd2 = 0.0, nAtY2 = 0.0;
      if (prm.need_dual) {
        __syncthreads();
        double v[2] = {0.0, 0.0};
        gemm_tpo(n, m, r, [&](int i, int k) { return cconj(Arm[(size_t)i * n + k]); },
                 [&](int i, int c) { return dY[i + (size_t)m * c]; },
                 [&](int k, int c, cd acc) {
                   const size_t q = k + (size_t)n * c;
                   cd a = AtY[q];
                   a.x += acc.x; a.y += acc.y;
                   AtY[q] = a;
                   v[0] += cabs2(acc); v[1] += cabs2(a);
                 }, tile, ksred);
        block_sum<2>(v, red);
        nAtYd2 = v[0]; nAtY2 = v[1];
      }
      const double res_prim = sqrt(pJM2), res_dual = mu * sqrt(nAtYd2);
      res_comb = sqrt(pJM2 + pYd2);
      iters = it;
      if (tk.trace != nullptr && tid == 0) tk.trace[it - 1] = res_comb;
      if (prm.need_dual) {
        const double mx = fmax(sqrt(pAX2), sqrt(pY2));
        const double th_prim = prm.tol_abs * sqrt((double)m * r) + prm.tol_rel * mx;
        const double th_dual = prm.tol_abs * sqrt((double)n * r) + prm.tol_rel * sqrt(nAtY2);
        const double th_comb = prm.tol_abs * sqrt((double)m * r * 2.0) + prm.tol_rel * sqrt(mx * mx + pY2);
        if ((res_prim < th_prim && res_dual < th_dual) || (res_comb < th_comb)) { converged = 1; break; }
      }
      if (res_comb > last_res * 0.9) { mu *= rho; ++bumps; }
      last_res = res_comb;
      __syncthreads();
    }
    __syncthreads();
    if (!have_opt) {
      const int rout = tk.sbr ? r : 1;
      if (tk.Xout) for (size_t idx = tid; idx < (size_t)n * rout; idx += NT) tk.Xout[idx] = cmk(NAN, NAN);
      if (tk.Yout) for (size_t idx = tid; idx < (size_t)m * rout; idx += NT) tk.Yout[idx] = cmk(NAN, NAN);
    }
    if (tk.state) {   // [X Z N (n x r each; Z = N = 0 here) | Y M (m x r each)]
      cd* st = tk.state;
      const size_t nr = (size_t)n * r, mr = (size_t)m * r;
      for (size_t idx = tid; idx < nr; idx += NT) { st[idx] = X[idx]; st[nr + idx] = cmk(0.0, 0.0); st[2 * nr + idx] = cmk(0.0, 0.0); }
      for (size_t idx = tid; idx < mr; idx += NT) { st[3 * nr + idx] = Y[idx]; st[3 * nr + mr + idx] = M[idx]; }
    }
    if (tk.scal && tid == 0) {
      tk.scal[SC_MU] = mu; tk.scal[SC_OPT_OBJ] = opt_obj; tk.scal[SC_ITERS] = iters;
      tk.scal[SC_OPT_ITER] = opt_iter; tk.scal[SC_OPT_COL] = opt_col; tk.scal[SC_BUMPS] = bumps;
      tk.scal[SC_CONVERGED] = converged; tk.scal[SC_RES_COMB] = res_comb; tk.scal[SC_SWEEPS] = 0;
    }
    __syncthreads();
  }
}

}  // namespace twoace
