// Evaluation metrics of one recovered channel against the ground truth, one CTA per instance
// (Numerical_Simulation/src/evaluate_plot_results/Evaluation_H.m:81-115): they are the
// payload of the statistics reduce of the multi-GPU path.
//   out[0] MSE_H      = || x_gt - (x'x_gt)/(x'x) x ||^2 / ||x_gt||^2                          (:87-89)
//   out[1] gain_ana   = | w' H f |, w / f = leading singular vectors of the estimate, phases quantised to
//                       Phase_Bit bits, constant modulus 1/sqrt(length)           (:94-102, Quantize_PS.m)
//   out[2] gain_dig   = | u' H v | with the unquantised unit vectors                           (:99-103)
//   out[3] proj_error = the MSE_H-type residual (not squared) between the rank-one approximations of the
//                       true and the estimated channel                                        (:106-115)
// svd() is replaced by the Hermitian eigenproblem of H'H (Jacobi) + u = H v / sigma.  A singular pair is only
// defined up to a common phase (u, v) -> e^{i phi} (u, v); gain_dig and proj_error do not depend on it, the
// quantised gain_ana does (LAPACK's phase is an implementation artefact of MATLAB's svd).  Canonical choice
// here and in oracle/metrics.py: the entry of v with the largest modulus (first on ties) is real positive.
#pragma once
#include "common.cuh"

namespace twoace {

constexpr int MET_WORDS = 4;
constexpr int MET_DMAX = 32;   // antennas per side

struct MetSmem {
  cd *He, *Ht, *G, *V, *ue, *ve, *ut, *vt, *tmp;
  double* red;
  JacobiScratch js;
};

__host__ __device__ inline size_t met_smem_bytes() {
  size_t b = 0;
  b += 4 * (size_t)MET_DMAX * MET_DMAX * sizeof(cd);          // He Ht G V
  b += 5 * (size_t)MET_DMAX * sizeof(cd);                     // ue ve ut vt tmp
  b += (size_t)(MET_DMAX / 2 + 2) * (sizeof(cd) + 2 * sizeof(double));
  b += 8 * NW * sizeof(double) + 64;
  return b + 64;
}

// leading singular triplet of the nr x nt matrix H (shared, column-major): sigma, u [nr], v [nt]
__device__ inline double leading_triplet(const cd* H, int nr, int nt, MetSmem& sm, cd* u, cd* v) {
  const int tid = threadIdx.x;
  for (int e = tid; e < nt * nt; e += NT) {
    const int a = e % nt, b = e / nt;
    cd acc = cmk(0.0, 0.0);
    for (int r = 0; r < nr; ++r) cfmac(acc, H[r + nr * a], H[r + nr * b]);   // conj(H[r,a]) * H[r,b]
    if (a == b) acc.y = 0.0;
    sm.G[e] = acc;
  }
  __syncthreads();
  jacobi_heig(sm.G, nt, sm.V, nt, nt, true, sm.js);
  __syncthreads();
  int kb = 0;
  double lb = -INFINITY;
  for (int k = 0; k < nt; ++k) {
    const double l = sm.G[k + nt * k].x;
    if (l > lb) { lb = l; kb = k; }
  }
  const double sigma = sqrt(fmax(lb, 0.0));
  // canonical phase: largest-modulus entry of v real positive
  int jb = 0;
  double mb = -1.0;
  for (int j = 0; j < nt; ++j) {
    const double a2 = cabs2(sm.V[j + nt * kb]);
    if (a2 > mb) { mb = a2; jb = j; }
  }
  const cd pv = sm.V[jb + nt * kb];
  const double pm = sqrt(cabs2(pv));
  const cd ph = pm > 0.0 ? cmk(pv.x / pm, -pv.y / pm) : cmk(1.0, 0.0);       // conj(phase)
  __syncthreads();
  for (int j = tid; j < nt; j += NT) v[j] = cmul(sm.V[j + nt * kb], ph);
  __syncthreads();
  for (int r = tid; r < nr; r += NT) {
    cd acc = cmk(0.0, 0.0);
    for (int j = 0; j < nt; ++j) cfma(acc, H[r + nr * j], v[j]);
    u[r] = sigma > 0.0 ? cscale(acc, 1.0 / sigma) : cmk(r == 0 ? 1.0 : 0.0, 0.0);
  }
  __syncthreads();
  return sigma;
}

// Quantize_PS.m: nearest of phi = -pi : 2 pi / 2^bits : pi (first minimum), modulus 1/sqrt(len)
__device__ __forceinline__ cd quantize_ps(cd x, int bits, int len) {
  const double PI = 3.141592653589793238462643383279502884;
  const int nps = 1 << bits;
  const double ang = atan2(x.y, x.x), stepq = 2.0 * PI / nps;
  int best = 0;
  double bd = INFINITY;
  for (int q = 0; q <= nps; ++q) {
    const double dq = fabs(ang - (-PI + q * stepq));
    if (dq < bd) { bd = dq; best = q; }
  }
  const double phi = -PI + best * stepq, s = 1.0 / sqrt((double)len);
  return cmk(cos(phi) * s, sin(phi) * s);
}

__global__ void __launch_bounds__(NT) metrics_kernel(const cd* __restrict__ Xest, const cd* __restrict__ Xtrue,
                                                     int nb, int nt, int nr, int phase_bit, double* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char met_raw[];
  MetSmem sm;
  {
    unsigned char* p = met_raw;
    sm.He = (cd*)p; p += MET_DMAX * MET_DMAX * sizeof(cd);
    sm.Ht = (cd*)p; p += MET_DMAX * MET_DMAX * sizeof(cd);
    sm.G = (cd*)p;  p += MET_DMAX * MET_DMAX * sizeof(cd);
    sm.V = (cd*)p;  p += MET_DMAX * MET_DMAX * sizeof(cd);
    sm.ue = (cd*)p; p += MET_DMAX * sizeof(cd);
    sm.ve = (cd*)p; p += MET_DMAX * sizeof(cd);
    sm.ut = (cd*)p; p += MET_DMAX * sizeof(cd);
    sm.vt = (cd*)p; p += MET_DMAX * sizeof(cd);
    sm.tmp = (cd*)p; p += MET_DMAX * sizeof(cd);
    const int h = MET_DMAX / 2 + 2;
    sm.js.e = (cd*)p; p += h * sizeof(cd);
    sm.js.cs = (double*)p; p += h * sizeof(double);
    sm.js.sn = (double*)p; p += h * sizeof(double);
    sm.red = (double*)p; p += 8 * NW * sizeof(double);
    sm.js.gscale = (double*)p; p += 8;
    sm.js.flag = (int*)p;
  }
  const int tid = threadIdx.x, n = nt * nr;
  for (int b = blockIdx.x; b < nb; b += gridDim.x) {
    __syncthreads();
    for (int e = tid; e < n; e += NT) { sm.He[e] = Xest[(size_t)b * n + e]; sm.Ht[e] = Xtrue[(size_t)b * n + e]; }
    __syncthreads();
    // ---- MSE_H (:87-89)
    double v5[5] = {0.0, 0.0, 0.0, 0.0, 0.0};   // x'x, Re/Im x'x_gt, |x_gt|^2, NaN flag
    for (int e = tid; e < n; e += NT) {
      const cd x = sm.He[e], g = sm.Ht[e];
      v5[0] += cabs2(x);
      v5[1] += x.x * g.x + x.y * g.y;
      v5[2] += x.x * g.y - x.y * g.x;
      v5[3] += cabs2(g);
      if (x.x != x.x || x.y != x.y) v5[4] += 1.0;
    }
    block_sum<5>(v5, sm.red);
    const bool bad = v5[4] > 0.0 || !(v5[0] > 0.0);
    double mse = NAN, gain_ana = NAN, gain_dig = NAN, proj = NAN;
    if (!bad) {   // (uniform)
      const cd alpha = cmk(v5[1] / v5[0], v5[2] / v5[0]);
      double r1[1] = {0.0};
      for (int e = tid; e < n; e += NT) r1[0] += cabs2(csub(sm.Ht[e], cmul(alpha, sm.He[e])));
      block_sum<1>(r1, sm.red);
      mse = r1[0] / v5[3];
      // ---- leading singular triplets (:94-96, :106-111)
      const double se = leading_triplet(sm.He, nr, nt, sm, sm.ue, sm.ve);
      const double st = leading_triplet(sm.Ht, nr, nt, sm, sm.ut, sm.vt);
      // ---- beamforming gains (:97-103): t = H f, gain = |w' t|
      double g4[4] = {0.0, 0.0, 0.0, 0.0};
      for (int r = tid; r < nr; r += NT) {
        cd ta = cmk(0.0, 0.0), td = cmk(0.0, 0.0);
        for (int j = 0; j < nt; ++j) {
          const cd h = sm.Ht[r + nr * j];
          cfma(ta, h, quantize_ps(sm.ve[j], phase_bit, nt));
          cfma(td, h, sm.ve[j]);
        }
        const cd wa = quantize_ps(sm.ue[r], phase_bit, nr), wd = sm.ue[r];
        const cd pa = cmulc(wa, ta), pd = cmulc(wd, td);      // conj(w) * t
        g4[0] += pa.x; g4[1] += pa.y; g4[2] += pd.x; g4[3] += pd.y;
      }
      block_sum<4>(g4, sm.red);
      gain_ana = sqrt(g4[0] * g4[0] + g4[1] * g4[1]);
      gain_dig = sqrt(g4[2] * g4[2] + g4[3] * g4[3]);
      // ---- 1-D projection error (:106-115) between the rank-one approximations sigma u v'
      double p4[4] = {0.0, 0.0, 0.0, 0.0};   // x'x, Re/Im x'x_gt, |x_gt|^2
      for (int e = tid; e < n; e += NT) {
        const int r = e % nr, c = e / nr;
        const cd x = cscale(cmul(sm.ue[r], cconj(sm.ve[c])), se), g = cscale(cmul(sm.ut[r], cconj(sm.vt[c])), st);
        p4[0] += cabs2(x);
        p4[1] += x.x * g.x + x.y * g.y;
        p4[2] += x.x * g.y - x.y * g.x;
        p4[3] += cabs2(g);
      }
      block_sum<4>(p4, sm.red);
      const cd beta = cmk(p4[1] / p4[0], p4[2] / p4[0]);
      double r2[1] = {0.0};
      for (int e = tid; e < n; e += NT) {
        const int r = e % nr, c = e / nr;
        const cd x = cscale(cmul(sm.ue[r], cconj(sm.ve[c])), se), g = cscale(cmul(sm.ut[r], cconj(sm.vt[c])), st);
        r2[0] += cabs2(csub(g, cmul(beta, x)));
      }
      block_sum<1>(r2, sm.red);
      proj = sqrt(r2[0]) / sqrt(p4[3]);
    }
    if (tid == 0) {
      out[(size_t)b * MET_WORDS + 0] = mse;
      out[(size_t)b * MET_WORDS + 1] = gain_ana;
      out[(size_t)b * MET_WORDS + 2] = gain_dig;
      out[(size_t)b * MET_WORDS + 3] = proj;
    }
  }
}

}  // namespace twoace

// ------------------------------------------------------------------------------------------------------------------
// AoD / AoA estimation error (Numerical_Simulation/src/evaluate_plot_results/Evaluation_Recovery.m:85-146) of an
// H-domain estimate: its angular spectrum z = vec(A_Rx' H A_Tx) on the virtual-angle dictionary of
// generate_channel/Sparse_Channel_Formulation.m:83-103, restricted to the searching area (:120-152, AoD index outer),
// plays the role of `recoveredSig`; the L largest entries give the estimated angles.  One CTA per instance.
//   out[0..5] = AoD_Err_to_True, AoA_Err_to_True, AoDA_Err, AoD_Err_to_True_Quantized, AoA_Err_to_True_Quantized,
//               AoDA_Err_Quantized   (degrees; NaN for a non-finite estimate)
// Kept quirk: :133 `AoA_True(order_True) = AoA_True(order_True)` is a no-op, the true AoAs stay in path order.
namespace twoace {

constexpr int ANG_WORDS = 6;
constexpr int ANG_LMAX = 32;

struct AngDims {
  int nt, nr, L, nqt, nqr;
  int u0, u1, v0, v1;        // inclusive grid index ranges covering the searching area
  double kph;                // 2 pi d / lambda
  size_t ws_stride;          // doubles per CTA: 2 nv nt (W) + nv nu (|z|^2)
};

__global__ void __launch_bounds__(NT) angle_metrics_kernel(const cd* __restrict__ Xest, const double* __restrict__ ang_true,
                                                          int nb, AngDims dm, double* __restrict__ out, double* wsbase) {
  __shared__ double s_val[NW];
  __shared__ int s_idx[NW];
  __shared__ int s_top[ANG_LMAX];
  __shared__ int s_bad;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nt = dm.nt, nr = dm.nr, L = dm.L, n = nt * nr;
  const int nu = dm.u1 - dm.u0 + 1, nv = dm.v1 - dm.v0 + 1;
  cd* W = (cd*)(wsbase + (size_t)blockIdx.x * dm.ws_stride);
  double* mag = (double*)(W + (size_t)nv * nt);
  const double PI = 3.14159265358979323846;
  for (int b = blockIdx.x; b < nb; b += gridDim.x) {
    const cd* H = Xest + (size_t)b * n;            // H[kr + nr kt]
    if (tid == 0) s_bad = 0;
    __syncthreads();
    for (int idx = tid; idx < nv * nt; idx += NT) {          // W = A_Rx(:, v)' H
      const int vi = idx % nv, kt = idx / nv;
      const double phi = dm.kph * (-1.0 + 2.0 * (double)(dm.v0 + vi) / (double)dm.nqr);
      cd acc = cmk(0.0, 0.0);
      for (int kr = 0; kr < nr; ++kr) {
        double s, c;
        sincos(phi * kr, &s, &c);
        cfma(acc, cmk(c, s), H[kr + nr * kt]);               // conj(exp(-j phi kr))
      }
      W[idx] = cscale(acc, 1.0 / sqrt((double)nr));
    }
    __syncthreads();
    for (int idx = tid; idx < nv * nu; idx += NT) {          // z = W A_Tx(:, u), entry ui * nv + vi
      const int vi = idx % nv, ui = idx / nv;
      const double phi = dm.kph * (-1.0 + 2.0 * (double)(dm.u0 + ui) / (double)dm.nqt);
      cd acc = cmk(0.0, 0.0);
      for (int kt = 0; kt < nt; ++kt) {
        double s, c;
        sincos(phi * kt, &s, &c);
        cfma(acc, cmk(c, -s), W[vi + nv * kt]);
      }
      const double m2 = cabs2(acc) / (double)nt;
      if (!(m2 == m2) || isinf(m2)) s_bad = 1;
      mag[idx] = m2;
    }
    __syncthreads();
    if (s_bad) {
      if (tid < ANG_WORDS) out[(size_t)b * ANG_WORDS + tid] = NAN;
      __syncthreads();
      continue;
    }
    for (int l = 0; l < L; ++l) {                            // the L largest entries, first index on ties (:86-87)
      double bv = -1.0;
      int bi = 0x7fffffff;
      for (int idx = tid; idx < nv * nu; idx += NT) {
        const double v = mag[idx];
        if (v > bv) { bv = v; bi = idx; }
      }
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      if (lane == 0) { s_val[warp] = bv; s_idx[warp] = bi; }
      __syncthreads();
      if (tid == 0) {
        for (int w = 1; w < NW; ++w)
          if (s_val[w] > bv || (s_val[w] == bv && s_idx[w] < bi)) { bv = s_val[w]; bi = s_idx[w]; }
        s_top[l] = bi;
        mag[bi] = -2.0;
      }
      __syncthreads();
    }
    if (tid == 0) {
      int ind[ANG_LMAX];
      double aod_e[ANG_LMAX], aoa_e[ANG_LMAX], aod_t[ANG_LMAX], aoa_t[ANG_LMAX], aod_q[ANG_LMAX], aoa_q[ANG_LMAX];
      for (int l = 0; l < L; ++l) ind[l] = s_top[l];
      for (int i = 1; i < L; ++i) {                          // :88 sort ascending
        const int key = ind[i];
        int j = i - 1;
        while (j >= 0 && ind[j] > key) { ind[j + 1] = ind[j]; --j; }
        ind[j + 1] = key;
      }
      const double r2d = 180.0 / PI;
      for (int l = 0; l < L; ++l) {
        const int ui = ind[l] / nv, vi = ind[l] % nv;
        aod_e[l] = asin(-1.0 + 2.0 * (double)(dm.u0 + ui) / (double)dm.nqt) * r2d;       // :104-113
        aoa_e[l] = asin(-1.0 + 2.0 * (double)(dm.v0 + vi) / (double)dm.nqr) * r2d;
        const double td = ang_true[(size_t)b * 2 * L + l], ta = ang_true[(size_t)b * 2 * L + L + l];
        aod_t[l] = td; aoa_t[l] = ta;
        // nearest grid point of the true virtual angle, first minimum (Sparse_Channel_Formulation.m:105-113)
        const double sd = sin(td * PI / 180.0), sa = sin(ta * PI / 180.0);
        int pd = 0, pa = 0;
        double ed = INFINITY, ea = INFINITY;
        for (int q = 0; q < dm.nqt; ++q) { const double e = fabs(dm.kph * (-1.0 + 2.0 * q / (double)dm.nqt) - dm.kph * sd); if (e < ed) { ed = e; pd = q; } }
        for (int q = 0; q < dm.nqr; ++q) { const double e = fabs(dm.kph * (-1.0 + 2.0 * q / (double)dm.nqr) - dm.kph * sa); if (e < ea) { ea = e; pa = q; } }
        aod_q[l] = asin(-1.0 + 2.0 * pd / (double)dm.nqt) * r2d;                         // :115-123
        aoa_q[l] = asin(-1.0 + 2.0 * pa / (double)dm.nqr) * r2d;
      }
      // :132-137 stable descending sorts: true AoD (carrying the quantised pairs, NOT the true AoA), estimated AoD
      for (int i = 1; i < L; ++i) {
        const double kd = aod_t[i], kq = aod_q[i], ka = aoa_q[i];
        int j = i - 1;
        while (j >= 0 && aod_t[j] < kd) { aod_t[j + 1] = aod_t[j]; aod_q[j + 1] = aod_q[j]; aoa_q[j + 1] = aoa_q[j]; --j; }
        aod_t[j + 1] = kd; aod_q[j + 1] = kq; aoa_q[j + 1] = ka;
      }
      for (int i = 1; i < L; ++i) {
        const double kd = aod_e[i], ka = aoa_e[i];
        int j = i - 1;
        while (j >= 0 && aod_e[j] < kd) { aod_e[j + 1] = aod_e[j]; aoa_e[j + 1] = aoa_e[j]; --j; }
        aod_e[j + 1] = kd; aoa_e[j + 1] = ka;
      }
      double e_d = 0.0, e_a = 0.0, e_dq = 0.0, e_aq = 0.0;
      for (int l = 0; l < L; ++l) {
        e_d += fabs(aod_e[l] - aod_t[l]); e_a += fabs(aoa_e[l] - aoa_t[l]);
        e_dq += fabs(aod_e[l] - aod_q[l]); e_aq += fabs(aoa_e[l] - aoa_q[l]);
      }
      e_d /= L; e_a /= L; e_dq /= L; e_aq /= L;
      double* o = out + (size_t)b * ANG_WORDS;
      o[0] = e_d; o[1] = e_a;
      o[2] = nt == 1 ? e_a : (nr == 1 ? e_d : 0.5 * (e_d + e_a));                         // :144-151
      o[3] = e_dq; o[4] = e_aq; o[5] = 0.5 * (e_dq + e_aq);
    }
    __syncthreads();
  }
}

}  // namespace twoace
