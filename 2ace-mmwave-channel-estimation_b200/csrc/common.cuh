// Common device helpers for the twoace sm_100a kernels: complex FP64 arithmetic on double2,
// block reductions, and block-cooperative small dense linear algebra (Hermitian Jacobi eigensolver,
// SPD inverse).  All matrices are column-major complex (MATLAB layout) unless noted.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace twoace {

typedef double2 cd;

constexpr int NT = 256;       // threads per CTA of every block-cooperative kernel
constexpr int NW = NT / 32;   // warps per CTA

__device__ __forceinline__ cd cmk(double x, double y) { return make_double2(x, y); }
__device__ __forceinline__ cd cadd(cd a, cd b) { return cmk(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cd csub(cd a, cd b) { return cmk(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cd cscale(cd a, double s) { return cmk(a.x * s, a.y * s); }
__device__ __forceinline__ cd cconj(cd a) { return cmk(a.x, -a.y); }
__device__ __forceinline__ cd cmul(cd a, cd b) {
  return cmk(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ cd cmulc(cd a, cd b) {  // conj(a) * b
  return cmk(fma(a.x, b.x, a.y * b.y), fma(a.x, b.y, -a.y * b.x));
}
__device__ __forceinline__ double cabs2(cd a) { return fma(a.x, a.x, a.y * a.y); }
// acc += a * b
__device__ __forceinline__ void cfma(cd& acc, cd a, cd b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(a.y, b.x, acc.y);
}
// acc += conj(a) * b
__device__ __forceinline__ void cfmac(cd& acc, cd a, cd b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(-a.y, b.x, acc.y);
}
// acc += a * conj(b)
__device__ __forceinline__ void cfmabc(cd& acc, cd a, cd b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(a.y, b.x, acc.y);
  acc.y = fma(-a.x, b.y, acc.y);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic block-wide sum of K doubles; every thread returns with the totals in v[].
// red: shared scratch of at least K*NW doubles.  Contains two __syncthreads().
template <int K>
__device__ __forceinline__ void block_sum(double (&v)[K], double* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double s = warp_sum(v[k]);
    if (lane == 0) red[k * NW + w] = s;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < NW; ++j) s += red[k * NW + j];
    v[k] = s;
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------
// Cyclic (round-robin parallel ordering) Jacobi eigensolver for a d x d complex Hermitian matrix.
//   G   : d x d, leading dimension ldg, overwritten (diagonal -> eigenvalues, off-diagonal -> ~0)
//   V   : d x d, leading dimension ldv; if init_v it is set to I first, else it is used as the
//         accumulated basis (caller pre-rotated G = V' * G0 * V for a warm start)
//   rot : shared scratch, >= 3 * (d/2+1) doubles... laid out as cd e[], double cs[], double sn[]
// G and V may live in shared or global memory (generic pointers).  Returns #sweeps used.
// Replaces MATLAB's eig() at inferLowRankV4.m:242,407,549 (LAPACK zheev inside MATLAB); ordering
// and clamping of the eigenvalues is done by the callers exactly as the reference does.
struct JacobiScratch {
  cd* e;        // [d/2+1] e^{-i theta} of each pair
  double* cs;   // [d/2+1]
  double* sn;   // [d/2+1]
  int* flag;    // [1] rotations done in this sweep
  double* gscale;  // [1] max |diag| at entry
};

__device__ __forceinline__ void rr_pair(int d2, int round, int k, int& p, int& q) {
  // round-robin tournament on dd = 2*d2 players; round in [0, dd-1), k in [0, d2)
  const int dd = 2 * d2, md = dd - 1;
  if (k == 0) {
    p = md;
    q = round;
  } else {
    p = (round + k) % md;
    q = (round - k + md) % md;
  }
  if (p > q) { int t = p; p = q; q = t; }
}

__device__ inline int jacobi_heig(cd* G, int ldg, cd* V, int ldv, int d, bool init_v,
                                  JacobiScratch js, int max_sweeps = 40) {
  const int tid = threadIdx.x;
  if (init_v) {
    for (int idx = tid; idx < d * d; idx += NT) {
      int i = idx % d, j = idx / d;
      V[i + (size_t)ldv * j] = cmk(i == j ? 1.0 : 0.0, 0.0);
    }
  }
  __syncthreads();
  if (d < 2) return 0;
  const int d2 = (d + 1) / 2;      // pairs per round (one dummy player when d is odd)
  const int rounds = 2 * d2 - 1;
  // absolute floor: off-diagonals below 1e-18 * max|diag| cannot move any eigenvalue visibly
  if (tid == 0) {
    double g = 0.0;
    for (int i = 0; i < d; ++i) g = fmax(g, fabs(G[i + (size_t)ldg * i].x));
    *js.gscale = g;
  }
  __syncthreads();
  const double floor_abs = 1.0e-18 * (*js.gscale);
  int sweeps = 0;
  for (; sweeps < max_sweeps; ++sweeps) {
    if (tid == 0) *js.flag = 0;
    __syncthreads();
    for (int rd = 0; rd < rounds; ++rd) {
      // --- rotation parameters
      if (tid < d2) {
        int p, q;
        rr_pair(d2, rd, tid, p, q);
        double c = 1.0, s = 0.0;
        cd e = cmk(1.0, 0.0);
        if (q < d) {
          cd b = G[p + (size_t)ldg * q];
          double a = G[p + (size_t)ldg * p].x, cc = G[q + (size_t)ldg * q].x;
          double ab = sqrt(cabs2(b));
          // rotate unless the off-diagonal is negligible against the diagonal pair
          if (ab > floor_abs && ab > 1.0e-17 * sqrt(fabs(a) * fabs(cc)) && ab > 1e-300) {
            double tau = (cc - a) / (2.0 * ab);
            double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
            c = 1.0 / sqrt(1.0 + t * t);
            s = t * c;
            e = cmk(b.x / ab, -b.y / ab);  // e^{-i theta}
            *js.flag = 1;
          }
        }
        js.cs[tid] = c;
        js.sn[tid] = s;
        js.e[tid] = e;
      }
      __syncthreads();
      // --- column update of G and V:  [gp gq] <- [gp gq] * J,  J = [[c, s],[-s e, c e]]
      for (int idx = tid; idx < 2 * d2 * d; idx += NT) {
        int which = idx / (d2 * d);
        int rem = idx - which * (d2 * d);
        int k = rem / d, i = rem - k * d;
        double s = js.sn[k];
        if (s == 0.0) continue;
        int p, q;
        rr_pair(d2, rd, k, p, q);
        double c = js.cs[k];
        cd e = js.e[k];
        cd* Mx = which ? V : G;
        int ld = which ? ldv : ldg;
        cd gp = Mx[i + (size_t)ld * p], gq = Mx[i + (size_t)ld * q];
        cd eq = cmul(e, gq);
        Mx[i + (size_t)ld * p] = cmk(c * gp.x - s * eq.x, c * gp.y - s * eq.y);
        Mx[i + (size_t)ld * q] = cmk(s * gp.x + c * eq.x, s * gp.y + c * eq.y);
      }
      __syncthreads();
      // --- row update of G:  [gp; gq] <- J' * [gp; gq]
      for (int idx = tid; idx < d2 * d; idx += NT) {
        int k = idx / d, j = idx - k * d;
        double s = js.sn[k];
        if (s == 0.0) continue;
        int p, q;
        rr_pair(d2, rd, k, p, q);
        double c = js.cs[k];
        cd ec = cconj(js.e[k]);
        cd gp = G[p + (size_t)ldg * j], gq = G[q + (size_t)ldg * j];
        cd eq = cmul(ec, gq);
        cd np_ = cmk(c * gp.x - s * eq.x, c * gp.y - s * eq.y);
        cd nq_ = cmk(s * gp.x + c * eq.x, s * gp.y + c * eq.y);
        if (j == p) np_.y = 0.0;
        if (j == q) nq_.y = 0.0;
        if (j == q) np_ = cmk(0.0, 0.0);   // annihilated element (exactly)
        if (j == p) nq_ = cmk(0.0, 0.0);
        G[p + (size_t)ldg * j] = np_;
        G[q + (size_t)ldg * j] = nq_;
      }
      __syncthreads();
    }
    int f = *js.flag;
    __syncthreads();
    if (!f) break;
  }
  return sweeps;
}

// In-place inverse of a d x d Hermitian positive-definite matrix (Gauss-Jordan, no pivoting).
// S column-major with leading dimension d (generic pointer).  colk/rowk: shared scratch [d] each.
// Replaces inv(A'*A + eye(n)) at inferLowRankV4.m:221,267 (or its Woodbury core I + A*A').
__device__ inline void spd_inverse(cd* S, int d, cd* colk, cd* rowk) {
  const int tid = threadIdx.x;
  for (int k = 0; k < d; ++k) {
    // save pivot row and column
    for (int i = tid; i < d; i += NT) {
      colk[i] = S[i + (size_t)d * k];
      rowk[i] = S[k + (size_t)d * i];
    }
    __syncthreads();
    const cd pv = colk[k];
    // 1/pv (pivot is real positive for an HPD matrix up to rounding; keep it complex for safety)
    const double den = cabs2(pv);
    const cd ip = cmk(pv.x / den, -pv.y / den);
    for (int idx = tid; idx < d * d; idx += NT) {
      int i = idx % d, j = idx / d;
      cd v;
      if (i == k && j == k) {
        v = ip;
      } else if (i == k) {
        v = cmul(rowk[j], ip);
      } else if (j == k) {
        cd t = cmul(colk[i], ip);
        v = cmk(-t.x, -t.y);
      } else {
        cd f = cmul(colk[i], ip);
        cd cur = S[idx];
        cd rj = rowk[j];
        v = cmk(cur.x - (f.x * rj.x - f.y * rj.y), cur.y - (f.x * rj.y + f.y * rj.x));
      }
      S[idx] = v;
    }
    __syncthreads();
  }
}

}  // namespace twoace
