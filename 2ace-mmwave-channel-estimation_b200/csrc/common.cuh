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

// Rotation J = [[c, s],[-s e, c e]] (e = e^{-i theta}) that annihilates the (p,q) entry of the Hermitian
// 2x2 block [[a, b],[conj b, cc]].  Returns false (identity) when |b| is negligible.
__device__ __forceinline__ bool jacobi_rot(double a, double cc, cd b, double floor2, double& c, double& s, cd& e) {
  const double ab2 = cabs2(b);
  c = 1.0; s = 0.0; e = cmk(1.0, 0.0);
  if (!(ab2 > floor2) || !(ab2 > 1.0e-34 * fabs(a * cc)) || !(ab2 > 1e-300)) return false;
  // rsqrt / rcp (1 ulp) keep the dependent chain short; J stays unitary to ~1 ulp
  const double rab = rsqrt(ab2);
  const double tau = 0.5 * (cc - a) * rab;
  const double x1 = fma(tau, tau, 1.0);
  const double w = x1 * rsqrt(x1);
  const double t = copysign(__drcp_rn(fabs(tau) + w), tau);
  c = rsqrt(fma(t, t, 1.0));
  s = t * c;
  e = cmk(b.x * rab, -b.y * rab);
  return true;
}

__device__ inline int jacobi_heig(cd* G, int ldg, cd* V, int ldv, int d, bool init_v,
                                  JacobiScratch js, int max_sweeps = 40) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (init_v) {
    for (int j = warp; j < d; j += NW)
      for (int i = lane; i < d; i += 32) V[i + (size_t)ldv * j] = cmk(i == j ? 1.0 : 0.0, 0.0);
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
  const double floor2 = floor_abs * floor_abs;
  // the pair of (round, k) is packed as (p << 16) | q in the low bits of js.cs-adjacent storage: we keep
  // it in registers of the d2 parameter threads and publish it through js.sn/js.cs companions below
  int sweeps = 0;
  for (; sweeps < max_sweeps; ++sweeps) {
    if (tid == 0) *js.flag = 0;
    __syncthreads();
    for (int rd = 0; rd < rounds; ++rd) {
      // --- rotation parameters (one thread per pair); the pair indices ride along in e.y of a 2nd slot
      for (int k = tid; k < d2; k += NT) {
        int p, q;
        rr_pair(d2, rd, k, p, q);
        double c = 1.0, s = 0.0;
        cd e = cmk(1.0, 0.0);
        if (q < d) {
          const cd b = G[p + (size_t)ldg * q];
          const double a = G[p + (size_t)ldg * p].x, cc = G[q + (size_t)ldg * q].x;
          if (jacobi_rot(a, cc, b, floor2, c, s, e)) *js.flag = 1;
        }
        js.cs[k] = c;
        js.sn[k] = s;
        js.e[k] = e;
      }
      __syncthreads();
      // --- column update of G and V:  [gp gq] <- [gp gq] * J   (warp per pair, lanes over rows)
      for (int k = warp; k < d2; k += NW) {
        const double s = js.sn[k];
        if (s == 0.0) continue;
        int p, q;
        rr_pair(d2, rd, k, p, q);
        const double c = js.cs[k];
        const cd e = js.e[k];
        for (int i = lane; i < 2 * d; i += 32) {
          cd* Mx = (i < d) ? G : V;
          const int ld = (i < d) ? ldg : ldv;
          const int ii = (i < d) ? i : i - d;
          const cd gp = Mx[ii + (size_t)ld * p], gq = Mx[ii + (size_t)ld * q];
          const cd eq = cmul(e, gq);
          Mx[ii + (size_t)ld * p] = cmk(c * gp.x - s * eq.x, c * gp.y - s * eq.y);
          Mx[ii + (size_t)ld * q] = cmk(s * gp.x + c * eq.x, s * gp.y + c * eq.y);
        }
      }
      __syncthreads();
      // --- row update of G:  [gp; gq] <- J' * [gp; gq]
      for (int k = warp; k < d2; k += NW) {
        const double s = js.sn[k];
        if (s == 0.0) continue;
        int p, q;
        rr_pair(d2, rd, k, p, q);
        const double c = js.cs[k];
        const cd ec = cconj(js.e[k]);
        for (int j = lane; j < d; j += 32) {
          const cd gp = G[p + (size_t)ldg * j], gq = G[q + (size_t)ldg * j];
          const cd eq = cmul(ec, gq);
          cd np_ = cmk(c * gp.x - s * eq.x, c * gp.y - s * eq.y);
          cd nq_ = cmk(s * gp.x + c * eq.x, s * gp.y + c * eq.y);
          if (j == p) np_.y = 0.0;
          if (j == q) nq_.y = 0.0;
          if (j == q) np_ = cmk(0.0, 0.0);   // annihilated element (exactly)
          if (j == p) nq_ = cmk(0.0, 0.0);
          G[p + (size_t)ldg * j] = np_;
          G[q + (size_t)ldg * j] = nq_;
        }
      }
      __syncthreads();
    }
    int f = *js.flag;
    __syncthreads();
    if (!f) break;
  }
  return sweeps;
}

// ------------------------------------------------------------------------------------------
// 16 x 16 specialisation for the ArgMinZ eigenproblem of the shared-memory kernel (exactly NT = 256
// threads).  Two-sided Jacobi, double-buffered (column update G -> H, row update H -> G) so that each
// round needs two barriers; every thread derives the rotation of its own pair redundantly from G, the
// pair table `pairs` ([15][8][2] bytes) is precomputed.  V is updated in place; with init_v == false the
// caller supplies V and the already rotated G = V' G0 V (warm start).
__device__ __forceinline__ void jacobi16_pairs(unsigned char* pairs) {
  for (int idx = threadIdx.x; idx < 15 * 8; idx += NT) {
    int p, q;
    rr_pair(8, idx >> 3, idx & 7, p, q);
    pairs[2 * idx] = (unsigned char)p;
    pairs[2 * idx + 1] = (unsigned char)q;
  }
}

__device__ inline int jacobi16(cd* G, cd* H, cd* V, const unsigned char* pairs, double* prm, bool init_v,
                               int max_sweeps = 30) {
  const int tid = threadIdx.x;
  const int which = tid >> 7, k = (tid >> 4) & 7, i = tid & 15;
  const int leader = tid & 16;   // lane (within the warp) of the first thread of this half-warp
  if (init_v) V[tid] = cmk((i == (tid >> 4)) ? 1.0 : 0.0, 0.0);
  double g = 0.0;
#pragma unroll
  for (int q = 0; q < 16; ++q) g = fmax(g, fabs(G[17 * q].x));
  const double floor_abs = 1.0e-18 * g;
  const double floor2 = floor_abs * floor_abs;
  __syncthreads();
  cd* Mx = which ? V : G;
  cd* Mo = which ? V : H;
  int sweeps = 0;
  for (; sweeps < max_sweeps;) {
    double smax = 0.0;
    for (int rd = 0; rd < 15; ++rd) {
      const int p = pairs[2 * (rd * 8 + k)], q = pairs[2 * (rd * 8 + k) + 1];
      double c = 1.0, s = 0.0;
      cd e = cmk(1.0, 0.0);
      if (i == 0) {   // one thread per half-warp derives the rotation, the other 15 receive it
        jacobi_rot(G[17 * p].x, G[17 * q].x, G[p + 16 * q], floor2, c, s, e);
        smax = fmax(smax, fabs(s));
      }
      c = __shfl_sync(0xffffffffu, c, leader);
      s = __shfl_sync(0xffffffffu, s, leader);
      e.x = __shfl_sync(0xffffffffu, e.x, leader);
      e.y = __shfl_sync(0xffffffffu, e.y, leader);
      {
        const cd gp = Mx[i + 16 * p], gq = Mx[i + 16 * q];
        const cd eq = cmul(e, gq);
        Mo[i + 16 * p] = cmk(c * gp.x - s * eq.x, c * gp.y - s * eq.y);
        Mo[i + 16 * q] = cmk(s * gp.x + c * eq.x, s * gp.y + c * eq.y);
      }
      if (tid < 128 && i == 0) { prm[4 * k] = c; prm[4 * k + 1] = s; prm[4 * k + 2] = e.x; prm[4 * k + 3] = e.y; }
      __syncthreads();
      if (tid < 128) {   // row update H -> G, item (pair k, column j = i)
        const double c2 = prm[4 * k], s2 = prm[4 * k + 1];
        const cd ec = cmk(prm[4 * k + 2], -prm[4 * k + 3]);
        const cd hp = H[p + 16 * i], hq = H[q + 16 * i];
        const cd eq = cmul(ec, hq);
        cd np_ = cmk(c2 * hp.x - s2 * eq.x, c2 * hp.y - s2 * eq.y);
        cd nq_ = cmk(s2 * hp.x + c2 * eq.x, s2 * hp.y + c2 * eq.y);
        if (s2 != 0.0) {
          if (i == p) np_.y = 0.0;
          if (i == q) nq_.y = 0.0;
          if (i == q) np_ = cmk(0.0, 0.0);
          if (i == p) nq_ = cmk(0.0, 0.0);
        }
        G[p + 16 * i] = np_;
        G[q + 16 * i] = nq_;
      }
      __syncthreads();
    }
    ++sweeps;
    // quadratic convergence: a sweep whose largest rotation had |sin| <= 1e-8 leaves off-diagonals at the
    // 1e-16 level, so no verification sweep is needed
    if (!__syncthreads_or(smax > 1.0e-8)) break;
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
