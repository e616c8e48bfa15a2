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

// Rotation J = [[c, sg],[-conj(sg), c]] annihilating the (p,q) entry b of [[a, b],[conj b, cc]].
// A dependent FP64 op costs ~23 cycles on B200 (FP32: 4), so the ANGLE is derived in FP32 (relative
// accuracy ~1e-7: the rotation then leaves a residual of 1e-7 |b|, which costs no extra sweep because
// Jacobi's quadratic convergence only beats a 1e-7 contraction in its very last sweep) and the pair
// (c, sg) is renormalised in FP64 so that J is unitary to 1 ulp: n2 = c^2 + |sg|^2 = 1 + eps, |eps| ~ 1e-7,
// 1/sqrt(n2) = 1 - eps/2 + 3 eps^2/8 (eps^3 ~ 1e-21).
__device__ __forceinline__ void jacobi_rot_sg(double a, double cc, cd b, double floor2, double inv_g, double& c,
                                              cd& sg, double skip_below = 0.0) {
  const double ab2 = cabs2(b);
  c = 1.0;
  sg = cmk(0.0, 0.0);
  if (!(ab2 > floor2) || !(ab2 > 1.0e-34 * fabs(a * cc)) || !(ab2 > 1e-300)) return;
  // optional: leave a pair alone when its whole 2x2 block lies below `skip_below` (|a|, |cc|, |b| < skip):
  // the caller does not need that part of the spectrum resolved (see the nuclear ArgMinZ)
  if (fmax(fabs(a), fabs(cc)) < skip_below && ab2 < skip_below * skip_below) return;
  // FP32 angle on inputs normalised by g = max|diag| (inv_g = 1/g): |b|^2/g^2 lies in (1e-36, ~1] and
  // |cc - a|/g <= 2, safely inside the float range; tau = (cc - a) / (2 |b|)
  const float bxn = (float)(b.x * inv_g), byn = (float)(b.y * inv_g), dn = (float)((cc - a) * inv_g);
  const float rabf = rsqrtf(fmaxf(fmaf(bxn, bxn, byn * byn), 1e-37f));
  const float tauf = fminf(fmaxf(0.5f * dn * rabf, -1e18f), 1e18f);
  const float bxf = bxn * rabf, byf = byn * rabf;
  const float x1f = fmaf(tauf, tauf, 1.0f);
  const float wf = x1f * rsqrtf(x1f);
  const float tf = copysignf(__fdividef(1.0f, fabsf(tauf) + wf), tauf);
  const float cf = rsqrtf(fmaf(tf, tf, 1.0f));
  const float sf = tf * cf;
  const double c0 = (double)cf, sx = (double)(sf * bxf), sy = (double)(sf * byf);
  const double eps = fma(c0, c0, fma(sx, sx, sy * sy)) - 1.0;
  const double rn = fma(eps, fma(eps, 0.375, -0.5), 1.0);
  c = c0 * rn;
  sg = cmk(sx * rn, sy * rn);
}

// D x D two-sided Jacobi (D even, D <= 20), one barrier per round, one thread per matrix element (two when
// D*D > NT).  Round = D/2 disjoint pairs (round-robin table).  Thread (i, j) produces element (i, j) of
// J' G J from the read-only previous G (ping-pong between Ga and Gb):
//   n_ij = ca (cb g_ij + bp g_i,pj) + ap (cb g_pi,j + bp g_pi,pj)
// with pi / pj the partners of i / j in this round and (ca, ap) / (cb, bp) the entries of J' / J that mix
// them.  The D/2 rotations are derived by the first lanes of every warp (FP32 angle, FP64 renormalisation)
// and shuffled to the lanes that need them; the first D*D/2 threads also rotate the columns of V in place.
//   tab   [D-1][D]      bytes: pair index (bits 0-3) | is-larger-member flag (bit 7) ; partner in tab2
//   tab2  [D-1][D]      bytes: partner index
//   pairs [D-1][D/2][2] bytes: (p, q), p < q
// Result: eigenvalues on the diagonal of Ga, eigenvectors in V.  With init_v == false the caller supplies
// V and Ga = V' G0 V (warm start).  Requires exactly NT = 256 threads.
template <int D>
struct JacobiTab {
  static constexpr int ROUNDS = D - 1, HALF = D / 2;
  static constexpr int PAIRS_BYTES = ROUNDS * HALF * 2, TAB_BYTES = ROUNDS * D;
  static constexpr int TABLES = (PAIRS_BYTES + 2 * TAB_BYTES + 15) / 16 * 16;
  static constexpr int PRM_BYTES = 2 * HALF * 4 * 8;   // rotation parameters (c, sg.x, sg.y, |sg|^2), double-buffered
  static constexpr int BYTES = TABLES + PRM_BYTES;
};

template <int D>
__device__ __forceinline__ void jacobi_tables(unsigned char* mem) {
  unsigned char* pairs = mem;
  unsigned char* tab = mem + JacobiTab<D>::PAIRS_BYTES;
  unsigned char* tab2 = tab + JacobiTab<D>::TAB_BYTES;
  constexpr int H = D / 2;
  for (int idx = threadIdx.x; idx < (D - 1) * H; idx += NT) {
    const int rd = idx / H, k = idx - rd * H;
    int p, q;
    rr_pair(H, rd, k, p, q);
    pairs[2 * idx] = (unsigned char)p;
    pairs[2 * idx + 1] = (unsigned char)q;
    tab[D * rd + p] = (unsigned char)k;
    tab[D * rd + q] = (unsigned char)(k | 128);
    tab2[D * rd + p] = (unsigned char)q;
    tab2[D * rd + q] = (unsigned char)p;
  }
}

// Tables of the "cross" ordering: H rounds, round r pairs index k of the first half with H + (k + r) % H of the
// second half -- every (first half, second half) pair exactly once, no pair inside a half.  Used by the block
// Jacobi for block pairs whose diagonal blocks were already swept in this sweep.
template <int D>
__device__ __forceinline__ void jacobi_tables_cross(unsigned char* mem) {
  unsigned char* pairs = mem;
  unsigned char* tab = mem + JacobiTab<D>::PAIRS_BYTES;
  unsigned char* tab2 = tab + JacobiTab<D>::TAB_BYTES;
  constexpr int H = D / 2;
  for (int idx = threadIdx.x; idx < H * H; idx += NT) {
    const int rd = idx / H, k = idx - rd * H;
    const int p = k, q = H + (k + rd) % H;
    pairs[2 * idx] = (unsigned char)p;
    pairs[2 * idx + 1] = (unsigned char)q;
    tab[D * rd + p] = (unsigned char)k;
    tab[D * rd + q] = (unsigned char)(k | 128);
    tab2[D * rd + p] = (unsigned char)q;
    tab2[D * rd + q] = (unsigned char)p;
  }
}

template <int D>
__device__ inline int jacobi_small(cd* Ga, cd* Gb, cd* V, const unsigned char* mem, bool init_v,
                                   int max_sweeps = 30, int* any_rotation = nullptr, double skip_below = 0.0,
                                   double floor_rel = 1.0e-18, double stop_sin2 = 1.0e-16, int nrounds = D - 1) {
  static_assert(D % 2 == 0 && D <= 32 && D * D <= 4 * NT, "unsupported dimension");
  constexpr int H = D / 2, E = D * D, EPT = (E + NT - 1) / NT;
  const unsigned char* pairs = mem;
  const unsigned char* tab = mem + JacobiTab<D>::PAIRS_BYTES;
  const unsigned char* tab2 = tab + JacobiTab<D>::TAB_BYTES;
  const int tid = threadIdx.x, lane = tid & 31;
  if (init_v) {
    for (int e = tid; e < E; e += NT) V[e] = cmk((e % D == e / D) ? 1.0 : 0.0, 0.0);
  }
  double g = 0.0;
#pragma unroll
  for (int q = 0; q < D; ++q) g = fmax(g, fabs(Ga[(D + 1) * q].x));
  const double floor_abs = floor_rel * g;
  const double floor2 = floor_abs * floor_abs;
  const double inv_g = (g > 0.0) ? 1.0 / g : 0.0;
  __syncthreads();
  const int kp = lane % H;          // rotation derived by this lane
  cd* Gin = Ga;
  cd* Gout = Gb;
  int sweeps = 0;
  while (sweeps < max_sweeps) {
    double smax = 0.0;
    for (int rd = 0; rd < nrounds; ++rd) {
      const unsigned char* pr = pairs + 2 * H * rd;
      const int p = pr[2 * kp], q = pr[2 * kp + 1];
      double c;
      cd sg;
      jacobi_rot_sg(Gin[(D + 1) * p].x, Gin[(D + 1) * q].x, Gin[p + D * q], floor2, inv_g, c, sg, skip_below);
      smax = fmax(smax, cabs2(sg));
#pragma unroll
      for (int u = 0; u < EPT; ++u) {
        const int e = tid + u * NT;
        const bool on = e < E;
        const int i = on ? e % D : 0, j = on ? e / D : 0;
        const int ti = tab[D * rd + i], tj = tab[D * rd + j];
        const int pi = tab2[D * rd + i], pj = tab2[D * rd + j];
        const int ka = ti & 127, kb = tj & 127;
        const double ca = __shfl_sync(0xffffffffu, c, ka), cb = __shfl_sync(0xffffffffu, c, kb);
        const cd sa = cmk(__shfl_sync(0xffffffffu, sg.x, ka), __shfl_sync(0xffffffffu, sg.y, ka));
        const cd sb = cmk(__shfl_sync(0xffffffffu, sg.x, kb), __shfl_sync(0xffffffffu, sg.y, kb));
        if (on) {
          const cd gij = Gin[i + D * j], gipj = Gin[i + D * pj], gpij = Gin[pi + D * j], gpipj = Gin[pi + D * pj];
          // J' row i: (ca, ap): p-member: new_p = c row_p - sg row_q ; q-member: new_q = conj(sg) row_p + c row_q
          const cd ap = (ti & 128) ? cmk(sa.x, -sa.y) : cmk(-sa.x, -sa.y);
          // J col j:  (cb, bp): p-member: new_p = c col_p - conj(sg) col_q ; q-member: new_q = sg col_p + c col_q
          const cd bp = (tj & 128) ? sb : cmk(-sb.x, sb.y);
          const cd t1 = cadd(cscale(gij, cb), cmul(bp, gipj));
          const cd t2 = cadd(cscale(gpij, cb), cmul(bp, gpipj));
          cd n = cadd(cscale(t1, ca), cmul(ap, t2));
          if (i == j) n.y = 0.0;
          Gout[e] = n;
        }
      }
      // V columns: item (pair kv, row iv), H*D items
      for (int it = tid; it < ((H * D + 31) / 32) * 32; it += NT) {
        const bool on = it < H * D;
        const int kv = on ? it / D : 0, iv = on ? it % D : 0;
        const double cv = __shfl_sync(0xffffffffu, c, kv);
        const cd sv = cmk(__shfl_sync(0xffffffffu, sg.x, kv), __shfl_sync(0xffffffffu, sg.y, kv));
        if (on) {
          const int pv = pr[2 * kv], qv = pr[2 * kv + 1];
          const cd vp = V[iv + D * pv], vq = V[iv + D * qv];
          V[iv + D * pv] = csub(cscale(vp, cv), cmulc(sv, vq));
          V[iv + D * qv] = cadd(cmul(sv, vp), cscale(vq, cv));
        }
      }
      __syncthreads();
      cd* t = Gin; Gin = Gout; Gout = t;
    }
    ++sweeps;
    // quadratic convergence: a sweep whose largest rotation had |sin| <= 1e-8 leaves off-diagonals at the
    // 1e-16 level, so no verification sweep is needed (smax holds |sin|^2)
    const int big = __syncthreads_or(smax > stop_sin2);
    if (any_rotation) *any_rotation = big;     // (same value in every thread)
    if (!big) break;
  }
  if (Gin != Ga) {   // odd number of rounds: result is in Gb
    for (int e = tid; e < E; e += NT) Ga[e] = Gb[e];
    __syncthreads();
  }
  return sweeps;
}

// ------------------------------------------------------------------------------------------
// Software-pipelined variant of jacobi_small: same rotations, same element formula, bit-identical results.
// The serial part of a round is the derivation of the next rotations (FP32 angle + FP64 renormalisation,
// ~400 cycles of dependent latency), which needs only THREE updated entries per pair.  Warp 0 ("pilot") computes
// those entries for the pairs of the NEXT round straight from the old matrix and the current rotations, derives
// the next rotations and publishes them in shared memory, while warps 1..7 apply the current rotations to all
// D*D entries and to V.  One barrier per round; the derivation is off the critical path of the update.
// mem: JacobiTab<D>::BYTES (tables followed by the double-buffered rotation parameters).
template <int D>
__device__ inline int jacobi_small_p(cd* Ga, cd* Gb, cd* V, unsigned char* mem, bool init_v,
                                     int max_sweeps = 30, int* any_rotation = nullptr, double skip_below = 0.0,
                                     double floor_rel = 1.0e-18, double stop_sin2 = 1.0e-16, int nrounds = D - 1) {
  static_assert(D % 2 == 0 && D <= 32, "unsupported dimension");
  constexpr int H = D / 2, E = D * D, NB = NT - 32;
  const unsigned char* pairs = mem;
  const unsigned char* tab = mem + JacobiTab<D>::PAIRS_BYTES;
  const unsigned char* tab2 = tab + JacobiTab<D>::TAB_BYTES;
  double* prm = reinterpret_cast<double*>(mem + JacobiTab<D>::TABLES);   // [2][H][4]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (init_v) {
    for (int e = tid; e < E; e += NT) V[e] = cmk((e % D == e / D) ? 1.0 : 0.0, 0.0);
  }
  double g = 0.0;
#pragma unroll
  for (int q = 0; q < D; ++q) g = fmax(g, fabs(Ga[(D + 1) * q].x));
  const double floor_abs = floor_rel * g;
  const double floor2 = floor_abs * floor_abs;
  const double inv_g = (g > 0.0) ? 1.0 / g : 0.0;
  if (tid < H) {   // rotations of round 0 from the matrix as given
    const int p = pairs[2 * tid], q = pairs[2 * tid + 1];
    double c;
    cd sg;
    jacobi_rot_sg(Ga[(D + 1) * p].x, Ga[(D + 1) * q].x, Ga[p + D * q], floor2, inv_g, c, sg, skip_below);
    prm[4 * tid] = c; prm[4 * tid + 1] = sg.x; prm[4 * tid + 2] = sg.y; prm[4 * tid + 3] = cabs2(sg);
  }
  __syncthreads();
  cd* Gin = Ga;
  cd* Gout = Gb;
  int sweeps = 0, cur = 0;
  // entry (i, j) of J' G J for the rotations P of round rd, from the read-only Gin
  auto elem = [&](const cd* Gi, const double* P, int rd, int i, int j) -> cd {
    const int ti = tab[D * rd + i], tj = tab[D * rd + j];
    const int pi = tab2[D * rd + i], pj = tab2[D * rd + j];
    const int ka = ti & 127, kb = tj & 127;
    const double ca = P[4 * ka], cb = P[4 * kb];
    const cd sa = cmk(P[4 * ka + 1], P[4 * ka + 2]), sb = cmk(P[4 * kb + 1], P[4 * kb + 2]);
    const cd gij = Gi[i + D * j], gipj = Gi[i + D * pj], gpij = Gi[pi + D * j], gpipj = Gi[pi + D * pj];
    const cd ap = (ti & 128) ? cmk(sa.x, -sa.y) : cmk(-sa.x, -sa.y);
    const cd bp = (tj & 128) ? sb : cmk(-sb.x, sb.y);
    const cd t1 = cadd(cscale(gij, cb), cmul(bp, gipj));
    const cd t2 = cadd(cscale(gpij, cb), cmul(bp, gpipj));
    cd n = cadd(cscale(t1, ca), cmul(ap, t2));
    if (i == j) n.y = 0.0;
    return n;
  };
  while (sweeps < max_sweeps) {
    double smax = 0.0;
    for (int rd = 0; rd < nrounds; ++rd) {
      const double* P = prm + cur * (4 * H);
      double* Pn = prm + (cur ^ 1) * (4 * H);
      if (warp == 0) {
        if (lane < H) {
          smax = fmax(smax, P[4 * lane + 3]);
          const int rn = (rd + 1 == nrounds) ? 0 : rd + 1;
          const int p = pairs[2 * (H * rn + lane)], q = pairs[2 * (H * rn + lane) + 1];
          const cd npp = elem(Gin, P, rd, p, p), nqq = elem(Gin, P, rd, q, q), npq = elem(Gin, P, rd, p, q);
          double c;
          cd sg;
          jacobi_rot_sg(npp.x, nqq.x, npq, floor2, inv_g, c, sg, skip_below);
          Pn[4 * lane] = c; Pn[4 * lane + 1] = sg.x; Pn[4 * lane + 2] = sg.y; Pn[4 * lane + 3] = cabs2(sg);
        }
      } else {
        const int bt = tid - 32;
        for (int e = bt; e < E; e += NB) Gout[e] = elem(Gin, P, rd, e % D, e / D);
        const unsigned char* pr = pairs + 2 * H * rd;
        for (int it = bt; it < H * D; it += NB) {
          const int kv = it / D, iv = it - kv * D;
          const double cv = P[4 * kv];
          const cd sv = cmk(P[4 * kv + 1], P[4 * kv + 2]);
          const int pv = pr[2 * kv], qv = pr[2 * kv + 1];
          const cd vp = V[iv + D * pv], vq = V[iv + D * qv];
          V[iv + D * pv] = csub(cscale(vp, cv), cmulc(sv, vq));
          V[iv + D * qv] = cadd(cmul(sv, vp), cscale(vq, cv));
        }
      }
      __syncthreads();
      cd* t = Gin; Gin = Gout; Gout = t;
      cur ^= 1;
    }
    ++sweeps;
    const int big = __syncthreads_or(smax > stop_sin2);
    if (any_rotation) *any_rotation = big;
    if (!big) break;
  }
  if (Gin != Ga) {
    for (int e = tid; e < E; e += NT) Ga[e] = Gb[e];
    __syncthreads();
  }
  return sweeps;
}

// ------------------------------------------------------------------------------------------
// Block Jacobi for a d x d Hermitian matrix in global memory (d up to a few hundred): 16-wide index
// blocks, round-robin over block pairs; each pair (I, J) is a 32 x 32 subproblem solved in shared memory by
// one sweep of jacobi_small<32> (accumulating its rotation Q), after which Q is applied to the block
// columns / rows of G and the block columns of V with small dense products.  Compared with element-wise
// rotations on the global matrix this divides the memory traffic per sweep by ~16 (6 d^3 complex MACs per
// sweep either way).  S, Sb, Q: shared 32 x 32 buffers with Sb == S + 1024 (S | Sb also hold the padded copy of
// Q during the updates); tab: 2 * JacobiTab<32>::BYTES + 64 bytes of shared memory (full and cross orderings,
// block flags).  act_thr: after the first sweep only block pairs that contain a diagonal entry > act_thr are
// processed (the caller only needs the eigenpairs above that threshold resolved to full accuracy; the others
// stay at the accuracy one sweep gives, which keeps the basis a good warm start); in the first sweep such pairs
// are skipped when their off-diagonal block is below skip_loose2 (squared Frobenius norm).
// Warm start: init_v == false, V holds an orthonormal basis and G = V' G0 V.  stop_sin2: a sweep whose largest
// rotation has |sin|^2 <= stop_sin2 ends the iteration.  skip_abs2 > 0: a block pair whose off-diagonal block
// has squared Frobenius norm <= skip_abs2 is left alone (its rotations would be below the caller's accuracy).
__device__ inline int block_jacobi_heig(cd* G, int ldg, cd* V, int ldv, int d, cd* S, cd* Sb, cd* Q,
                                        unsigned char* tab, int max_sweeps = 30, bool init_v = true,
                                        double stop_sin2 = 1.0e-16, double skip_abs2 = 0.0,
                                        long long* prof = nullptr, double act_thr = -INFINITY,
                                        double skip_loose2 = 0.0) {
  constexpr int B = 16, D2 = 2 * B;
  const int tid = threadIdx.x;
  if (init_v) {
    for (int idx = tid; idx < d * d; idx += NT) {
      const int i = idx % d, j = idx / d;
      V[i + (size_t)ldv * j] = cmk(i == j ? 1.0 : 0.0, 0.0);
    }
  }
  jacobi_tables<D2>(tab);
  unsigned char* tabx = tab + JacobiTab<D2>::BYTES;
  jacobi_tables_cross<D2>(tabx);
  __syncthreads();
  const int nbk = (d + B - 1) / B;
  if (nbk < 2) {   // single block: pad to 32 and solve directly
    for (int e = tid; e < D2 * D2; e += NT) {
      const int i = e % D2, j = e / D2;
      S[e] = (i < d && j < d) ? G[i + (size_t)ldg * j] : cmk(0.0, 0.0);
    }
    __syncthreads();
    const int sw = jacobi_small_p<D2>(S, Sb, Q, tab, true, max_sweeps, nullptr, 0.0, 1.0e-18, stop_sin2);
    for (int e = tid; e < D2 * D2; e += NT) {
      const int i = e % D2, j = e / D2;
      if (i < d && j < d) G[i + (size_t)ldg * j] = S[e];
    }
    if (init_v) {
      for (int e = tid; e < D2 * D2; e += NT) {
        const int i = e % D2, j = e / D2;
        if (i < d && j < d) V[i + (size_t)ldv * j] = Q[e];
      }
    } else {   // V <- V * Q (d <= 16: one row of V per thread)
      cd row[B];
      const bool on = tid < d;
      if (on) {
        for (int c = 0; c < d; ++c) {
          cd a = cmk(0.0, 0.0);
          for (int u = 0; u < d; ++u) cfma(a, V[tid + (size_t)ldv * u], Q[u + D2 * c]);
          row[c] = a;
        }
      }
      __syncthreads();
      if (on) for (int c = 0; c < d; ++c) V[tid + (size_t)ldv * c] = row[c];
    }
    __syncthreads();
    return sw;
  }
  const int nb2 = (nbk + 1) / 2, rounds = 2 * nb2 - 1;
  int* imp = reinterpret_cast<int*>(tab + 2 * JacobiTab<D2>::BYTES);   // [16] block holds an entry > act_thr
  int sweeps = 0;
  for (; sweeps < max_sweeps; ++sweeps) {
    int rotated = 0;
    if (act_thr > -INFINITY) {
      if (tid < 16) imp[tid] = 0;
      __syncthreads();
      for (int i = tid; i < d; i += NT)
        if (G[i + (size_t)ldg * i].x > act_thr) imp[i / B] = 1;
      __syncthreads();
    }
    for (int rd = 0; rd < rounds; ++rd) {
      for (int kb = 0; kb < nb2; ++kb) {
        int bi, bj;
        rr_pair(nb2, rd, kb, bi, bj);
        if (sweeps > 0 && act_thr > -INFINITY && !(imp[bi] || (bj < nbk && imp[bj]))) continue;
        if (bj >= nbk && rd != 0) continue;      // dummy block (odd block count); in round 0 its partner is
                                                 // still swept internally (zero-padded second half)
        // global index of subproblem index u (0..31)
        auto gidx = [&](int u) { return (u < B) ? bi * B + u : bj * B + (u - B); };
        double off2 = 0.0;
        for (int e = tid; e < D2 * D2; e += NT) {
          const int ui = e % D2, uj = e / D2;
          const int gi = gidx(ui), gj = gidx(uj);
          const cd v = (gi < d && gj < d) ? G[gi + (size_t)ldg * gj] : cmk(0.0, 0.0);
          S[e] = v;
          // round 0 sweeps the whole subproblem (all pairs), the other rounds only the off-diagonal block
          if (rd == 0 ? (ui > uj) : (ui >= B && uj < B)) off2 += cabs2(v);
        }
        if (skip_abs2 > 0.0) {   // (uniform branch; contains the barrier the load needs)
          double t[1] = {off2};
          block_sum<1>(t, reinterpret_cast<double*>(Sb));
          // block pairs without an entry above act_thr only have to keep the basis a good warm start
          const bool important = !(act_thr > -INFINITY) || imp[bi] || (bj < nbk && imp[bj]);
          if (t[0] <= (important ? skip_abs2 : fmax(skip_abs2, skip_loose2))) continue;
        } else {
          __syncthreads();
        }
        int big = 0;
        const long long tp0 = prof ? clock64() : 0;
        // round 0 is a perfect matching of the blocks: the full 32 x 32 ordering there sweeps the pairs inside
        // every diagonal block once per sweep; all other block pairs only need their cross pairs
        if (rd == 0) jacobi_small_p<D2>(S, Sb, Q, tab, true, 1, &big, 0.0, 1.0e-18, stop_sin2);
        else jacobi_small_p<D2>(S, Sb, Q, tabx, true, 1, &big, 0.0, 1.0e-18, stop_sin2, B);
        if (prof) prof[0] += clock64() - tp0;
        rotated |= big;
        // write the rotated diagonal/off-diagonal blocks back
        for (int e = tid; e < D2 * D2; e += NT) {
          const int gi = gidx(e % D2), gj = gidx(e / D2);
          if (gi < d && gj < d) G[gi + (size_t)ldg * gj] = S[e];
        }
        // ---- columns: [G; V][:, IJ] <- [G; V][:, IJ] * Q  for the rows outside the pair (G) / all rows (V).
        // item (row, quarter): row `rr` of G and row `rr` of V share every Q operand (one shared-memory load per
        // two complex FMAs); 2 x 8 output columns.  The four quarters of a row sit in adjacent lanes: a warp-level
        // barrier separates their loads from their stores.  Column, row and S updates touch disjoint entries, so
        // no block barrier is needed between them.
        constexpr int QC = B / 2;   // output columns per item: qt, qt + 4, ..., qt + 4 (QC - 1)
        // Q re-laid with leading dimension 33 in the (now free) S | Sb area: the four lanes of a row read columns
        // qt + 4c, i.e. four adjacent padded columns = four different bank groups (no conflict; with ld 32 every
        // column of Q starts in bank 0)
        constexpr int LDQ = D2 + 1;
        cd* Qp = S;
        __syncthreads();
        for (int e = tid; e < D2 * D2; e += NT) Qp[(e % D2) + LDQ * (e / D2)] = Q[e];
        __syncthreads();
        for (int base = 0; base < 4 * d; base += NT) {
          const int it = base + tid;
          const bool on = it < 4 * d;
          const int rr = on ? (it >> 2) : 0, qt = it & 3;
          const bool doG = on && !((rr / B) == bi || (rr / B) == bj);   // pair rows of G: done in S
          cd accg[QC], accv[QC];
#pragma unroll
          for (int c = 0; c < QC; ++c) { accg[c] = cmk(0.0, 0.0); accv[c] = cmk(0.0, 0.0); }
          if (on) {
#pragma unroll 1
            for (int u0 = 0; u0 < D2; u0 += 8) {
              cd xg[8], xv[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const int gu = gidx(u0 + u);
                xg[u] = (doG && gu < d) ? G[rr + (size_t)ldg * gu] : cmk(0.0, 0.0);
                xv[u] = (gu < d) ? V[rr + (size_t)ldv * gu] : cmk(0.0, 0.0);
              }
#pragma unroll
              for (int u = 0; u < 8; ++u) {
#pragma unroll
                for (int c = 0; c < QC; ++c) {
                  const cd q = Qp[(u0 + u) + LDQ * (qt + 4 * c)];
                  cfma(accg[c], xg[u], q);
                  cfma(accv[c], xv[u], q);
                }
              }
            }
          }
          __syncwarp();
          if (on) {
#pragma unroll
            for (int c = 0; c < QC; ++c) {
              const int gc = gidx(qt + 4 * c);
              if (gc < d) {
                if (doG) G[rr + (size_t)ldg * gc] = accg[c];
                V[rr + (size_t)ldv * gc] = accv[c];
              }
            }
          }
        }
        // ---- rows: G[IJ, :] <- Q' * G[IJ, :]  for the columns outside the pair; item (column pair, quarter):
        // columns `col` and `col + hc` share the Q operands
        const int hc = (d + 1) / 2;
        for (int base = 0; base < 4 * hc; base += NT) {
          const int it = base + tid;
          const bool on = it < 4 * hc;
          const int c0 = on ? (it >> 2) : 0, c1 = c0 + hc, qt = it & 3;
          const bool do0 = on && !((c0 / B) == bi || (c0 / B) == bj);
          const bool do1 = on && c1 < d && !((c1 / B) == bi || (c1 / B) == bj);
          cd acc0[QC], acc1[QC];
#pragma unroll
          for (int c = 0; c < QC; ++c) { acc0[c] = cmk(0.0, 0.0); acc1[c] = cmk(0.0, 0.0); }
          if (do0 || do1) {
#pragma unroll 1
            for (int u0 = 0; u0 < D2; u0 += 8) {
              cd x0[8], x1[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const int gu = gidx(u0 + u);
                x0[u] = (do0 && gu < d) ? G[gu + (size_t)ldg * c0] : cmk(0.0, 0.0);
                x1[u] = (do1 && gu < d) ? G[gu + (size_t)ldg * c1] : cmk(0.0, 0.0);
              }
#pragma unroll
              for (int u = 0; u < 8; ++u) {
#pragma unroll
                for (int c = 0; c < QC; ++c) {
                  const cd q = Qp[(u0 + u) + LDQ * (qt + 4 * c)];
                  cfmac(acc0[c], q, x0[u]);   // conj(Q[u, c']) * x
                  cfmac(acc1[c], q, x1[u]);
                }
              }
            }
          }
          __syncwarp();
#pragma unroll
          for (int c = 0; c < QC; ++c) {
            const int gc = gidx(qt + 4 * c);
            if (gc < d) {
              if (do0) G[gc + (size_t)ldg * c0] = acc0[c];
              if (do1) G[gc + (size_t)ldg * c1] = acc1[c];
            }
          }
        }
        __syncthreads();
      }
    }
    if (!rotated) break;
  }
  return sweeps;
}

// In-place inverse of a d x d Hermitian positive-definite matrix (Gauss-Jordan, no pivoting).
// S column-major with leading dimension d (generic pointer).  colk/rowk: shared scratch [d] each.
// Replaces inv(A'*A + eye(n)) at inferLowRankV4.m:221,267 (or its Woodbury core I + A*A').
// The columns can be split over `nparts` cooperating CTAs (a cluster): part p updates columns [d p / nparts,
// d (p + 1) / nparts) of every pivot step and `sync_all()` (a cluster barrier that also orders global memory) separates
// the steps; with nparts = 1 it is a block barrier.  One thread owns one row per pass, walks its columns four at a time
// (independent load / FMA / store chains: the matrix lives in L2, so the step time is load latency over the number of
// loads in flight) and keeps its pivot-column entry in a register.  Element arithmetic as in the textbook update, so the
// result does not depend on the split.
template <class SyncF>
__device__ inline void spd_inverse_part(cd* S, int d, cd* colk, cd* rowk, int part, int nparts, SyncF sync_all) {
  const int tid = threadIdx.x;
  const int j0 = (int)((long long)d * part / nparts), j1 = (int)((long long)d * (part + 1) / nparts);
  for (int k = 0; k < d; ++k) {
    for (int i = tid; i < d; i += NT) {      // (L2 loads: with nparts > 1 other CTAs wrote these entries in the last step)
      colk[i] = __ldcg(S + i + (size_t)d * k);
      rowk[i] = __ldcg(S + k + (size_t)d * i);
    }
    __syncthreads();
    const cd pv = colk[k];
    // 1/pv (pivot is real positive for an HPD matrix up to rounding; keep it complex for safety)
    const double den = cabs2(pv);
    const cd ip = cmk(pv.x / den, -pv.y / den);
    for (int i = tid; i < d; i += NT) {
      cd* row = S + i;
      if (i == k) {
        for (int j = j0; j < j1; ++j) row[(size_t)d * j] = (j == k) ? ip : cmul(rowk[j], ip);
      } else {
        const cd f = cmul(colk[i], ip);
        int j = j0;
        for (; j + 3 < j1; j += 4) {
          const cd c0 = row[(size_t)d * j], c1 = row[(size_t)d * (j + 1)], c2 = row[(size_t)d * (j + 2)], c3 = row[(size_t)d * (j + 3)];
          const cd r0 = rowk[j], r1 = rowk[j + 1], r2 = rowk[j + 2], r3 = rowk[j + 3];
          cd v0 = cmk(c0.x - (f.x * r0.x - f.y * r0.y), c0.y - (f.x * r0.y + f.y * r0.x));
          cd v1 = cmk(c1.x - (f.x * r1.x - f.y * r1.y), c1.y - (f.x * r1.y + f.y * r1.x));
          cd v2 = cmk(c2.x - (f.x * r2.x - f.y * r2.y), c2.y - (f.x * r2.y + f.y * r2.x));
          cd v3 = cmk(c3.x - (f.x * r3.x - f.y * r3.y), c3.y - (f.x * r3.y + f.y * r3.x));
          if (k >= j && k < j + 4) {
            const cd nf = cmk(-f.x, -f.y);
            if (k == j) v0 = nf; else if (k == j + 1) v1 = nf; else if (k == j + 2) v2 = nf; else v3 = nf;
          }
          row[(size_t)d * j] = v0; row[(size_t)d * (j + 1)] = v1; row[(size_t)d * (j + 2)] = v2; row[(size_t)d * (j + 3)] = v3;
        }
        for (; j < j1; ++j) {
          const cd cur = row[(size_t)d * j], rj = rowk[j];
          row[(size_t)d * j] = (j == k) ? cmk(-f.x, -f.y)
                                        : cmk(cur.x - (f.x * rj.x - f.y * rj.y), cur.y - (f.x * rj.y + f.y * rj.x));
        }
      }
    }
    sync_all();
  }
}

__device__ inline void spd_inverse(cd* S, int d, cd* colk, cd* rowk) {
  spd_inverse_part(S, d, colk, rowk, 0, 1, [] { __syncthreads(); });
}

}  // namespace twoace
