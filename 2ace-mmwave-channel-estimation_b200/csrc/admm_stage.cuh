// General-size InferADMM kernel: one CTA per (instance, stage), FP64 complex, state in a per-CTA
// workspace that stays L2-resident, operands staged through shared memory.
//
// Restates, as a batched device program, the loop of
//   main/src/my_recovery_algorithms/ADMM_v2/inferLowRankV4.m:260-365 (InferADMM)
// with ArgMinX :380-388, ArgMinY :490-512, normalize_rows :517-538, ArgMinZ :402-464 and the
// nuclear ArgMinZ of inferLowRank_Nuclear.m:411-439.  Differences from the reference formulation
// (all exact in exact arithmetic):
//   * inv(A'A+I) is never formed when m^2 + n*m < n^2: with Q = Z - N/mu, T = Y - M/mu, S = I + A A'
//     (m x m), Woodbury gives  X = Q + A' W,  A X = T - W,  W = S^-1 (T - A Q): two A-products.
//   * A'*Y (:309) is only computed when a tolerance is non-zero; it feeds nothing but res_dual.
#pragma once
#include "common.cuh"
#include "tasks.h"

namespace twoace {

constexpr int RCH = 10;        // columns of the iterate held in registers per pass
constexpr int QT = 64;         // reduction-tile length staged in shared memory
constexpr int EC = 64;         // E-columns per ArgMinZ tile
constexpr int ECP = EC + 1;    // padded (odd) so row-strided shared reads are conflict-free
constexpr int SMALL_DMAX = 32; // largest tx (or r for the nuclear Gram) handled by the shared eig

struct StageDims {
  int n, tx, rx;
  int maxm, maxr;
  int dmax;          // max over tasks of the inverse dimension (m if Woodbury else n)
  int ds;            // dimension of the shared-memory eigenproblem (tx, or max(tx, maxr) if nuclear)
  int wpr;           // 2-bit code words per row of A (n / 16) when the launch keeps the codes of A in shared memory, else 0
  size_t ws_stride;  // workspace elements (cd) per CTA slot
};

__host__ __device__ inline bool use_woodbury(int m, int n) {
  return (long long)m * m + (long long)n * m < (long long)n * n;
}

struct StageWS {
  cd *Acm, *Sinv, *X, *Z, *N, *optX, *Vb, *Y, *M, *AX, *Wb, *optY, *AtY0, *AtY1;
};

__host__ __device__ inline size_t stage_ws_elems(const StageDims& d) {
  size_t nr = (size_t)d.n * d.maxr, mr = (size_t)d.maxm * d.maxr;
  return (size_t)d.maxm * d.n + (size_t)d.dmax * d.dmax + 5 * nr + 5 * mr + 2 * nr;
}

__device__ inline StageWS carve_ws(cd* p, const StageDims& d) {
  StageWS w;
  size_t nr = (size_t)d.n * d.maxr, mr = (size_t)d.maxm * d.maxr;
  w.Acm = p;  p += (size_t)d.maxm * d.n;
  w.Sinv = p; p += (size_t)d.dmax * d.dmax;
  w.X = p; p += nr;  w.Z = p; p += nr;  w.N = p; p += nr;  w.optX = p; p += nr;  w.Vb = p; p += nr;
  w.Y = p; p += mr;  w.M = p; p += mr;  w.AX = p; p += mr; w.Wb = p; p += mr;  w.optY = p; p += mr;
  w.AtY0 = p; p += nr; w.AtY1 = p; p += nr;
  return w;
}

struct StageSmem {
  double* Bs;      // [maxm]
  int* rows_s;     // [maxm]
  cd* tile;        // [QT*RCH]
  cd* big;         // [max(NT*RCH, ds*ECP, 2*dmax)]  K-split partials | ArgMinZ tile | GJ pivot row/col
  cd *G, *U, *P;   // [ds*ds] each
  JacobiScratch js;
  double* red;     // [16*NW]
  double* s2;      // [ds]
  double* s2s;     // [ds] sqrt(s2_scale) (V4) or shrink factor (nuclear)
  double* colsc;   // [SMALL_DMAX] per-column scalars
  double* sc;      // [32] broadcast scalars
  int* ifl;        // [8] broadcast ints
  unsigned char* jtab;   // tables of jacobi_small<16> / jacobi_small_p<32>
  uint32_t* cw;          // [maxm * (wpr + 1)] 2-bit codes of the task's rows of A (row stride padded: conflict-free)
};

__host__ __device__ inline size_t stage_big_elems(const StageDims& d) {
  size_t a = (size_t)NT * RCH, b = (size_t)d.ds * ECP, c = 2 * (size_t)d.dmax;
  size_t m = a > b ? a : b;
  return m > c ? m : c;
}

__host__ __device__ inline size_t stage_smem_bytes(const StageDims& d) {
  size_t b = 0;
  b += ((size_t)d.maxm * sizeof(double) + 15) / 16 * 16;
  b += ((size_t)d.maxm * sizeof(int) + 15) / 16 * 16;
  b += (size_t)QT * RCH * sizeof(cd);
  b += stage_big_elems(d) * sizeof(cd);
  b += 3 * (size_t)d.ds * d.ds * sizeof(cd);
  b += (size_t)(d.ds / 2 + 2) * (sizeof(cd) + 2 * sizeof(double));
  b += 16 * NW * sizeof(double);
  b += 2 * (size_t)d.ds * sizeof(double);
  b += SMALL_DMAX * sizeof(double) + 32 * sizeof(double) + 16 * sizeof(int);
  b += (d.tx == 32 ? JacobiTab<32>::BYTES : JacobiTab<16>::BYTES) + 16;
  if (d.wpr > 0) b += (size_t)d.maxm * (d.wpr + 1) * sizeof(uint32_t) + 16;
  return b + 64;
}

__device__ inline StageSmem carve_smem(unsigned char* p, const StageDims& d) {
  StageSmem s;
  s.Bs = (double*)p;   p += ((size_t)d.maxm * sizeof(double) + 15) / 16 * 16;
  s.rows_s = (int*)p;  p += ((size_t)d.maxm * sizeof(int) + 15) / 16 * 16;
  s.tile = (cd*)p;     p += (size_t)QT * RCH * sizeof(cd);
  s.big = (cd*)p;      p += stage_big_elems(d) * sizeof(cd);
  s.G = (cd*)p;        p += (size_t)d.ds * d.ds * sizeof(cd);
  s.U = (cd*)p;        p += (size_t)d.ds * d.ds * sizeof(cd);
  s.P = (cd*)p;        p += (size_t)d.ds * d.ds * sizeof(cd);
  const int h = d.ds / 2 + 2;
  s.js.e = (cd*)p;       p += (size_t)h * sizeof(cd);
  s.js.cs = (double*)p;  p += (size_t)h * sizeof(double);
  s.js.sn = (double*)p;  p += (size_t)h * sizeof(double);
  s.red = (double*)p;    p += 16 * NW * sizeof(double);
  s.s2 = (double*)p;     p += (size_t)d.ds * sizeof(double);
  s.s2s = (double*)p;    p += (size_t)d.ds * sizeof(double);
  s.colsc = (double*)p;  p += SMALL_DMAX * sizeof(double);
  s.sc = (double*)p;     p += 32 * sizeof(double);
  s.ifl = (int*)p;       p += 16 * sizeof(int);
  s.jtab = p;            p += ((d.tx == 32 ? JacobiTab<32>::BYTES : JacobiTab<16>::BYTES) + 15) / 16 * 16;
  s.cw = (uint32_t*)p;
  s.js.flag = s.ifl + 8;
  s.js.gscale = s.sc + 31;
  return s;
}

// ------------------------------------------------------------------------------------------
// out(o, c) = store( sum_q mat(q, o) * opnd(q, c) ), o in [0,nout), c in [0,r), q in [0,K).
// One thread per output row o holds RCH column accumulators; opnd is staged in shared tiles of
// QT x RCH; mat(q, o) must be coalesced over o.  When nout is small the reduction is split over
// up to 8 thread groups and the partials are combined through `ksred`.
template <int RC, class MatF, class OpF, class StoreF>
__device__ __forceinline__ void gemm_tpo_t(int nout, int K, int r, MatF mat, OpF opnd, StoreF store,
                                         cd* tile, cd* ksred) {
  const int tid = threadIdx.x;
  int ks = 1;
  while (ks < 8 && 2 * ks * nout <= NT) ks *= 2;
  const int OB = (ks == 1) ? NT : nout;   // outputs per pass
  const int o_in = tid % OB, kslice = tid / OB;
  for (int c0 = 0; c0 < r; c0 += RC) {
    const int rc = min(RC, r - c0);
    for (int ob = 0; ob < nout; ob += OB) {
      const int o = ob + o_in;
      const bool act = (o < nout) && (kslice < ks);
      cd acc[RC];
#pragma unroll
      for (int j = 0; j < RC; ++j) acc[j] = cmk(0.0, 0.0);
      for (int q0 = 0; q0 < K; q0 += QT) {
        const int ql = min(QT, K - q0);
        __syncthreads();
        for (int idx = tid; idx < QT * RC; idx += NT) {
          int j = idx / QT, q = idx - j * QT;   // q fastest: coalesced operand reads
          tile[q * RC + j] = (q < ql && j < rc) ? opnd(q0 + q, c0 + j) : cmk(0.0, 0.0);
        }
        __syncthreads();
        if (act) {
          // GL elements of `mat` are loaded before the first use: the loop is bound by the latency of those
          // loads (L2 / DRAM), not by their bandwidth
          constexpr int GL = 8;
          int q = kslice;
          for (; q + (GL - 1) * ks < ql; q += GL * ks) {
            cd a[GL];
#pragma unroll
            for (int u = 0; u < GL; ++u) a[u] = mat(q0 + q + u * ks, o);
#pragma unroll
            for (int u = 0; u < GL; ++u) {
              const cd* tr = tile + (q + u * ks) * RC;
#pragma unroll
              for (int j = 0; j < RC; ++j) cfma(acc[j], a[u], tr[j]);
            }
          }
          for (; q < ql; q += ks) {
            const cd a = mat(q0 + q, o);
            const cd* tr = tile + q * RC;
#pragma unroll
            for (int j = 0; j < RC; ++j) cfma(acc[j], a, tr[j]);
          }
        }
      }
      if (ks > 1) {
        __syncthreads();
        if (act) {
#pragma unroll
          for (int j = 0; j < RC; ++j) ksred[((size_t)kslice * nout + o) * RC + j] = acc[j];
        }
        __syncthreads();
        if (act && kslice == 0) {
          for (int s = 1; s < ks; ++s) {
#pragma unroll
            for (int j = 0; j < RC; ++j) {
              cd v = ksred[((size_t)s * nout + o) * RC + j];
              acc[j].x += v.x;
              acc[j].y += v.y;
            }
          }
        }
      }
      if (act && kslice == 0) {
#pragma unroll
        for (int j = 0; j < RC; ++j)
          if (j < rc) store(o, c0 + j, acc[j]);
      }
    }
  }
  __syncthreads();
}

// Two output rows per thread (o and o + half of the pass) and RC columns: every operand read from the shared tile
// feeds two complex FMAs.  With `mat` cheap (2-bit codes) the one-row form is bound by the throughput of those
// warp-broadcast reads (ncu: short scoreboard 55 % of the samples in the loop).  A pass of at most NT rows splits the
// reduction over two thread groups so that all warps stay busy.  nout > NT / 2.
template <int RC, class MatF, class OpF, class StoreF>
__device__ __forceinline__ void gemm_tpo_2r(int nout, int K, int r, MatF mat, OpF opnd, StoreF store, cd* tile,
                                            cd* ksred) {
  constexpr int QT2 = QT * RCH / RC;     // the tile buffer holds QT x RCH elements: longer tiles, fewer fills and barriers
  const int tid = threadIdx.x;
  for (int c0 = 0; c0 < r; c0 += RC) {
    const int rc = min(RC, r - c0);
    for (int ob = 0; ob < nout; ob += 2 * NT) {
      const int rows = min(2 * NT, nout - ob), half = (rows + 1) / 2;
      const int ks = (2 * half <= NT) ? 2 : 1;
      const int kslice = tid / half, o_in = tid - kslice * half;
      const bool act0 = kslice < ks, act1 = act0 && (half + o_in < rows);
      const int o0 = ob + o_in, o1 = act1 ? ob + half + o_in : o0;
      cd acc0[RC], acc1[RC];
#pragma unroll
      for (int j = 0; j < RC; ++j) { acc0[j] = cmk(0.0, 0.0); acc1[j] = cmk(0.0, 0.0); }
      for (int q0 = 0; q0 < K; q0 += QT2) {
        const int ql = min(QT2, K - q0);
        __syncthreads();
        for (int idx = tid; idx < QT2 * RC; idx += NT) {
          int j = idx / QT2, q = idx - j * QT2;
          tile[q * RC + j] = (q < ql && j < rc) ? opnd(q0 + q, c0 + j) : cmk(0.0, 0.0);
        }
        __syncthreads();
        if (act0) {
          constexpr int GL = 4;
          int q = kslice;
          for (; q + (GL - 1) * ks < ql; q += GL * ks) {
            cd a0[GL], a1[GL];
#pragma unroll
            for (int u = 0; u < GL; ++u) { a0[u] = mat(q0 + q + u * ks, o0); a1[u] = mat(q0 + q + u * ks, o1); }
#pragma unroll
            for (int u = 0; u < GL; ++u) {
              const cd* tr = tile + (q + u * ks) * RC;
#pragma unroll
              for (int j = 0; j < RC; ++j) { const cd t = tr[j]; cfma(acc0[j], a0[u], t); cfma(acc1[j], a1[u], t); }
            }
          }
          for (; q < ql; q += ks) {
            const cd a0 = mat(q0 + q, o0), a1 = mat(q0 + q, o1);
            const cd* tr = tile + q * RC;
#pragma unroll
            for (int j = 0; j < RC; ++j) { const cd t = tr[j]; cfma(acc0[j], a0, t); cfma(acc1[j], a1, t); }
          }
        }
      }
      if (ks > 1) {
        __syncthreads();
        if (act0 && kslice == 1) {
#pragma unroll
          for (int j = 0; j < RC; ++j) { ksred[(size_t)o_in * 2 * RC + j] = acc0[j]; ksred[(size_t)o_in * 2 * RC + RC + j] = acc1[j]; }
        }
        __syncthreads();
        if (act0 && kslice == 0) {
#pragma unroll
          for (int j = 0; j < RC; ++j) {
            const cd v0 = ksred[(size_t)o_in * 2 * RC + j], v1 = ksred[(size_t)o_in * 2 * RC + RC + j];
            acc0[j].x += v0.x; acc0[j].y += v0.y;
            acc1[j].x += v1.x; acc1[j].y += v1.y;
          }
        }
      }
      if (act0 && kslice == 0) {
#pragma unroll
        for (int j = 0; j < RC; ++j)
          if (j < rc) {
            store(o0, c0 + j, acc0[j]);
            if (act1) store(o1, c0 + j, acc1[j]);
          }
      }
    }
  }
  __syncthreads();
}

// r == 1 (the refinement stages) runs with one column accumulator instead of RCH; wide outputs with two rows per thread.
template <class MatF, class OpF, class StoreF>
__device__ __forceinline__ void gemm_tpo(int nout, int K, int r, MatF mat, OpF opnd, StoreF store,
                                         cd* tile, cd* ksred) {
  if (r == 1) gemm_tpo_t<1>(nout, K, r, mat, opnd, store, tile, ksred);
  else if (nout > NT / 2) gemm_tpo_2r<RCH / 2>(nout, K, r, mat, opnd, store, tile, ksred);
  else gemm_tpo_t<RCH>(nout, K, r, mat, opnd, store, tile, ksred);
}

// Rank-shaping profile of inferLowRankV4.m:416-443.  Returns the number of (r_k, f_k) stages.
__device__ inline int rank_profile_dev(int tx, int rx, int m, int n, int rank_one, int* rl, double* fl) {
  const int sz = min(rx, tx);
  const double sq = sqrt((double)sz);
  const int r0 = (int)ceil(sq * 0.5), r1 = (int)ceil(sq * 0.7), r2 = (int)ceil(sq);
  const int r3 = min(sz, (int)ceil(sq * 2.0));
  // rank_one also carries the profile of the older solver versions: 2 = inferLowRank.m:407-418,437 (single
  // stage [r2]), 3 = inferLowRankV2.m:418-431 (small-array fallback [r2 r3] instead of [r2])
  if (rank_one == 1) { rl[0] = 1; fl[0] = 0.95; return 1; }
  if (rank_one == 2) { rl[0] = min(sz, r2); fl[0] = 0.95; return 1; }
  if ((long long)m >= (long long)n * 3) { rl[0] = r3; fl[0] = 0.995; return 1; }
  if (r1 <= 2) {
    if (rank_one == 3) { rl[0] = r2; rl[1] = r3; fl[0] = 0.95; fl[1] = 0.995; return 2; }
    rl[0] = r2; fl[0] = 0.95; return 1;
  }
  if (r0 <= 2) { rl[0] = r1; rl[1] = r2; rl[2] = r3; fl[0] = 0.9; fl[1] = 0.95; fl[2] = 0.995; return 3; }
  rl[0] = r0; rl[1] = r1; rl[2] = r2; rl[3] = r3;
  fl[0] = 0.8; fl[1] = 0.9; fl[2] = 0.95; fl[3] = 0.995;
  return 4;
}

// ArgMinZ for the V4 family (inferLowRankV4.m:402-464) fused with the N update (:319-320) and the
// Z/N/X norms of :343-349.  Z_in = X + N*imu is viewed as E = tx x (n*r/tx).
//   init_mode: N == 0, mu == 1, only Z is written (the call at :288).
// Returns (through nrm[4], valid in all threads) |X-Z|^2, |Z-Z0|^2, |X|^2, |Z|^2.
__device__ inline void argmin_z_v4(const StageTask& tk, const StageDims& dm, const StageWS& ws,
                                   const StageSmem& sm, double mu, bool init_mode, int rank_one,
                                   double* nrm, int* sweeps_acc) {
  const int tid = threadIdx.x;
  const int n = dm.n, tx = dm.tx, r = tk.r;
  const int ne = n * r / tx;         // E columns
  const double imu = 1.0 / mu;
  cd* zs = sm.big;
  // ---- Gram G = E E'
  for (int idx = tid; idx < tx * tx; idx += NT) sm.G[idx] = cmk(0.0, 0.0);
  for (int e0 = 0; e0 < ne; e0 += EC) {
    const int ecnt = min(EC, ne - e0);
    __syncthreads();
    for (int idx = tid; idx < tx * ecnt; idx += NT) {
      const int i = idx % tx, el = idx / tx;
      const size_t lin = (size_t)tx * e0 + idx;
      cd v = ws.X[lin];
      if (!init_mode) { cd nn = ws.N[lin]; v.x = fma(nn.x, imu, v.x); v.y = fma(nn.y, imu, v.y); }
      zs[el + ECP * i] = v;
    }
    __syncthreads();
    for (int idx = tid; idx < tx * tx; idx += NT) {
      const int i = idx % tx, j = idx / tx;
      cd acc = cmk(0.0, 0.0);
      const cd* zi = zs + ECP * i;
      const cd* zj = zs + ECP * j;
      for (int e = 0; e < ecnt; ++e) cfmabc(acc, zi[e], zj[e]);
      cd g = sm.G[idx];
      sm.G[idx] = cmk(g.x + acc.x, g.y + acc.y);
    }
  }
  __syncthreads();
  // enforce exact Hermitian symmetry (MATLAB's E*E' is exactly Hermitian)
  for (int idx = tid; idx < tx * tx; idx += NT) {
    const int i = idx % tx, j = idx / tx;
    if (i == j) sm.G[idx].y = 0.0;
    else if (i > j) { cd u = sm.G[j + tx * i]; sm.G[idx] = cmk(u.x, -u.y); }
  }
  __syncthreads();
  // ---- eigen-decomposition.  tx == 16: warm start from the previous eigenvectors (kept in sm.U), exact
  // Schur-Horn screen (if the constraints of :449-459 already hold for the sorted diagonal of U'GU they hold
  // for the spectrum, no stage fires and Z = Z_in without any eigen-decomposition), element-form Jacobi.
  const bool small16 = (tx == 16), small32 = (tx == 32);
  const bool warm = (small16 || small32) && !init_mode && (sm.ifl[2] & 63) != 0;
  bool need_eig = true;
  if (warm) {
    // G <- U' G U (Hermitian result built from its lower triangle)
    const int tt = tx * tx;
    for (int idx = tid; idx < tt; idx += NT) {
      const int i = idx % tx, j = idx / tx;
      cd t0 = cmk(0.0, 0.0), t1 = t0;
      for (int k = 0; k < tx; k += 2) {
        cfma(t0, sm.G[i + tx * k], sm.U[k + tx * j]);
        cfma(t1, sm.G[i + tx * (k + 1)], sm.U[(k + 1) + tx * j]);
      }
      sm.P[idx] = cmk(t0.x + t1.x, t0.y + t1.y);
    }
    __syncthreads();
    for (int idx = tid; idx < tt; idx += NT) {
      const int i = idx % tx, j = idx / tx;
      if (i >= j) {
        cd t0 = cmk(0.0, 0.0), t1 = t0;
        for (int k = 0; k < tx; k += 2) {
          cfmac(t0, sm.U[k + tx * i], sm.P[k + tx * j]);
          cfmac(t1, sm.U[(k + 1) + tx * i], sm.P[(k + 1) + tx * j]);
        }
        cd t = cmk(t0.x + t1.x, t0.y + t1.y);
        if (i == j) t.y = 0.0;
        sm.G[i + tx * j] = t;
        if (i != j) sm.G[j + tx * i] = cmk(t.x, -t.y);
      }
    }
    __syncthreads();
    if (tid == 0) {
      double dg[SMALL_DMAX], pre = 0.0, tot = 0.0;
      for (int i = 0; i < tx; ++i) { dg[i] = fmax(0.0, sm.G[(tx + 1) * i].x); tot += dg[i]; }
      for (int i = 1; i < tx; ++i) {   // descending insertion sort
        const double v = dg[i]; int j = i - 1;
        while (j >= 0 && dg[j] < v) { dg[j + 1] = dg[j]; --j; }
        dg[j + 1] = v;
      }
      int rl[4]; double fl[4];
      const int ns = rank_profile_dev(tx, dm.rx, tk.m, n, rank_one, rl, fl);
      int ok = 1, done = 0;
      for (int k = 0; k < ns; ++k) {
        for (; done < rl[k] && done < tx; ++done) pre += dg[done];
        ok &= (pre >= tot * fl[k] * (1.0 + 1e-12)) ? 1 : 0;
      }
      sm.ifl[1] = ok;
    }
    __syncthreads();
    need_eig = sm.ifl[1] == 0;
  }
  if (!need_eig) {
    if (tid == 0) sm.ifl[0] = 0;
  } else {
  const long long te0 = clock64();
  int sw = small16 ? jacobi_small<16>(sm.G, sm.P, sm.U, sm.jtab, !warm)
         : small32 ? jacobi_small_p<32>(sm.G, sm.P, sm.U, sm.jtab, !warm)
                   : jacobi_heig(sm.G, tx, sm.U, tx, tx, true, sm.js);
  if (tid == 0) {
    sm.sc[30] += (double)(clock64() - te0);     // cycles in the eigensolver (reported in scal[9])
    sm.ifl[2] += 1;
    *sweeps_acc += sw;
    // eigenvalues, clamped (:408); stable descending order (:409)
    int ord[SMALL_DMAX];
    double s2[SMALL_DMAX], scl[SMALL_DMAX];
    for (int i = 0; i < tx; ++i) { s2[i] = fmax(0.0, sm.G[i + tx * i].x); ord[i] = i; scl[i] = 1.0; }
    for (int i = 1; i < tx; ++i) {            // insertion sort, descending, stable
      int oi = ord[i]; double v = s2[oi]; int j = i - 1;
      while (j >= 0 && s2[ord[j]] < v) { ord[j + 1] = ord[j]; --j; }
      ord[j + 1] = oi;
    }
    double ss[SMALL_DMAX];
    for (int i = 0; i < tx; ++i) ss[i] = s2[ord[i]];
    int rl[4]; double fl[4];
    const int ns = rank_profile_dev(tx, dm.rx, tk.m, n, rank_one, rl, fl);
    for (int k = 0; k < ns; ++k) {            // cascade :449-459 (s2 rescaled cumulatively)
      const int rr = rl[k]; const double f = fl[k];
      double vr = 0.0, v = 0.0;
      for (int i = 0; i < rr && i < tx; ++i) vr += ss[i];
      for (int i = 0; i < tx; ++i) v += ss[i];
      if (vr < v * f) {
        const double scale = fmin(1.0, vr / (v - vr) * (1.0 / f - 1.0));
        for (int i = rr; i < tx; ++i) { ss[i] *= scale; scl[ord[i]] *= scale; }
      }
    }
    int any = 0;
    for (int i = 0; i < tx; ++i) { if (scl[i] < 1.0) any = 1; sm.s2s[i] = sqrt(scl[i]); }
    sm.ifl[0] = any;
  }
  }   // need_eig
  __syncthreads();
  const int any = sm.ifl[0];
  if (any) {   // P = U diag(sqrt(s2_scale)) U'   (:462)
    for (int idx = tid; idx < tx * tx; idx += NT) {
      const int i = idx % tx, j = idx / tx;
      cd acc = cmk(0.0, 0.0);
      for (int k = 0; k < tx; ++k) {
        cd ui = sm.U[i + tx * k], uj = sm.U[j + tx * k];
        const double s = sm.s2s[k];
        cfmabc(acc, cmk(ui.x * s, ui.y * s), uj);
      }
      sm.P[idx] = acc;
    }
  }
  __syncthreads();
  // ---- apply, update N, norms
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  const int ng = tx / 4;
  for (int e0 = 0; e0 < ne; e0 += EC) {
    const int ecnt = min(EC, ne - e0);
    __syncthreads();
    for (int idx = tid; idx < tx * ecnt; idx += NT) {
      const int i = idx % tx, el = idx / tx;
      const size_t lin = (size_t)tx * e0 + idx;
      cd v = ws.X[lin];
      if (!init_mode) { cd nn = ws.N[lin]; v.x = fma(nn.x, imu, v.x); v.y = fma(nn.y, imu, v.y); }
      zs[el + ECP * i] = v;
    }
    __syncthreads();
    if (any) {
      // items (el, ig): 4 consecutive output rows per thread; at most ceil(EC*ng/NT) items each
      constexpr int MAXIT = (EC * (SMALL_DMAX / 4) + NT - 1) / NT;
      cd outv[MAXIT][4];
#pragma unroll
      for (int cnt = 0; cnt < MAXIT; ++cnt) {
        const int it = tid + cnt * NT;
        if (it < ecnt * ng) {
          const int el = it % ecnt, ig = it / ecnt;
          cd acc[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) acc[u] = cmk(0.0, 0.0);
          for (int k = 0; k < tx; ++k) {
            const cd z = zs[el + ECP * k];
#pragma unroll
            for (int u = 0; u < 4; ++u) cfma(acc[u], sm.P[(4 * ig + u) + tx * k], z);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) outv[cnt][u] = acc[u];
        }
      }
      __syncthreads();
#pragma unroll
      for (int cnt = 0; cnt < MAXIT; ++cnt) {
        const int it = tid + cnt * NT;
        if (it < ecnt * ng) {
          const int el = it % ecnt, ig = it / ecnt;
#pragma unroll
          for (int u = 0; u < 4; ++u) zs[el + ECP * (4 * ig + u)] = outv[cnt][u];
        }
      }
      __syncthreads();
    }
    for (int idx = tid; idx < tx * ecnt; idx += NT) {
      const int i = idx % tx, el = idx / tx;
      const size_t lin = (size_t)tx * e0 + idx;
      const cd zn = zs[el + ECP * i];
      if (init_mode) {
        ws.Z[lin] = zn;
      } else {
        const cd x = ws.X[lin], nn = ws.N[lin], zo = ws.Z[lin];
        const cd jn = cmk(x.x - zn.x, x.y - zn.y);
        ws.Z[lin] = zn;
        ws.N[lin] = cmk(fma(mu, jn.x, nn.x), fma(mu, jn.y, nn.y));
        a0 += cabs2(jn);
        a1 += cabs2(cmk(zn.x - zo.x, zn.y - zo.y));
        a2 += cabs2(x);
        a3 += cabs2(zn);
      }
    }
  }
  __syncthreads();
  if (!init_mode) {
    double v[4] = {a0, a1, a2, a3};
    block_sum<4>(v, sm.red);
    nrm[0] = v[0]; nrm[1] = v[1]; nrm[2] = v[2]; nrm[3] = v[3];
  }
}

// Nuclear-norm ArgMinZ (inferLowRank_Nuclear.m:411-439): singular-value soft threshold of the
// n x r matrix Z_in by 1/mu, computed through the r x r Gram eigenproblem
//   Z_in = W S V',  Z_in' Z_in = V S^2 V'  =>  Z = Z_in * (V diag(max(0, s - 1/mu)/s) V').
__device__ inline void argmin_z_nuclear(const StageTask& tk, const StageDims& dm, const StageWS& ws,
                                        const StageSmem& sm, double mu, bool init_mode, double* nrm,
                                        int* sweeps_acc) {
  const int tid = threadIdx.x;
  const int n = dm.n, r = tk.r;
  const double imu = 1.0 / mu;
  // zin tile: [KT rows][r] with row pitch r (+1 pad when even) in sm.big
  const int pitch = r | 1;
  const int KT = (int)min((size_t)64, (size_t)(NT * RCH) / pitch);
  cd* zt = sm.big;
  for (int idx = tid; idx < r * r; idx += NT) sm.G[idx] = cmk(0.0, 0.0);
  for (int k0 = 0; k0 < n; k0 += KT) {
    const int kc = min(KT, n - k0);
    __syncthreads();
    for (int idx = tid; idx < kc * r; idx += NT) {
      const int kl = idx % kc, c = idx / kc;
      const size_t lin = (size_t)(k0 + kl) + (size_t)n * c;
      cd v = ws.X[lin];
      if (!init_mode) { cd nn = ws.N[lin]; v.x = fma(nn.x, imu, v.x); v.y = fma(nn.y, imu, v.y); }
      zt[kl * pitch + c] = v;
    }
    __syncthreads();
    for (int idx = tid; idx < r * r; idx += NT) {
      const int c = idx % r, c2 = idx / r;   // G[c, c2] = sum_k conj(z[k,c]) z[k,c2]
      cd acc = cmk(0.0, 0.0);
      for (int kl = 0; kl < kc; ++kl) cfmac(acc, zt[kl * pitch + c], zt[kl * pitch + c2]);
      cd g = sm.G[idx];
      sm.G[idx] = cmk(g.x + acc.x, g.y + acc.y);
    }
  }
  __syncthreads();
  for (int idx = tid; idx < r * r; idx += NT) {
    const int i = idx % r, j = idx / r;
    if (i == j) sm.G[idx].y = 0.0;
    else if (i > j) { cd u = sm.G[j + r * i]; sm.G[idx] = cmk(u.x, -u.y); }
  }
  __syncthreads();
  int sw = jacobi_heig(sm.G, r, sm.U, r, r, true, sm.js);
  if (tid == 0) *sweeps_acc += sw;
  if (tid < r) {
    const double s = sqrt(fmax(0.0, sm.G[tid + r * tid].x));
    sm.s2s[tid] = (s > 0.0) ? fmax(0.0, s - imu) / s : 0.0;
  }
  __syncthreads();
  for (int idx = tid; idx < r * r; idx += NT) {   // P = V diag(f) V'
    const int i = idx % r, j = idx / r;
    cd acc = cmk(0.0, 0.0);
    for (int k = 0; k < r; ++k) {
      cd ui = sm.U[i + r * k], uj = sm.U[j + r * k];
      const double s = sm.s2s[k];
      cfmabc(acc, cmk(ui.x * s, ui.y * s), uj);
    }
    sm.P[idx] = acc;
  }
  __syncthreads();
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  for (int k0 = 0; k0 < n; k0 += KT) {
    const int kc = min(KT, n - k0);
    __syncthreads();
    for (int idx = tid; idx < kc * r; idx += NT) {
      const int kl = idx % kc, c = idx / kc;
      const size_t lin = (size_t)(k0 + kl) + (size_t)n * c;
      cd v = ws.X[lin];
      if (!init_mode) { cd nn = ws.N[lin]; v.x = fma(nn.x, imu, v.x); v.y = fma(nn.y, imu, v.y); }
      zt[kl * pitch + c] = v;
    }
    __syncthreads();
    for (int idx = tid; idx < kc * r; idx += NT) {
      const int kl = idx % kc, c2 = idx / kc;
      cd zn = cmk(0.0, 0.0);
      for (int c = 0; c < r; ++c) cfma(zn, zt[kl * pitch + c], sm.P[c + r * c2]);
      const size_t lin = (size_t)(k0 + kl) + (size_t)n * c2;
      if (init_mode) {
        ws.Z[lin] = zn;
      } else {
        const cd x = ws.X[lin], nn = ws.N[lin], zo = ws.Z[lin];
        const cd jn = cmk(x.x - zn.x, x.y - zn.y);
        ws.Z[lin] = zn;
        ws.N[lin] = cmk(fma(mu, jn.x, nn.x), fma(mu, jn.y, nn.y));
        a0 += cabs2(jn);
        a1 += cabs2(cmk(zn.x - zo.x, zn.y - zo.y));
        a2 += cabs2(x);
        a3 += cabs2(zn);
      }
    }
  }
  __syncthreads();
  if (!init_mode) {
    double v[4] = {a0, a1, a2, a3};
    block_sum<4>(v, sm.red);
    nrm[0] = v[0]; nrm[1] = v[1]; nrm[2] = v[2]; nrm[3] = v[3];
  }
}

// ------------------------------------------------------------------------------------------
// CM: A = (*tk.cscale) * u(code) with the 2-bit codes of the task's rows held in shared memory: the products of the
// iteration read no A from global memory (for n = 1024 the dense A of one instance is 3 MB, of a batch far more than
// L2, and the dense loop is bound by the latency of those loads).  The set-up (S and its inverse) still reads dense A.
template <bool CM>
__device__ inline void run_stage(const StageTask& tk, const DevParams& prm, const StageDims& dm,
                                 const StageWS& ws, const StageSmem& sm) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = dm.n, m = tk.m, r = tk.r;
  const bool wood = use_woodbury(m, n);
  const int dS = wood ? m : n;
  const cd* Ab = tk.A.base;
  const double asc_dense = *tk.A.scale;
  const double asc = CM ? *tk.cscale : asc_dense;      // scale applied to the A' products
  const int cws = dm.wpr + 1;
  const size_t nr = (size_t)n * r, mr = (size_t)m * r;

  // ---- stage-local copies: row ids, B, column-major A
  const double bsc = *tk.bscale;
  const int rank_one = tk.rank_one_ptr ? *tk.rank_one_ptr : tk.rank_one;
  for (int i = tid; i < m; i += NT) {
    sm.rows_s[i] = tk.A.rows ? tk.A.rows[i] : i;
    sm.Bs[i] = bsc * tk.B[tk.brows ? tk.brows[i] : i];
  }
  __syncthreads();
  if constexpr (!CM) {   // (with codes nothing below reads the column-major copy)
    for (size_t idx = tid; idx < (size_t)m * n; idx += NT) {
      const int k = (int)(idx % n), i = (int)(idx / n);
      cd a = Ab[(size_t)sm.rows_s[i] * n + k];
      ws.Acm[i + (size_t)m * k] = cmk(a.x * asc_dense, a.y * asc_dense);
    }
  }
  if constexpr (CM) {
    for (int idx = tid; idx < m * dm.wpr; idx += NT) {
      const int i = idx / dm.wpr, w = idx - i * dm.wpr;
      sm.cw[i * cws + w] = tk.codes[(size_t)sm.rows_s[i] * dm.wpr + w];
    }
  }
  double nb2;
  {
    double v[1] = {0.0};
    for (int i = tid; i < m; i += NT) v[0] += sm.Bs[i] * sm.Bs[i];
    block_sum<1>(v, sm.red);   // (also orders the Acm writes before the reads below)
    nb2 = v[0];
  }
  const double normB = sqrt(nb2);
  // device cycle counters of the phases (thread 0; scal[9..14], the layout of the cluster kernels)
  long long cyc_x = 0, cyc_y = 0, cyc_z = 0, cyc_loop = 0;
  const long long t_begin = clock64();
  if (tid == 0) sm.sc[30] = 0.0;

  // ---- S = I + A A' (Woodbury) or A'A + I, then its inverse (inferLowRankV4.m:221 / :267)
  if (wood) {
    if constexpr (CM) {
      // u_i(k) conj(u_j(k)) = j^((c_i - c_j) mod 4): the 16 differences of a word pair come from one 2-bit-field
      // subtraction, their counts from three popcounts; the sum over k is exact integer arithmetic
      constexpr uint32_t HB = 0xAAAAAAAAu, LB = 0x55555555u;
      const double as2 = asc * asc;
      for (int idx = tid; idx < m * m; idx += NT) {
        const int i = idx % m, j = idx / m;
        if (i >= j) {
          const uint32_t* wi = sm.cw + i * cws;
          const uint32_t* wj = sm.cw + j * cws;
          int re = 0, im = 0;
          for (int w = 0; w < dm.wpr; ++w) {
            const uint32_t x = wi[w], y = wj[w];
            const uint32_t d = ((x | HB) - (y & LB)) ^ ((x ^ ~y) & HB);
            const uint32_t hi = (d >> 1) & LB, lo = d & LB;
            const int n3 = __popc(hi & lo), n2 = __popc(hi & ~lo), n1 = __popc(lo & ~hi);
            re += 16 - n1 - 2 * n2 - n3;     // n0 - n2
            im += n1 - n3;
          }
          ws.Sinv[idx] = cmk(fma(as2, (double)re, i == j ? 1.0 : 0.0), as2 * (double)im);
        }
      }
    } else {
      for (int idx = tid; idx < m * m; idx += NT) {
        const int i = idx % m, j = idx / m;
        cd acc = cmk(i == j ? 1.0 : 0.0, 0.0);
        if (i >= j) {
          for (int k = 0; k < n; ++k) cfmabc(acc, ws.Acm[i + (size_t)m * k], ws.Acm[j + (size_t)m * k]);
          ws.Sinv[idx] = acc;
        }
      }
    }
    __syncthreads();
    for (int idx = tid; idx < m * m; idx += NT) {
      const int i = idx % m, j = idx / m;
      if (i < j) { cd u = ws.Sinv[j + (size_t)m * i]; ws.Sinv[idx] = cmk(u.x, -u.y); }
      else if (i == j) ws.Sinv[idx].y = 0.0;
    }
  } else {
    const double as2 = asc_dense * asc_dense;
    for (size_t idx = tid; idx < (size_t)n * n; idx += NT) {
      const int k = (int)(idx % n), l = (int)(idx / n);
      cd acc = cmk(0.0, 0.0);
      for (int i = 0; i < m; ++i) {
        const cd* row = Ab + (size_t)sm.rows_s[i] * n;
        cfmac(acc, row[k], row[l]);
      }
      ws.Sinv[idx] = cmk(fma(acc.x, as2, (k == l) ? 1.0 : 0.0), acc.y * as2);
    }
  }
  __syncthreads();
  spd_inverse(ws.Sinv, dS, sm.big, sm.big + dS);

  // ---- X = X0, M = N = 0
  for (size_t idx = tid; idx < nr; idx += NT) { ws.X[idx] = tk.X0[idx]; ws.N[idx] = cmk(0.0, 0.0); }
  for (size_t idx = tid; idx < mr; idx += NT) ws.M[idx] = cmk(0.0, 0.0);
  __syncthreads();

  auto ucode = [&](int i, int k) -> unsigned { return (sm.cw[i * cws + (k >> 4)] >> (2 * (k & 15))) & 3u; };
  auto matA = [&](int k, int i) -> cd {                                                  // A(i,k), out=i
    if constexpr (CM) {
      const unsigned c = ucode(i, k);
      const double v = (c & 2u) ? -asc : asc;
      return (c & 1u) ? cmk(0.0, v) : cmk(v, 0.0);
    } else {
      return ws.Acm[i + (size_t)m * k];
    }
  };
  auto matAh = [&](int i, int k) -> cd {                                                 // conj(A(i,k))/scale, out=k
    if constexpr (CM) {
      const unsigned c = ucode(i, k);
      const double v = (c & 2u) ? -1.0 : 1.0;
      return (c & 1u) ? cmk(0.0, -v) : cmk(v, 0.0);
    } else {
      cd a = Ab[(size_t)sm.rows_s[i] * n + k];
      return cmk(a.x, -a.y);
    }
  };

  // AX = A * X                                                                   (:278)
  gemm_tpo(m, n, r, matA, [&](int k, int c) -> cd { return ws.X[k + (size_t)n * c]; },
           [&](int i, int c, cd v) { ws.AX[i + (size_t)m * c] = v; }, sm.tile, sm.big);
  // rescale X so |A X| matches |B|                                              (:279-286)
  if (tk.sbr) {
    double v[1] = {0.0};
    for (size_t idx = tid; idx < mr; idx += NT) v[0] += cabs2(ws.AX[idx]);
    block_sum<1>(v, sm.red);
    const double s = normB / sqrt(v[0]);
    for (size_t idx = tid; idx < nr; idx += NT) ws.X[idx] = cscale(ws.X[idx], s);
    for (size_t idx = tid; idx < mr; idx += NT) ws.AX[idx] = cscale(ws.AX[idx], s);
  } else {
    for (int c = warp; c < r; c += NW) {
      double a = 0.0;
      for (int i = lane; i < m; i += 32) a += cabs2(ws.AX[i + (size_t)m * c]);
      a = warp_sum(a);
      if (lane == 0) sm.colsc[c] = normB / sqrt(a);
    }
    __syncthreads();
    for (size_t idx = tid; idx < nr; idx += NT) ws.X[idx] = cscale(ws.X[idx], sm.colsc[idx / n]);
    for (size_t idx = tid; idx < mr; idx += NT) ws.AX[idx] = cscale(ws.AX[idx], sm.colsc[idx / m]);
  }
  __syncthreads();
  // Y = normalize_rows(AX, B)                                                    (:287, :517-538)
  if (tk.sbr) {
    const double isr = 1.0 / sqrt((double)r);
    for (int i = tid; i < m; i += NT) {
      double d2 = 0.0;
      for (int c = 0; c < r; ++c) d2 += cabs2(ws.AX[i + (size_t)m * c]);
      double D = sqrt(d2);
      const bool z = (D == 0.0);
      if (z) D = 1.0;
      const double f = sm.Bs[i] / D;
      for (int c = 0; c < r; ++c) {
        cd v = z ? cmk(isr, 0.0) : ws.AX[i + (size_t)m * c];
        ws.Y[i + (size_t)m * c] = cscale(v, f);
      }
    }
  } else {
    for (size_t idx = tid; idx < mr; idx += NT) {
      const int i = (int)(idx % m);
      cd v = ws.AX[idx];
      double D = sqrt(cabs2(v));
      if (D == 0.0) { v = cmk(1.0, 0.0); D = 1.0; }
      ws.Y[idx] = cscale(v, sm.Bs[i] / D);
    }
  }
  __syncthreads();
  int sweeps = 0;
  double nz[4];
  if (tid == 0) sm.ifl[2] = 0;
  if (dm.tx == 16) jacobi_tables<16>(sm.jtab);
  else if (dm.tx == 32) jacobi_tables<32>(sm.jtab);
  __syncthreads();
  // Z = ArgMinZ(X, N=0, mu=1)                                                    (:288)
  if (tk.nuclear) argmin_z_nuclear(tk, dm, ws, sm, 1.0, true, nz, &sweeps);
  else argmin_z_v4(tk, dm, ws, sm, 1.0, true, rank_one, nz, &sweeps);
  __syncthreads();
  cd* AtY = ws.AtY0;
  cd* AtYp = ws.AtY1;
  if (prm.need_dual) {   // AtY = A' * Y                                          (:289)
    gemm_tpo(n, m, r, matAh, [&](int i, int c) -> cd { return ws.Y[i + (size_t)m * c]; },
             [&](int k, int c, cd v) { AtY[k + (size_t)n * c] = cscale(v, asc); }, sm.tile, sm.big);
  }

  double mu = prm.mu0, opt_obj = INFINITY, last_res = INFINITY, res_comb = 0.0;
  int iters = 0, opt_iter = -1, opt_col = -1, bumps = 0, converged = 0, have_opt = 0;
  const long long t_loop = clock64();

  for (int it = 1; it <= prm.maxiter; ++it) {
    const double imu = 1.0 / mu;
    const long long t0 = clock64();
    // ---- X update (:304, :380-388):  X = inv(A'A+I) (A'T + Q),  T = Y - M/mu,  Q = Z - N/mu
    if (wood) {
      // two-product Woodbury form:  X = Q + A' W,  A X = T - W,  W = S^-1 (T - A Q),  S = I + A A'
      for (size_t idx = tid; idx < nr; idx += NT) {
        const cd z = ws.Z[idx], nn = ws.N[idx];
        ws.X[idx] = cmk(fma(-nn.x, imu, z.x), fma(-nn.y, imu, z.y));
      }
      __syncthreads();
      gemm_tpo(m, n, r, matA, [&](int k, int c) -> cd { return ws.X[k + (size_t)n * c]; },
               [&](int i, int c, cd v) {
                 const size_t p = i + (size_t)m * c;
                 const cd y = ws.Y[p], mm = ws.M[p];
                 ws.Wb[p] = cmk(fma(-mm.x, imu, y.x) - v.x, fma(-mm.y, imu, y.y) - v.y);
               },
               sm.tile, sm.big);
      gemm_tpo(m, m, r, [&](int j, int i) -> cd { return ws.Sinv[i + (size_t)m * j]; },
               [&](int j, int c) -> cd { return ws.Wb[j + (size_t)m * c]; },
               [&](int i, int c, cd v) { ws.AX[i + (size_t)m * c] = v; }, sm.tile, sm.big);   // W (for now)
      gemm_tpo(n, m, r, matAh, [&](int i, int c) -> cd { return ws.AX[i + (size_t)m * c]; },
               [&](int k, int c, cd v) {
                 const size_t p = k + (size_t)n * c;
                 const cd x = ws.X[p];
                 ws.X[p] = cmk(fma(v.x, asc, x.x), fma(v.y, asc, x.y));
               },
               sm.tile, sm.big);
      for (size_t idx = tid; idx < mr; idx += NT) {     // A X = T - W
        const cd y = ws.Y[idx], mm = ws.M[idx], w = ws.AX[idx];
        ws.AX[idx] = cmk(fma(-mm.x, imu, y.x) - w.x, fma(-mm.y, imu, y.y) - w.y);
      }
      __syncthreads();
    } else {
      gemm_tpo(n, m, r, matAh,
               [&](int i, int c) -> cd {
                 cd y = ws.Y[i + (size_t)m * c], mm = ws.M[i + (size_t)m * c];
                 return cmk(fma(-mm.x, imu, y.x), fma(-mm.y, imu, y.y));
               },
               [&](int k, int c, cd v) {
                 const size_t p = k + (size_t)n * c;
                 cd z = ws.Z[p], nn = ws.N[p];
                 ws.Vb[p] = cmk(fma(v.x, asc, fma(-nn.x, imu, z.x)), fma(v.y, asc, fma(-nn.y, imu, z.y)));
               },
               sm.tile, sm.big);
      gemm_tpo(n, n, r, [&](int l, int k) -> cd { return ws.Sinv[k + (size_t)n * l]; },
               [&](int l, int c) -> cd { return ws.Vb[l + (size_t)n * c]; },
               [&](int k, int c, cd v) { ws.X[k + (size_t)n * c] = v; }, sm.tile, sm.big);
      gemm_tpo(m, n, r, matA, [&](int k, int c) -> cd { return ws.X[k + (size_t)n * c]; },
               [&](int i, int c, cd v) { ws.AX[i + (size_t)m * c] = v; }, sm.tile, sm.big);
    }
    // ---- Y update (:308, :490-512), M update (:315-316), objective (:323-340), Y/AX norms
    const long long t1 = clock64();
    cyc_x += t1 - t0;
    double nYd2 = 0.0, nJM2 = 0.0, nAX2 = 0.0, nY2 = 0.0, obj2 = 0.0;
    const double i1mu = 1.0 / (1.0 + mu);
    if (tk.sbr) {
      const double isr = 1.0 / sqrt((double)r);
      for (int i = tid; i < m; i += NT) {
        double d2 = 0.0;
        for (int c = 0; c < r; ++c) {
          cd ax = ws.AX[i + (size_t)m * c], mm = ws.M[i + (size_t)m * c];
          d2 += cabs2(cmk(fma(mm.x, imu, ax.x), fma(mm.y, imu, ax.y)));
        }
        double D = sqrt(d2);
        const bool z = (D == 0.0);
        if (z) D = 1.0;
        const double f = (sm.Bs[i] / D + mu) * i1mu;
        double rax2 = 0.0;
        for (int c = 0; c < r; ++c) {
          const size_t p = i + (size_t)m * c;
          cd ax = ws.AX[p], mm = ws.M[p], yo = ws.Y[p];
          cd cc = z ? cmk(isr, 0.0) : cmk(fma(mm.x, imu, ax.x), fma(mm.y, imu, ax.y));
          cd yn = cscale(cc, f);
          cd jm = cmk(ax.x - yn.x, ax.y - yn.y);
          ws.Y[p] = yn;
          ws.M[p] = cmk(fma(mu, jm.x, mm.x), fma(mu, jm.y, mm.y));
          nYd2 += cabs2(cmk(yn.x - yo.x, yn.y - yo.y));
          nJM2 += cabs2(jm);
          rax2 += cabs2(ax);
          nY2 += cabs2(yn);
        }
        nAX2 += rax2;
        const double dd = sqrt(rax2) - sm.Bs[i];
        obj2 += dd * dd;
      }
    } else {
      for (int c = warp; c < r; c += NW) {
        double oc = 0.0;
        for (int i = lane; i < m; i += 32) {
          const size_t p = i + (size_t)m * c;
          cd ax = ws.AX[p], mm = ws.M[p], yo = ws.Y[p];
          cd cc = cmk(fma(mm.x, imu, ax.x), fma(mm.y, imu, ax.y));
          double D = sqrt(cabs2(cc));
          if (D == 0.0) { cc = cmk(1.0, 0.0); D = 1.0; }
          const double f = (sm.Bs[i] / D + mu) * i1mu;
          cd yn = cscale(cc, f);
          cd jm = cmk(ax.x - yn.x, ax.y - yn.y);
          ws.Y[p] = yn;
          ws.M[p] = cmk(fma(mu, jm.x, mm.x), fma(mu, jm.y, mm.y));
          nYd2 += cabs2(cmk(yn.x - yo.x, yn.y - yo.y));
          nJM2 += cabs2(jm);
          const double a2 = cabs2(ax);
          nAX2 += a2;
          nY2 += cabs2(yn);
          const double dd = sqrt(a2) - sm.Bs[i];
          oc += dd * dd;
        }
        oc = warp_sum(oc);
        if (lane == 0) sm.colsc[c] = sqrt(oc);
      }
    }
    {
      double v[5] = {nYd2, nJM2, nAX2, nY2, obj2};
      block_sum<5>(v, sm.red);
      nYd2 = v[0]; nJM2 = v[1]; nAX2 = v[2]; nY2 = v[3]; obj2 = v[4];
    }
    // ---- AtY = A' * Y (:309) — only feeds res_dual
    double nAtYd2 = 0.0, nAtY2 = 0.0;
    if (prm.need_dual) {
      cd* t = AtY; AtY = AtYp; AtYp = t;
      gemm_tpo(n, m, r, matAh, [&](int i, int c) -> cd { return ws.Y[i + (size_t)m * c]; },
               [&](int k, int c, cd v) { AtY[k + (size_t)n * c] = cscale(v, asc); }, sm.tile, sm.big);
      double v[2] = {0.0, 0.0};
      for (size_t idx = tid; idx < nr; idx += NT) {
        cd a = AtY[idx], b = AtYp[idx];
        v[0] += cabs2(cmk(a.x - b.x, a.y - b.y));
        v[1] += cabs2(a);
      }
      block_sum<2>(v, sm.red);
      nAtYd2 = v[0]; nAtY2 = v[1];
    }
    // ---- Z update (:312), N update (:319-320), X/Z norms
    const long long t2 = clock64();
    cyc_y += t2 - t1;
    if (tk.nuclear) argmin_z_nuclear(tk, dm, ws, sm, mu, false, nz, &sweeps);
    else argmin_z_v4(tk, dm, ws, sm, mu, false, rank_one, nz, &sweeps);
    cyc_z += clock64() - t2;
    const double nJN2 = nz[0], nZd2 = nz[1], nX2 = nz[2], nZ2 = nz[3];

    // ---- best solution so far (:323-340).  NaN objectives never win (MATLAB min skips NaN).
    double obj; int jbest = -1;
    if (tk.sbr) {
      obj = sqrt(obj2);
    } else {
      obj = NAN;
      for (int c = 0; c < r; ++c) {
        const double oc = sm.colsc[c];
        if (oc == oc && (jbest < 0 || oc < obj)) { obj = oc; jbest = c; }
      }
    }
    if (obj < opt_obj) {   // uniform across the block: every thread holds the same scalars
      opt_obj = obj; opt_iter = it; opt_col = jbest; have_opt = 1;
      if (tk.sbr) {
        for (size_t idx = tid; idx < nr; idx += NT) ws.optX[idx] = ws.X[idx];
        for (size_t idx = tid; idx < mr; idx += NT) ws.optY[idx] = ws.Y[idx];
      } else {
        for (int k = tid; k < n; k += NT) ws.optX[k] = ws.X[k + (size_t)n * jbest];
        for (int i = tid; i < m; i += NT) ws.optY[i] = ws.Y[i + (size_t)m * jbest];
      }
    }
    // ---- residuals and stopping rule (:343-354)
    const double res_prim = sqrt(nJM2 + nJN2);
    const double res_dual = mu * sqrt(nAtYd2 + nZd2);
    res_comb = sqrt(nJM2 + nJN2 + nYd2 + nZd2);
    iters = it;
    if (tk.trace != nullptr && threadIdx.x == 0) tk.trace[it - 1] = res_comb;
    if (prm.need_dual) {
      const double mx1 = fmax(sqrt(nAX2), sqrt(nY2)), mx2 = fmax(sqrt(nX2), sqrt(nZ2));
      const double th_prim = prm.tol_abs * sqrt((double)(m + n) * r) + prm.tol_rel * sqrt(mx1 * mx1 + mx2 * mx2);
      const double th_dual = prm.tol_abs * sqrt((double)n * r * 2.0) + prm.tol_rel * sqrt(nAtY2 + nZ2);
      const double th_comb = prm.tol_abs * sqrt((double)(m + n) * r * 2.0) +
                             prm.tol_rel * sqrt(mx1 * mx1 + mx2 * mx2 + nY2 + nZ2);
      if ((res_prim < th_prim && res_dual < th_dual) || (res_comb < th_comb)) { converged = 1; break; }
    }
    // ---- mu adaptation (:358-361)
    if (res_comb > last_res * 0.9) { mu *= prm.rho; ++bumps; }
    last_res = res_comb;
    __syncthreads();
  }
  __syncthreads();
  cyc_loop = clock64() - t_loop;
  // ---- outputs (:363-364)
  const int rout = tk.sbr ? r : 1;
  for (size_t idx = tid; idx < (size_t)n * rout; idx += NT)
    if (tk.Xout) tk.Xout[idx] = have_opt ? ws.optX[idx] : cmk(NAN, NAN);
  for (size_t idx = tid; idx < (size_t)m * rout; idx += NT)
    if (tk.Yout) tk.Yout[idx] = have_opt ? ws.optY[idx] : cmk(NAN, NAN);
  if (tk.state) {
    cd* st = tk.state;
    for (size_t idx = tid; idx < nr; idx += NT) { st[idx] = ws.X[idx]; st[nr + idx] = ws.Z[idx]; st[2 * nr + idx] = ws.N[idx]; }
    for (size_t idx = tid; idx < mr; idx += NT) { st[3 * nr + idx] = ws.Y[idx]; st[3 * nr + mr + idx] = ws.M[idx]; }
  }
  if (tk.scal && tid == 0) {
    tk.scal[SC_MU] = mu; tk.scal[SC_OPT_OBJ] = opt_obj; tk.scal[SC_ITERS] = iters;
    tk.scal[SC_OPT_ITER] = opt_iter; tk.scal[SC_OPT_COL] = opt_col; tk.scal[SC_BUMPS] = bumps;
    tk.scal[SC_CONVERGED] = converged; tk.scal[SC_RES_COMB] = res_comb; tk.scal[SC_SWEEPS] = sweeps;
    tk.scal[9] = sm.sc[30]; tk.scal[10] = (double)cyc_x; tk.scal[11] = (double)cyc_loop; tk.scal[12] = (double)cyc_y;
    tk.scal[13] = (double)cyc_z; tk.scal[14] = (double)(t_loop - t_begin);     // [14]: set-up (S, its inverse, first Z)
  }
  __syncthreads();
}

// The kernel lives in its own translation unit (gen_kernel.cu; api.cu reaches it through these host functions): its two
// instantiations of run_stage are half of the library's compile time, and the translation units build in parallel.
cudaError_t gen_kernel_set_smem(size_t smem);
cudaError_t gen_kernel_occupancy(int* per_sm, size_t smem);
cudaError_t gen_kernel_launch(int grid, size_t smem, cudaStream_t stream, const StageTask* tasks, int ntasks, DevParams prm,
                              StageDims dm, cd* wsbase);

#ifdef TWOACE_GEN_KERNEL_TU
__global__ void __launch_bounds__(NT, 2)
admm_stage_kernel(const StageTask* __restrict__ tasks, int ntasks, DevParams prm, StageDims dm, cd* wsbase) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const StageSmem sm = carve_smem(smem_raw, dm);
  const StageWS ws = carve_ws(wsbase + (size_t)blockIdx.x * dm.ws_stride, dm);
  for (int t = blockIdx.x; t < ntasks; t += gridDim.x) {
    const StageTask tk = tasks[t];
    if (tk.active != nullptr && *tk.active != tk.active_expect) continue;
    if (tk.m <= 0 || tk.r <= 0) continue;
    if (dm.wpr > 0 && tk.codes != nullptr && tk.cscale != nullptr) run_stage<true>(tk, prm, dm, ws, sm);
    else run_stage<false>(tk, prm, dm, ws, sm);
  }
}

cudaError_t gen_kernel_set_smem(size_t smem) {
  return cudaFuncSetAttribute(admm_stage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}
cudaError_t gen_kernel_occupancy(int* per_sm, size_t smem) {
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, admm_stage_kernel, NT, smem);
}
cudaError_t gen_kernel_launch(int grid, size_t smem, cudaStream_t stream, const StageTask* tasks, int ntasks, DevParams prm,
                              StageDims dm, cd* wsbase) {
  admm_stage_kernel<<<grid, NT, smem, stream>>>(tasks, ntasks, prm, dm, wsbase);
  return cudaGetLastError();
}
#endif

}  // namespace twoace
