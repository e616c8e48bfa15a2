// Fast InferADMM kernel for the 16x16 (n = 256) quantised-codebook case: the whole iterate lives in
// shared memory and the r columns of X are split over a thread-block cluster (CS CTAs, RL = r/CS
// columns each).  Cross-CTA coupling is only
//   * the row norms of A X + M/mu (scale_by_row),
//   * the tx x tx Gram E E' of ArgMinZ,
//   * a handful of residual norms / per-column objectives,
// exchanged through distributed shared memory; the A-products are fully local.
//
// The sensing matrix is held as 2-bit phase codes: every shipped codebook entry is a 4th root of
// unity, A_eff = cscale * u, u in {1, j, -1, -j}; two packed copies ([i-word][k] and
// [k-word][i], 16 codes per 32-bit word) serve the A'(.) and A(.) products without a transpose.
// Same algorithm as admm_stage.cuh (inferLowRankV4.m:260-365); differences, exact in exact arithmetic:
//   * two-product Woodbury form of ArgMinX: X = Q + A'W, A X = T - W, W = S^-1 (T - A Q) with Q = Z - N/mu,
//     T = Y - M/mu, S = I + A A' (from exact integer phase counts, S^-1 streamed from L2);
//   * N is overwritten by Z_in = X + N/mu during ArgMinZ and rebuilt as N = mu (Z_in - Z)
//     (identical to N + mu (X - Z), :319-320);
//   * A'Y (:309) lives in global memory and is advanced by A'(Y - Y0), only when a tolerance is set;
//   * the best iterate (:323-340) is written straight to the output buffers.
#pragma once
#include <cooperative_groups.h>

#include "admm_stage.cuh"
#include "tc_prod.cuh"

namespace cg = cooperative_groups;

namespace twoace {

constexpr int FN = 256;    // n
constexpr int FTX = 16;    // tx (= rx)
constexpr int XS_SCAL = 16;

struct FastDims {
  int maxm;          // largest m in the launch (big_stage.cuh: rows per chunk)
  int mfull;         // big_stage.cuh: largest m in the launch (0 otherwise)
  int mw;            // ceil(maxm/16): words per k of the [i-word][k] code copy
  int r;             // total columns (CS * RL)
  int ds;            // shared eig dimension: 16 (V4: tx x tx) or max(16, r) (nuclear: r x r)
  int nuclear;
  int lean;          // big1_stage_kernel: no code copies, no row-exchange buffers in shared memory
  size_t ws_stride;  // global workspace elements (cd) per cluster
  TcDims tc;         // tensor-core products (tc.on == 0: FP64 SIMT products)
};

__host__ __device__ inline size_t fast_ws_elems(const FastDims& d) {
  // Sinv | AtY | operand blocks of the tensor-core products (streaming mode)
  return (size_t)d.maxm * d.maxm + (size_t)FN * d.r + (d.tc.on ? TC_AOP_BYTES / sizeof(cd) : 0);
}

template <int RL>
struct FastSmem {
  cd *X, *Z, *N, *Y, *M, *WT, *AX, *G, *U, *P;
  cd* xG;           // exchange: Gram partial
  cd* lut;          // [0..3] u(code), [4..7] conj(u(code))
  JacobiScratch js;
  unsigned char* pairs;   // [15][8][2] round-robin pair table of jacobi16
  double* jprm;           // [8][4] rotation parameters of the current round
  double* Bs;
  double* xrow;     // exchange: [2*maxm] row partial sums
  double* rowtot;   // [2*maxm] cluster totals
  double* xsc;      // exchange: [2][XS_SCAL] scalars (parity double-buffered)
  double* xcol;     // exchange: [2][SMALL_DMAX] per-column objectives (owner writes its slots)
  double *red, *s2s, *colsc, *sc;
  uint32_t* cki;    // [mw][256]   code(i = 16w + j, k) in bits 2j of cki[w*256 + k]   (SIMT products only)
  uint32_t* cik;    // [16][m]     code(i, k = 16w + j) in bits 2j of cik[w*m + i]     (tensor-core mode: setup only,
                    //             aliases the B operand)
  int* rows_s;
  int* ifl;
  // tensor-core products
  unsigned char* tc_bs;    // B operand
  unsigned char* tc_ov;    // overlay region (== G): ring slots [0, n1)
  unsigned char* tc_ex;    // ring slots [n1, nslot)
  uint64_t* tc_bars;
  uint32_t* tc_tslot;
  size_t bytes;            // total carved
};

// Shared-memory layout of the fast kernel; fast_smem_bytes() is the size of the same carve.
// G | P | xG are contiguous (the "overlay": dead outside ArgMinZ, so the tensor-core products use it as ring slots).
template <int RL>
__host__ __device__ inline FastSmem<RL> fast_carve(unsigned char* base, const FastDims& d) {
  FastSmem<RL> s;
  unsigned char* p = base;
  auto take = [&p](size_t bytes, size_t align) {
    size_t a = (size_t)p;
    a = (a + align - 1) / align * align;
    unsigned char* q = (unsigned char*)a;
    p = q + bytes;
    return q;
  };
  s.X = (cd*)take((size_t)FN * RL * sizeof(cd), 16);
  s.Z = (cd*)take((size_t)FN * RL * sizeof(cd), 16);
  s.N = (cd*)take((size_t)FN * RL * sizeof(cd), 16);
  s.Y = (cd*)take((size_t)d.maxm * RL * sizeof(cd), 16);
  s.M = (cd*)take((size_t)d.maxm * RL * sizeof(cd), 16);
  s.WT = (cd*)take((size_t)d.maxm * RL * sizeof(cd), 16);
  s.AX = (cd*)take((size_t)d.maxm * RL * sizeof(cd), 16);
  const size_t gpx = 3 * (size_t)d.ds * d.ds * sizeof(cd);
  const size_t ovb = d.tc.on ? (size_t)d.tc.ov_bytes : gpx;
  s.tc_ov = take(ovb, 128);
  s.G = (cd*)s.tc_ov;
  s.P = s.G + (size_t)d.ds * d.ds;
  s.xG = s.P + (size_t)d.ds * d.ds;
  s.U = (cd*)take((size_t)d.ds * d.ds * sizeof(cd), 16);   // persistent across iterations (warm start)
  s.pairs = take(2048, 16);   // JacobiTab<16 or 20>::BYTES: pair / element tables + rotation parameters
  s.jprm = (double*)take(32 * sizeof(double), 16);
  s.lut = (cd*)take(8 * sizeof(cd), 16);
  const int h = d.ds / 2 + 2;
  s.js.e = (cd*)take((size_t)h * sizeof(cd), 16);
  s.js.cs = (double*)take((size_t)h * sizeof(double), 8);
  s.js.sn = (double*)take((size_t)h * sizeof(double), 8);
  const size_t mb = (size_t)(d.mfull > d.maxm ? d.mfull : d.maxm);
  s.Bs = (double*)take(mb * sizeof(double), 8);
  s.xrow = (double*)take((size_t)2 * (d.lean ? 16 : d.maxm) * sizeof(double), 8);
  s.rowtot = (double*)take((size_t)2 * (d.lean ? 16 : d.maxm) * sizeof(double), 8);
  s.xsc = (double*)take(2 * XS_SCAL * sizeof(double), 8);
  s.xcol = (double*)take(2 * SMALL_DMAX * sizeof(double), 8);
  s.red = (double*)take(16 * NW * sizeof(double), 8);
  s.s2s = (double*)take(SMALL_DMAX * sizeof(double), 8);
  s.colsc = (double*)take(2 * SMALL_DMAX * sizeof(double), 8);
  s.sc = (double*)take(32 * sizeof(double), 8);
  s.rows_s = (int*)take(mb * sizeof(int), 4);
  s.ifl = (int*)take(48 * sizeof(int), 4);
  s.js.flag = s.ifl + 8;
  s.js.gscale = s.sc + 31;
  const size_t cik_bytes = (size_t)16 * d.maxm * 4;
  if (d.lean) {
    s.cki = nullptr; s.cik = nullptr;
    s.tc_bs = nullptr; s.tc_ex = nullptr; s.tc_bars = nullptr; s.tc_tslot = nullptr;
  } else if (d.tc.on) {
    s.cki = nullptr;
    s.tc_bs = take(cik_bytes > (size_t)TC_BS_BYTES ? cik_bytes : (size_t)TC_BS_BYTES, 128);
    s.cik = (uint32_t*)s.tc_bs;
    s.tc_bars = (uint64_t*)take((2 * TC_MAXSLOT + 1) * sizeof(uint64_t), 8);
    s.tc_tslot = (uint32_t*)take(16, 4);
    // the ring slots outside the overlay; 2 KB of slack behind them: a K-major read of a block with fewer than
    // 128 rows per slab runs past its last slab (the rows it reads there land in unused accumulator lanes)
    s.tc_ex = take((size_t)(d.tc.nslot - d.tc.n1) * d.tc.slot_bytes + 2048, 128);
  } else {
    s.cki = (uint32_t*)take((size_t)d.mw * 256 * 4, 16);
    s.cik = (uint32_t*)take(cik_bytes, 16);
    s.tc_bs = nullptr; s.tc_ex = nullptr; s.tc_bars = nullptr; s.tc_tslot = nullptr;
  }
  s.bytes = (size_t)(p - base);
  return s;
}

template <int RL>
__host__ __device__ inline size_t fast_smem_bytes(const FastDims& d) {
  // carve from a 1024-aligned dummy base (the kernel's dynamic shared memory is declared with that alignment)
  return fast_carve<RL>((unsigned char*)(uintptr_t)1024, d).bytes + 128;
}

// Ring geometry of the tensor-core products for a launch with the given maxm: the overlay region holds n1 slots,
// the rest of the ring takes what is left of `limit` bytes of shared memory.  Returns false when no slot fits.
template <int RL>
__host__ inline bool fast_tc_layout(FastDims& d, size_t limit) {
  const int mt = (d.maxm + 127) / 128;
  if (mt > 2) return false;
  const int R = mt > 1 ? 128 : (d.maxm + 31) / 32 * 32;
  const int nblk = 4 * mt;
  const size_t gpx = 3 * (size_t)d.ds * d.ds * sizeof(cd);
  d.tc.on = 1;
  d.tc.slot_bytes = R * 128;
  // try the overlay grown to one slot when the Gram buffers alone are smaller
  for (int grow = 0; grow < 2; ++grow) {
    size_t ov = gpx;
    if (grow && ov < (size_t)d.tc.slot_bytes) ov = d.tc.slot_bytes;
    ov = (ov + 127) / 128 * 128;
    d.tc.ov_bytes = (int)ov;
    d.tc.n1 = (int)std::min<size_t>(ov / d.tc.slot_bytes, (size_t)nblk);
    d.tc.nslot = d.tc.n1;
    const size_t base = fast_smem_bytes<RL>(d);
    if (base > limit) continue;
    // (up to nblk slots behind the overlay: with all of them the blocks stay resident for the whole stage)
    const int extra = (int)std::min<size_t>((limit - base) / d.tc.slot_bytes, (size_t)nblk);
    d.tc.nslot = std::min(d.tc.n1 + extra, TC_MAXSLOT);
    if (d.tc.nslot >= 1 && (grow || d.tc.nslot >= 2)) return true;
  }
  d.tc.on = 0;
  return d.tc.on != 0;
}

template <int CS>
__device__ __forceinline__ void cl_sync() {
  if constexpr (CS > 1) cg::this_cluster().sync();
  else __syncthreads();
}

template <class T, int CS>
__device__ __forceinline__ T* peer_ptr(T* p, int rank) {
  if constexpr (CS > 1) return cg::this_cluster().map_shared_rank(p, rank);
  else return p;
}

// Next task index from a device-wide counter, the same value in every CTA of the cluster (two cluster barriers).
template <int CS>
__device__ __forceinline__ int next_task(int* counter, int* slot, int rank) {
  if (rank == 0 && threadIdx.x == 0) *slot = atomicAdd(counter, 1);
  cl_sync<CS>();
  const int t = *peer_ptr<int, CS>(slot, 0);
  cl_sync<CS>();      // rank 0 may overwrite its slot only after every CTA has read it
  return t;
}

// A' product, register-tiled 2 (k) x RL (columns) with the reduction over i split over a lane pair:
//   thread t: s = t & 1, k0 = t >> 1, k1 = k0 + 128, rows i = s, s + 2, s + 4, ...
// Every operand Op[i, c] loaded from shared memory feeds 8 DFMAs (a broadcast LDS.128 costs 4 cycles of the
// shared-memory pipe per warp, so one load per 4 DFMAs made the products shared-memory bound).  After the
// pair reduction lane s keeps the result for k = k0 + 128 s, returned in acc[]; kout = that k.
template <int RL>
__device__ __forceinline__ int prod_ah(const uint32_t* cki, int m, const cd* Op, int ldo, const cd* lutc,
                                       cd (&acc)[RL]) {
  const int tid = threadIdx.x;
  const int s = tid & 1, k0 = tid >> 1, k1 = k0 + 128;
  cd a0[RL], a1[RL];
#pragma unroll
  for (int c = 0; c < RL; ++c) { a0[c] = cmk(0.0, 0.0); a1[c] = cmk(0.0, 0.0); }
  for (int w = 0; w * 16 < m; ++w) {
    uint32_t w0 = cki[w * 256 + k0] >> (2 * s), w1 = cki[w * 256 + k1] >> (2 * s);
    const int cnt = (min(16, m - w * 16) - s + 1) >> 1;   // rows 16w + s, 16w + s + 2, ... below m
    const cd* op = Op + w * 16 + s;
#pragma unroll 2
    for (int q = 0; q < cnt; ++q) {
      const cd u0 = lutc[w0 & 3u], u1 = lutc[w1 & 3u];
      w0 >>= 4;
      w1 >>= 4;
#pragma unroll
      for (int c = 0; c < RL; ++c) {
        const cd t = op[2 * q + ldo * c];
        cfma(a0[c], u0, t);
        cfma(a1[c], u1, t);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < RL; ++c) {
    // lane s keeps k = k0 + 128 s: it needs its own partial of that k plus the partner's
    const cd mine = s ? a1[c] : a0[c];
    const cd give = s ? a0[c] : a1[c];
    acc[c].x += mine.x + __shfl_xor_sync(0xffffffffu, give.x, 1);
    acc[c].y += mine.y + __shfl_xor_sync(0xffffffffu, give.y, 1);
  }
  return k0 + 128 * s;
}

// largest power of two ks <= 16 with ks * m <= NT: the reduction index is split over ks ADJACENT lanes
// (interleaved: lane s takes indices congruent to s mod ks, so shared-memory reads of consecutive lanes hit
// consecutive 16-byte words) and the partial sums are combined with warp shuffles.  ks depends on m only,
// so the summation order of an instance never depends on its batch mates.
__device__ __forceinline__ int ksplit_for(int m, int kmax) {
  int ks = 1;
  while (ks < kmax && 2 * ks * m <= NT) ks *= 2;
  return ks;
}

template <int RL>
__device__ __forceinline__ void group_reduce(cd (&acc)[RL], int ks) {
  for (int o = 1; o < ks; o <<= 1) {
#pragma unroll
    for (int c = 0; c < RL; ++c) {
      acc[c].x += __shfl_xor_sync(0xffffffffu, acc[c].x, o);
      acc[c].y += __shfl_xor_sync(0xffffffffu, acc[c].y, o);
    }
  }
}

// store(i, c, sum_k u(i,k) * V[k + 256*c])   (A product), register-tiled 2 (rows i) x RL (columns):
// thread (row pair ip, lane-group member s) handles rows 2 ip, 2 ip + 1 and k = s, s + ks, ...
template <int RL, class StoreF>
__device__ __forceinline__ void prod_a(const uint32_t* cik, int m, const cd* V, const cd* lut, StoreF store) {
  const int tid = threadIdx.x;
  const int mp = (m + 1) >> 1;               // row pairs
  const int ks = ksplit_for(mp, 16);
  const int ip = tid / ks, s = tid - ip * ks;
  const bool act = ip < mp;
  const int i0 = 2 * ip, i1 = min(2 * ip + 1, m - 1);
  cd a0[RL], a1[RL];
#pragma unroll
  for (int c = 0; c < RL; ++c) { a0[c] = cmk(0.0, 0.0); a1[c] = cmk(0.0, 0.0); }
  if (act) {
    const int per = 16 / ks;   // codes of each word handled by this lane
    for (int w = 0; w < 16; ++w) {
      const uint32_t w0 = cik[w * m + i0] >> (2 * s), w1 = cik[w * m + i1] >> (2 * s);
      const cd* v = V + w * 16 + s;
#pragma unroll 2
      for (int t = 0; t < per; ++t) {
        const cd u0 = lut[(w0 >> (2 * ks * t)) & 3u], u1 = lut[(w1 >> (2 * ks * t)) & 3u];
#pragma unroll
        for (int c = 0; c < RL; ++c) {
          const cd x = v[ks * t + FN * c];
          cfma(a0[c], u0, x);
          cfma(a1[c], u1, x);
        }
      }
    }
  }
  group_reduce<RL>(a0, ks);
  group_reduce<RL>(a1, ks);
  if (act && s == 0) {
#pragma unroll
    for (int c = 0; c < RL; ++c) store(i0, c, a0[c]);
    if (2 * ip + 1 < m) {
#pragma unroll
      for (int c = 0; c < RL; ++c) store(i1, c, a1[c]);
    }
  }
  __syncthreads();
}

// W = Sinv * R (R = T - A Q in `RW`, overwritten by W) and AX = T - W (T in `TAX`, overwritten by AX).
// Sinv lives in global memory / L2 (m^2 x 16 B per iteration: the largest stream of the kernel).  Register tile of
// NR = 4 (2 for RL = 10) rows x RL columns per thread: every R operand read from shared memory (a broadcast LDS.128)
// feeds NR complex FMAs and the Sinv elements of a thread are NR x 16 contiguous bytes.  The reduction index j is split over ks
// adjacent lanes (shuffle reduction); the Sinv values of the next two j are loaded while the current two are
// multiplied, so eight independent L2 loads per thread are always in flight.
template <int RL, int NR>
__device__ __forceinline__ void prod_sinv_t(const cd* __restrict__ Sinv, int m, cd* RW, cd* TAX) {
  const int tid = threadIdx.x;
  const int mq = (m + NR - 1) / NR;
  const int ks = ksplit_for(mq, 16);
  const int iq = tid / ks, s = tid - iq * ks;
  const bool act = iq < mq;
  int row[NR];
#pragma unroll
  for (int u = 0; u < NR; ++u) row[u] = min(NR * iq + u, m - 1);
  cd acc[NR][RL];
#pragma unroll
  for (int u = 0; u < NR; ++u)
#pragma unroll
    for (int c = 0; c < RL; ++c) acc[u][c] = cmk(0.0, 0.0);
  if (act) {
    const int nj = (m - s + ks - 1) / ks;           // this lane's j: s, s + ks, ...
    cd cur[2][NR], nxt[2][NR];
    auto fetch = [&](cd (&dst)[2][NR], int q) {      // j indices q, q + 1 of this lane (clamped: the tail is masked below)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int j = s + ks * min(q + h, nj - 1);
#pragma unroll
        for (int u = 0; u < NR; ++u) dst[h][u] = __ldg(Sinv + row[u] + (size_t)m * j);
      }
    };
    fetch(cur, 0);
    for (int q = 0; q < nj; q += 2) {
      if (q + 2 < nj) fetch(nxt, q + 2);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (q + h < nj) {
          const int j = s + ks * (q + h);
#pragma unroll
          for (int c = 0; c < RL; ++c) {
            const cd r = RW[j + m * c];
#pragma unroll
            for (int u = 0; u < NR; ++u) cfma(acc[u][c], cur[h][u], r);
          }
        }
      }
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int u = 0; u < NR; ++u) cur[h][u] = nxt[h][u];
    }
  }
#pragma unroll
  for (int u = 0; u < NR; ++u) group_reduce<RL>(acc[u], ks);
  __syncthreads();            // every thread has finished reading R before it is overwritten by W
  if (act && s == 0) {
#pragma unroll
    for (int u = 0; u < NR; ++u) {
      if (NR * iq + u < m) {
#pragma unroll
        for (int c = 0; c < RL; ++c) {
          const int p = NR * iq + u + m * c;
          const cd t = TAX[p];
          RW[p] = acc[u][c];
          TAX[p] = cmk(t.x - acc[u][c].x, t.y - acc[u][c].y);
        }
      }
    }
  }
  __syncthreads();
}

// Rows per thread by problem size: the wider tiles pay for their shuffle reduction only when m is large (measured
// on B200, cycles per call at m = 30 / 60 / 121 / 243:  NR 1: ~6 k / 7 k / 19 k / 62 k,  NR 4: 8.4 k / 12.4 k / 20 k / 51 k).
template <int RL>
__device__ __forceinline__ void prod_sinv(const cd* __restrict__ Sinv, int m, cd* RW, cd* TAX) {
  if (m > 128) prod_sinv_t<RL, (RL > 5 ? 2 : 4)>(Sinv, m, RW, TAX);
  else if (m > 64) prod_sinv_t<RL, 2>(Sinv, m, RW, TAX);
  else prod_sinv_t<RL, 1>(Sinv, m, RW, TAX);
}

// ArgMinZ (inferLowRankV4.m:402-464) + N update (:319-320) + X/Z norms.  Contains exactly one
// cluster sync.  On return nrm[0..3] hold THIS CTA's partial |X-Z|^2, |Z-Z0|^2, |X|^2, |Z|^2.
template <int RL, int CS>
__device__ inline void fast_argmin_z(int m, int rank_one, const FastSmem<RL>& sm, double mu, bool init_mode,
                                     bool warm, double* nrm, int* sweeps_acc) {
  const int tid = threadIdx.x;
  const double imu = 1.0 / mu;
  constexpr int NE = FN * RL;          // local elements
  constexpr int NEC = NE / FTX;        // local E columns
  // ---- N <- Z_in = X + N/mu
  for (int idx = tid; idx < NE; idx += NT) {
    cd v = sm.X[idx];
    if (!init_mode) { const cd nn = sm.N[idx]; v.x = fma(nn.x, imu, v.x); v.y = fma(nn.y, imu, v.y); }
    sm.N[idx] = v;
  }
  __syncthreads();
  // ---- partial Gram over the local E columns (lower triangle)
  const int gi = tid & 15, gj = tid >> 4;   // 256 threads = the 16 x 16 entries
  {
    cd acc = cmk(0.0, 0.0);
    if (gi >= gj) {   // 4 independent accumulators: a dependent DFMA costs ~23 cycles
      cd a0 = cmk(0.0, 0.0), a1 = a0, a2 = a0, a3 = a0;
      static_assert(NEC % 4 == 0, "NEC must be a multiple of 4");
#pragma unroll 2
      for (int e = 0; e < NEC; e += 4) {
        cfmabc(a0, sm.N[gi + FTX * e], sm.N[gj + FTX * e]);
        cfmabc(a1, sm.N[gi + FTX * (e + 1)], sm.N[gj + FTX * (e + 1)]);
        cfmabc(a2, sm.N[gi + FTX * (e + 2)], sm.N[gj + FTX * (e + 2)]);
        cfmabc(a3, sm.N[gi + FTX * (e + 3)], sm.N[gj + FTX * (e + 3)]);
      }
      acc = cmk((a0.x + a1.x) + (a2.x + a3.x), (a0.y + a1.y) + (a2.y + a3.y));
    }
    sm.xG[gi + FTX * gj] = acc;
  }
  cl_sync<CS>();
  if (gi >= gj) {
    cd g = cmk(0.0, 0.0);
#pragma unroll
    for (int rk = 0; rk < CS; ++rk) {
      const cd v = peer_ptr<cd, CS>(sm.xG, rk)[gi + FTX * gj];
      g.x += v.x;
      g.y += v.y;
    }
    if (gi == gj) g.y = 0.0;
    sm.G[gi + FTX * gj] = g;
    if (gi != gj) sm.G[gj + FTX * gi] = cmk(g.x, -g.y);
  }
  __syncthreads();
  if (warm) {   // rotate into the previous eigenbasis: G <- U' G U (nearly diagonal), then refine U
    cd t1 = cmk(0.0, 0.0), t1b = t1, t1c = t1, t1d = t1;
#pragma unroll
    for (int k = 0; k < FTX; k += 4) {
      cfma(t1, sm.G[gi + FTX * k], sm.U[k + FTX * gj]);
      cfma(t1b, sm.G[gi + FTX * (k + 1)], sm.U[(k + 1) + FTX * gj]);
      cfma(t1c, sm.G[gi + FTX * (k + 2)], sm.U[(k + 2) + FTX * gj]);
      cfma(t1d, sm.G[gi + FTX * (k + 3)], sm.U[(k + 3) + FTX * gj]);
    }
    sm.P[gi + FTX * gj] = cmk((t1.x + t1b.x) + (t1c.x + t1d.x), (t1.y + t1b.y) + (t1c.y + t1d.y));
    __syncthreads();
    cd t2 = cmk(0.0, 0.0);
    if (gi >= gj) {
      cd t2b = t2, t2c = t2, t2d = t2;
#pragma unroll
      for (int k = 0; k < FTX; k += 4) {
        cfmac(t2, sm.U[k + FTX * gi], sm.P[k + FTX * gj]);
        cfmac(t2b, sm.U[(k + 1) + FTX * gi], sm.P[(k + 1) + FTX * gj]);
        cfmac(t2c, sm.U[(k + 2) + FTX * gi], sm.P[(k + 2) + FTX * gj]);
        cfmac(t2d, sm.U[(k + 3) + FTX * gi], sm.P[(k + 3) + FTX * gj]);
      }
      t2 = cmk((t2.x + t2b.x) + (t2c.x + t2d.x), (t2.y + t2b.y) + (t2c.y + t2d.y));
      if (gi == gj) t2.y = 0.0;
    }
    __syncthreads();
    if (gi >= gj) {
      sm.G[gi + FTX * gj] = t2;
      if (gi != gj) sm.G[gj + FTX * gi] = cmk(t2.x, -t2.y);
    }
    __syncthreads();
  }
  // ---- cheap exact screen (warm only): by Schur-Horn the r largest diagonal entries of U'GU sum to at most
  // the r largest eigenvalues, while the trace is the same.  If every constraint C(r_k, f_k) of :449-459
  // already holds for the sorted DIAGONAL (with a 1e-12 margin) it holds for the spectrum, no stage fires,
  // s2_scale == 1 everywhere and :461 leaves Z = Z_in: the eigen-decomposition is not needed at all.
  bool need_eig = true;
  if (warm) {
    if (tid < FTX) sm.colsc[tid] = fmax(0.0, sm.G[tid + FTX * tid].x);
    __syncthreads();
    if (tid < FTX) {
      const double v = sm.colsc[tid];
      int rank = 0;
#pragma unroll
      for (int j = 0; j < FTX; ++j) { const double o = sm.colsc[j]; rank += (o > v) || (o == v && j < tid); }
      sm.colsc[FTX + rank] = v;
    }
    __syncthreads();
    if (tid == 0) {
      double pre[FTX + 1];
      pre[0] = 0.0;
#pragma unroll
      for (int i = 0; i < FTX; ++i) pre[i + 1] = pre[i] + sm.colsc[FTX + i];
      int rl[4]; double fl[4];
      const int ns = rank_profile_dev(FTX, FTX, m, FN, rank_one, rl, fl);
      int ok = 1;
      for (int k = 0; k < ns; ++k) ok &= (pre[min(rl[k], FTX)] >= pre[FTX] * fl[k] * (1.0 + 1e-12)) ? 1 : 0;
      sm.ifl[1] = ok;
    }
    __syncthreads();
    need_eig = sm.ifl[1] == 0;
  }
  int any = 0;
  if (need_eig) {
  const long long tj0 = clock64();
  const int sw = jacobi_small<FTX>(sm.G, sm.P, sm.U, sm.pairs, !warm);
  if (tid == 0) sm.sc[20] += (double)(clock64() - tj0);
  // eigenvalues clamped (:408) and ranked in descending order, stable (:409): one thread per eigenvalue
  if (tid < FTX) sm.colsc[tid] = fmax(0.0, sm.G[tid + FTX * tid].x);
  __syncthreads();
  if (tid < FTX) {
    const double v = sm.colsc[tid];
    int rank = 0;
#pragma unroll
    for (int j = 0; j < FTX; ++j) { const double o = sm.colsc[j]; rank += (o > v) || (o == v && j < tid); }
    sm.colsc[FTX + rank] = v;          // sorted values
    sm.ifl[16 + tid] = rank;           // position of eigenvalue tid in the sorted order
  }
  __syncthreads();
  if (tid == 0) {
    *sweeps_acc += sw;
    // prefix sums of the sorted spectrum (tree-ish: 4 chains) then the cascade of :449-459 in O(1) per
    // stage: with cf = product of the scales applied so far (to every index >= r_prev),
    //   head_k = head_{k-1} + cf (pre[r_k] - pre[r_{k-1}]),  v = head_k + cf (pre[16] - pre[r_k])
    double pre[FTX + 1];
    pre[0] = 0.0;
#pragma unroll
    for (int i = 0; i < FTX; ++i) pre[i + 1] = pre[i] + sm.colsc[FTX + i];
    int rl[4]; double fl[4];
    const int ns = rank_profile_dev(FTX, FTX, m, FN, rank_one, rl, fl);
    double cf = 1.0, head = 0.0;
    int rprev = 0;
    double cfs[4];
    for (int k = 0; k < ns; ++k) {
      const int rr = min(rl[k], FTX);
      head += cf * (pre[rr] - pre[rprev]);
      const double v = head + cf * (pre[FTX] - pre[rr]);
      if (head < v * fl[k]) cf *= fmin(1.0, head / (v - head) * (1.0 / fl[k] - 1.0));
      cfs[k] = cf;
      rprev = rr;
    }
    // s2_scale of sorted position pos: product of the scales of the stages with r_k <= pos
    for (int pos = 0; pos < FTX; ++pos) {
      double sc = 1.0;
      for (int k = 0; k < ns; ++k) if (pos >= rl[k]) sc = cfs[k];
      sm.colsc[2 * FTX + pos] = sc;
    }
    sm.ifl[0] = (cf < 1.0) ? 1 : 0;
  }
  __syncthreads();
  if (tid < FTX) sm.s2s[tid] = sqrt(sm.colsc[2 * FTX + sm.ifl[16 + tid]]);
  __syncthreads();
  any = sm.ifl[0];
  if (any) {   // P = U diag(sqrt(s2_scale)) U'
    cd ac[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) ac[u] = cmk(0.0, 0.0);
#pragma unroll
    for (int k = 0; k < FTX; ++k) {
      const cd ui = sm.U[gi + FTX * k], uj = sm.U[gj + FTX * k];
      const double s = sm.s2s[k];
      cfmabc(ac[k & 3], cmk(ui.x * s, ui.y * s), uj);
    }
    sm.P[gi + FTX * gj] = cmk((ac[0].x + ac[1].x) + (ac[2].x + ac[3].x), (ac[0].y + ac[1].y) + (ac[2].y + ac[3].y));
  }
  __syncthreads();
  if (tid == 0) sm.ifl[2] += 1;   // eigen-decompositions actually performed
  }   // need_eig
  // ---- Z <- P Z_in (or Z_in), norms.  Item (e, ig): rows 4ig..4ig+3 of E column e; the owner of an
  // element is the only thread touching Z there, so Z_old can be read and replaced in place.
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  constexpr int ITEMS = NEC * 4;
  for (int it = tid; it < ITEMS; it += NT) {
    const int ig = it & 3, e = it >> 2;
    const cd* z = sm.N + FTX * e;
    cd zn[4];
    if (any) {
#pragma unroll
      for (int u = 0; u < 4; ++u) zn[u] = cmk(0.0, 0.0);
#pragma unroll 4
      for (int k = 0; k < FTX; ++k) {
        const cd zz = z[k];
#pragma unroll
        for (int u = 0; u < 4; ++u) cfma(zn[u], sm.P[(4 * ig + u) + FTX * k], zz);
      }
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u) zn[u] = z[4 * ig + u];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = FTX * e + 4 * ig + u;
      if (!init_mode) {
        const cd zo = sm.Z[idx], x = sm.X[idx];
        a0 += cabs2(cmk(x.x - zn[u].x, x.y - zn[u].y));
        a1 += cabs2(cmk(zn[u].x - zo.x, zn[u].y - zo.y));
        a2 += cabs2(x);
        a3 += cabs2(zn[u]);
      }
      sm.Z[idx] = zn[u];
    }
  }
  __syncthreads();
  // ---- N <- mu (Z_in - Z)   (0 in init mode: N is not part of the :288 call)
  for (int idx = tid; idx < NE; idx += NT) {
    if (init_mode) {
      sm.N[idx] = cmk(0.0, 0.0);
    } else {
      const cd zi = sm.N[idx], zn = sm.Z[idx];
      sm.N[idx] = cmk(mu * (zi.x - zn.x), mu * (zi.y - zn.y));
    }
  }
  if (!init_mode) {
    double v[4] = {a0, a1, a2, a3};
    block_sum<4>(v, sm.red);
    nrm[0] = v[0]; nrm[1] = v[1]; nrm[2] = v[2]; nrm[3] = v[3];
  } else {
    __syncthreads();
  }
}

// Nuclear-norm ArgMinZ (inferLowRank_Nuclear.m:411-439) on the column-split iterate: singular-value soft
// threshold of Z_in (n x r) by tau = 1/mu through the r x r Gram eigenproblem,
//   Z = Z_in V diag(max(0, s - tau)/s) V',   Z_in' Z_in = V diag(s^2) V'.
// Columns of Z_in living in peer CTAs are staged through `stage` (cap elements) in chunks of whole columns.
// Exact screen: ||Z_in||_F <= tau  =>  every s_i <= tau  =>  Z = 0 (the usual case while mu is small).
// Cluster syncs: 1 (screen hit) or 3 (SVT), cluster-uniform.  nrm[] = this CTA's partial norms.
template <int RL, int CS>
__device__ inline void fast_argmin_z_nuclear(const FastDims& fd, const FastSmem<RL>& sm, int rank, double mu,
                                             bool init_mode, double* xsc_slot, cd* stage, size_t stage_cap,
                                             double* nrm, int* sweeps_acc) {
  const int tid = threadIdx.x;
  constexpr int r = RL * CS;
  const int c0 = rank * RL;
  const double imu = 1.0 / mu;
  constexpr int NE = FN * RL;
  double fro = 0.0;
  for (int idx = tid; idx < NE; idx += NT) {
    cd v = sm.X[idx];
    if (!init_mode) { const cd nn = sm.N[idx]; v.x = fma(nn.x, imu, v.x); v.y = fma(nn.y, imu, v.y); }
    sm.N[idx] = v;
    fro += cabs2(v);
  }
  {
    double v[1] = {fro};
    block_sum<1>(v, sm.red);
    if (tid == 0) *xsc_slot = v[0];
  }
  cl_sync<CS>();
  double tot = 0.0;
#pragma unroll
  for (int rk = 0; rk < CS; ++rk) tot += *peer_ptr<double, CS>(xsc_slot, rk);
  const bool all_zero = !(tot > imu * imu);      // ||Z_in||_F <= tau (NaN: falls through to the SVT)
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  if (all_zero && tot == tot) {
    for (int idx = tid; idx < NE; idx += NT) {
      if (!init_mode) {
        const cd zo = sm.Z[idx], x = sm.X[idx], zi = sm.N[idx];
        a0 += cabs2(x); a1 += cabs2(zo); a2 += cabs2(x);
        sm.N[idx] = cmk(mu * zi.x, mu * zi.y);
      } else {
        sm.N[idx] = cmk(0.0, 0.0);
      }
      sm.Z[idx] = cmk(0.0, 0.0);
    }
  } else {
    (void)stage; (void)stage_cap;   // (remote columns are read in place; no staging buffer any more)
    // ---- Gram rows of the local columns: G[c0 + cl, c'] = sum_k conj(z[k, c0+cl]) z[k, c'] for ALL c' (local and
    // remote).  One warp per column c': lane l takes k = l, l + 32, ...; the column -- read straight from the
    // owner's shared memory (DSMEM), every element exactly once -- is multiplied with the RL local columns, then
    // the RL partial sums are reduced over the warp.  No staging buffer, no block barrier.
    {
      const int lane = tid & 31, warp = tid >> 5;
      for (int col = warp; col < r; col += NW) {
        const int own = col / RL;
        const cd* zc = (own == rank ? sm.N : peer_ptr<cd, CS>(sm.N, own)) + FN * (col - own * RL);
        cd acc[RL];
#pragma unroll
        for (int cl = 0; cl < RL; ++cl) acc[cl] = cmk(0.0, 0.0);
#pragma unroll 2
        for (int k = lane; k < FN; k += 32) {
          const cd z = zc[k];
#pragma unroll
          for (int cl = 0; cl < RL; ++cl) cfmac(acc[cl], sm.N[k + FN * cl], z);
        }
#pragma unroll
        for (int cl = 0; cl < RL; ++cl) {
          const double gx = warp_sum(acc[cl].x), gy = warp_sum(acc[cl].y);
          if (lane == 0) sm.xG[(c0 + cl) + r * col] = cmk(gx, gy);
        }
      }
    }
    cl_sync<CS>();      // every CTA's row block of G is complete
    for (int idx = tid; idx < r * r; idx += NT) {
      const int i = idx % r, j = idx / r;
      const cd v = peer_ptr<cd, CS>(sm.xG, i / RL)[i + r * j];   // row i is owned by rank i / RL
      sm.G[idx] = v;
    }
    __syncthreads();
    for (int idx = tid; idx < r * r; idx += NT) {   // exact Hermitian symmetry from the lower triangle
      const int i = idx % r, j = idx / r;
      if (i == j) sm.G[idx].y = 0.0;
    }
    __syncthreads();
    for (int idx = tid; idx < r * r; idx += NT) {
      const int i = idx % r, j = idx / r;
      if (i < j) { const cd u = sm.G[j + r * i]; sm.G[idx] = cmk(u.x, -u.y); }
    }
    __syncthreads();
    int sw;
    if constexpr (r == 20) {
      // warm start from the previous eigenvectors (cold every 64 decompositions re-orthonormalises U)
      const bool warm = !init_mode && (sm.ifl[2] & 63) != 0;
      if (warm) {
        for (int idx = tid; idx < r * r; idx += NT) {     // P <- G U
          const int i = idx % r, j = idx / r;
          cd t0 = cmk(0.0, 0.0), t1 = t0;
          for (int k = 0; k < r; k += 2) {
            cfma(t0, sm.G[i + r * k], sm.U[k + r * j]);
            cfma(t1, sm.G[i + r * (k + 1)], sm.U[(k + 1) + r * j]);
          }
          sm.P[idx] = cmk(t0.x + t1.x, t0.y + t1.y);
        }
        __syncthreads();
        for (int idx = tid; idx < r * r; idx += NT) {     // G <- U' P (lower triangle, mirrored)
          const int i = idx % r, j = idx / r;
          if (i >= j) {
            cd t0 = cmk(0.0, 0.0), t1 = t0;
            for (int k = 0; k < r; k += 2) {
              cfmac(t0, sm.U[k + r * i], sm.P[k + r * j]);
              cfmac(t1, sm.U[(k + 1) + r * i], sm.P[(k + 1) + r * j]);
            }
            cd t = cmk(t0.x + t1.x, t0.y + t1.y);
            if (i == j) t.y = 0.0;
            sm.G[i + r * j] = t;
            if (i != j) sm.G[j + r * i] = cmk(t.x, -t.y);
          }
        }
        __syncthreads();
      }
      // (skipping the rotations inside the sub-threshold cluster was tried: it slows convergence, measured.)
      // Absolute rotation floor 1e-15 max|diag| instead of 1e-18: the explicitly formed Gram Z'Z carries
      // rounding noise of ~1e-16 max|diag| in every entry, so the numerically null cluster of a low-rank Z
      // can never be resolved below that; chasing it costs ~3 extra sweeps per call and changes the retained
      // singular values by at most r * 1e-15 relative to the largest one.
      const long long tj0 = clock64();
      sw = jacobi_small_p<20>(sm.G, sm.P, sm.U, sm.pairs, !warm, 30, nullptr, 0.0, 1.0e-15);
      if (tid == 0) { sm.ifl[2] += 1; sm.sc[20] += (double)(clock64() - tj0); }
    } else {
      sw = jacobi_heig(sm.G, r, sm.U, r, r, true, sm.js);
    }
    if (tid == 0) *sweeps_acc += sw;
    if (tid < r) {
      const double sg = sqrt(fmax(0.0, sm.G[tid + r * tid].x));
      sm.s2s[tid] = (sg > 0.0) ? fmax(0.0, sg - imu) / sg : 0.0;
    }
    __syncthreads();
    for (int idx = tid; idx < r * r; idx += NT) {   // P = V diag(f) V'
      const int i = idx % r, j = idx / r;
      cd p0 = cmk(0.0, 0.0), p1 = p0;
      for (int k = 0; k + 1 < r; k += 2) {
        const cd ui = sm.U[i + r * k], uj = sm.U[j + r * k];
        const double f0 = sm.s2s[k];
        cfmabc(p0, cmk(ui.x * f0, ui.y * f0), uj);
        const cd vi = sm.U[i + r * (k + 1)], vj = sm.U[j + r * (k + 1)];
        const double f1 = sm.s2s[k + 1];
        cfmabc(p1, cmk(vi.x * f1, vi.y * f1), vj);
      }
      if (r & 1) {
        const cd ui = sm.U[i + r * (r - 1)], uj = sm.U[j + r * (r - 1)];
        const double f0 = sm.s2s[r - 1];
        cfmabc(p0, cmk(ui.x * f0, ui.y * f0), uj);
      }
      sm.P[idx] = cmk(p0.x + p1.x, p0.y + p1.y);
    }
    __syncthreads();
    // ---- Z[:, c0 + c'] = sum_c Z_in[:, c] P[c, c0 + c']   (thread per row k; accumulators in registers)
    cd acc[RL];
#pragma unroll
    for (int c = 0; c < RL; ++c) acc[c] = cmk(0.0, 0.0);
    {
      const int k = tid;
      for (int c = 0; c < RL; ++c) {
        const cd z = sm.N[k + FN * c];
#pragma unroll
        for (int c2 = 0; c2 < RL; ++c2) cfma(acc[c2], z, sm.P[(c0 + c) + r * (c0 + c2)]);
      }
    }
    if constexpr (CS > 1) {
      // remote columns straight from their owner's shared memory (coalesced over k, every element read once)
      for (int pr = 1; pr < CS; ++pr) {
        const int rk = (rank + pr) % CS;
        const cd* rem = peer_ptr<cd, CS>(sm.N, rk);
        cd zr[RL];
#pragma unroll
        for (int cc = 0; cc < RL; ++cc) zr[cc] = rem[tid + FN * cc];
#pragma unroll
        for (int cc = 0; cc < RL; ++cc) {
#pragma unroll
          for (int c2 = 0; c2 < RL; ++c2) cfma(acc[c2], zr[cc], sm.P[(rk * RL + cc) + r * (c0 + c2)]);
        }
      }
    }
    cl_sync<CS>();      // peers are done reading this CTA's Z_in before it is overwritten
    {
      const int k = tid;
#pragma unroll
      for (int c = 0; c < RL; ++c) {
        const int idx = k + FN * c;
        const cd zn = acc[c];
        if (!init_mode) {
          const cd zo = sm.Z[idx], x = sm.X[idx], zi = sm.N[idx];
          a0 += cabs2(cmk(x.x - zn.x, x.y - zn.y));
          a1 += cabs2(cmk(zn.x - zo.x, zn.y - zo.y));
          a2 += cabs2(x);
          a3 += cabs2(zn);
          sm.N[idx] = cmk(mu * (zi.x - zn.x), mu * (zi.y - zn.y));
        } else {
          sm.N[idx] = cmk(0.0, 0.0);
        }
        sm.Z[idx] = zn;
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

template <int RL, int CS, bool TC>
__device__ inline void run_fast(const StageTask& tk, const DevParams& prm, const FastDims& fd,
                                const FastSmem<RL>& sm, cd* wsg, int rank, TcCtx& tc) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m = tk.m;
  constexpr int r = RL * CS;
  const int c0 = rank * RL;              // first global column owned by this CTA
  const double cs = *tk.cscale;          // A_eff = cs * u
  const double bsc = *tk.bscale;
  const int rank_one = tk.rank_one_ptr ? *tk.rank_one_ptr : tk.rank_one;
  cd* Sinv = tk.sinv != nullptr ? tk.sinv : wsg;      // (I + A A')^-1: per-instance store or cluster workspace
  const bool sinv_reuse = tk.sinv != nullptr && tk.sinv_state == 1;
  const long long ttask0 = clock64();
  cd* AtY = wsg + (size_t)fd.maxm * fd.maxm;   // [256 x r], column c0+c owned by this CTA
  const cd* lut = sm.lut;
  const cd* lutc = sm.lut + 4;
  cd* scratch = sm.WT;                   // WT | AX (pivot row / column of the Gauss-Jordan inverse)
  TcGeom tg = {};
  if constexpr (TC) {
    tg = tc_geom(m, tc.nslot_launch, tc.n1);
    tc.nslot = tc.nslot_launch;
    tc.wt_slot = -1;
    if (!tg.resident) {
      // streaming: one producer warp per slot (warps 1 .. NW-1); WT serves as one more slot while it is dead
      // (single-pass products only: with RL = 10 the second column pass still needs WT)
      tc.nslot = min(tc.nslot, NW - 2);
      if (RL == TC_NC && (size_t)fd.maxm * RL * sizeof(cd) >= (size_t)tc.slot_bytes) { tc.wt_slot = tc.nslot; tc.nslot += 1; }
    }
    tc.premask = 0;
    tc.aop = (unsigned char*)(wsg + (size_t)fd.maxm * fd.maxm + (size_t)FN * fd.r);
  }
  // store(i, c, sum_k u(i,k) V[k + 256 c]), c < RL: FP64 SIMT or exact int8 tensor-core product
  auto product_a = [&](const cd* V, auto store) {
    if constexpr (TC) {
      for (int cb = 0; cb < RL; cb += TC_NC)
        tc_product<false>(tc, tg, m, [&](int k, int c) { return V[k + FN * (cb + c)]; },
                          [&](int i, int c, cd v) { store(i, cb + c, v); }, (uint32_t*)sm.red);
    } else {
      prod_a<RL>(sm.cik, m, V, lut, store);
    }
  };
  // store(k, c, sum_i conj(u(i,k)) Op[i + m c]), every (k, c) exactly once, by the thread that owns it
  auto product_ah = [&](const cd* Op, auto store) {
    if constexpr (TC) {
      for (int cb = 0; cb < RL; cb += TC_NC)
        tc_product<true>(tc, tg, m, [&](int i, int c) { return Op[i + m * (cb + c)]; },
                         [&](int k, int c, cd v) { store(k, cb + c, v); }, (uint32_t*)sm.red);
    } else {
      cd acc[RL];
#pragma unroll
      for (int c = 0; c < RL; ++c) acc[c] = cmk(0.0, 0.0);
      const int kk = prod_ah<RL>(sm.cki, m, Op, m, lutc, acc);
#pragma unroll
      for (int c = 0; c < RL; ++c) store(kk, c, acc[c]);
    }
  };

  // ---- stage-local copies
  if (tid < 4) {
    const double re[4] = {1.0, 0.0, -1.0, 0.0}, im[4] = {0.0, 1.0, 0.0, -1.0};
    sm.lut[tid] = cmk(re[tid], im[tid]);
    sm.lut[4 + tid] = cmk(re[tid], -im[tid]);
  }
  for (int i = tid; i < m; i += NT) {
    sm.rows_s[i] = tk.A.rows ? tk.A.rows[i] : i;
    sm.Bs[i] = bsc * tk.B[tk.brows ? tk.brows[i] : i];
  }
  if (fd.nuclear && fd.r == 20) jacobi_tables<20>(sm.pairs); else jacobi_tables<FTX>(sm.pairs);
  __syncthreads();
  for (int idx = tid; idx < 16 * m; idx += NT) {
    const int w = idx / m, i = idx - w * m;
    sm.cik[w * m + i] = tk.codes[(size_t)sm.rows_s[i] * 16 + w];
  }
  __syncthreads();
  if constexpr (!TC) {
    const int k = tid;   // thread k packs code(i, k) for all i
    for (int w = 0; w * 16 < m; ++w) {
      uint32_t word = 0;
      const int cnt = min(16, m - w * 16);
      for (int j = 0; j < cnt; ++j) {
        const uint32_t c = (sm.cik[(k >> 4) * m + (w * 16 + j)] >> (2 * (k & 15))) & 3u;
        word |= c << (2 * j);
      }
      sm.cki[w * 256 + k] = word;
    }
  }
  double nb2;
  {
    double v[1] = {0.0};
    for (int i = tid; i < m; i += NT) v[0] += sm.Bs[i] * sm.Bs[i];
    block_sum<1>(v, sm.red);
    nb2 = v[0];
  }
  const double normB = sqrt(nb2);

  // ---- S = I + A A' from exact phase counts (pairs split over the cluster), inverse by rank 0; skipped when an
  // earlier stage of the same trial has left the inverse in the per-instance store (same rows, same scale)
  {
    const double cs2 = cs * cs;
    if (!sinv_reuse)
    for (int idx = tid + NT * rank; idx < m * m; idx += NT * CS) {
      const int i = idx % m, j = idx / m;
      if (i >= j) {
        int re = 0, im = 0;
        for (int w = 0; w < 16; ++w) {
          uint32_t wi = sm.cik[w * m + i], wj = sm.cik[w * m + j];
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const uint32_t d = (wi - wj) & 3u;     // u_i conj(u_j) = j^d
            re += (d == 0u) - (d == 2u);
            im += (d == 1u) - (d == 3u);
            wi >>= 2;
            wj >>= 2;
          }
        }
        const cd v = cmk(((i == j) ? 1.0 : 0.0) + cs2 * re, (i == j) ? 0.0 : cs2 * im);
        Sinv[i + (size_t)m * j] = v;
        if (i != j) Sinv[j + (size_t)m * i] = cmk(v.x, -v.y);
      }
    }
    // operand blocks of the tensor-core products (the code copy `cik` is dead after this: it aliases the B operand)
    if constexpr (TC) tc_build(tc, tg, sm.cik, m, tg.resident ? 0 : rank, tg.resident ? 1 : CS);
    __threadfence();
    cl_sync<CS>();
    if (!sinv_reuse) spd_inverse_part(Sinv, m, scratch, scratch + m, rank, CS, [] { __threadfence(); cl_sync<CS>(); });
    __threadfence();
    cl_sync<CS>();
  }

  // ---- X = X0 (own columns), M = N = 0
  for (int idx = tid; idx < FN * RL; idx += NT) {
    sm.X[idx] = tk.X0[(size_t)FN * c0 + idx];
    sm.N[idx] = cmk(0.0, 0.0);
  }
  for (int idx = tid; idx < m * RL; idx += NT) sm.M[idx] = cmk(0.0, 0.0);
  __syncthreads();
  // AX = A X  (:278)
  product_a(sm.X, [&](int i, int c, cd v) { sm.AX[i + m * c] = cscale(v, cs); });
  // rescale so |A X| matches |B|  (:279-286)
  if (tk.sbr) {
    double v[1] = {0.0};
    for (int idx = tid; idx < m * RL; idx += NT) v[0] += cabs2(sm.AX[idx]);
    block_sum<1>(v, sm.red);
    if (tid == 0) sm.xsc[0] = v[0];
    cl_sync<CS>();
    double tot = 0.0;
#pragma unroll
    for (int rk = 0; rk < CS; ++rk) tot += peer_ptr<double, CS>(sm.xsc, rk)[0];
    const double s = normB / sqrt(tot);
    for (int idx = tid; idx < FN * RL; idx += NT) sm.X[idx] = cscale(sm.X[idx], s);
    for (int idx = tid; idx < m * RL; idx += NT) sm.AX[idx] = cscale(sm.AX[idx], s);
    cl_sync<CS>();   // peers are done with xsc[0] before it is reused
  } else {
    for (int c = warp; c < RL; c += NW) {
      double a = 0.0;
      for (int i = lane; i < m; i += 32) a += cabs2(sm.AX[i + m * c]);
      a = warp_sum(a);
      if (lane == 0) sm.colsc[c] = normB / sqrt(a);
    }
    __syncthreads();
    for (int idx = tid; idx < FN * RL; idx += NT) sm.X[idx] = cscale(sm.X[idx], sm.colsc[idx / FN]);
    for (int idx = tid; idx < m * RL; idx += NT) sm.AX[idx] = cscale(sm.AX[idx], sm.colsc[idx / m]);
  }
  __syncthreads();
  // Y = normalize_rows(AX, B)  (:287, :517-538)
  if (tk.sbr) {
    for (int i = tid; i < m; i += NT) {
      double d2 = 0.0;
      for (int c = 0; c < RL; ++c) d2 += cabs2(sm.AX[i + m * c]);
      sm.xrow[i] = d2;
    }
    cl_sync<CS>();
    const double isr = 1.0 / sqrt((double)r);
    for (int i = tid; i < m; i += NT) {
      double d2 = 0.0;
#pragma unroll
      for (int rk = 0; rk < CS; ++rk) d2 += peer_ptr<double, CS>(sm.xrow, rk)[i];
      double D = sqrt(d2);
      const bool z = (D == 0.0);
      if (z) D = 1.0;
      const double f = sm.Bs[i] / D;
      for (int c = 0; c < RL; ++c) {
        const cd v = z ? cmk(isr, 0.0) : sm.AX[i + m * c];
        sm.Y[i + m * c] = cscale(v, f);
      }
    }
    cl_sync<CS>();
  } else {
    for (int idx = tid; idx < m * RL; idx += NT) {
      const int i = idx % m;
      cd v = sm.AX[idx];
      double D = sqrt(cabs2(v));
      if (D == 0.0) { v = cmk(1.0, 0.0); D = 1.0; }
      sm.Y[idx] = cscale(v, sm.Bs[i] / D);
    }
  }
  __syncthreads();
  int sweeps = 0;
  double nz[4];
  if (tid == 0) sm.ifl[2] = 0;
  __syncthreads();
  // Z = ArgMinZ(X, 0, 1)  (:288).  Nuclear staging area: WT | AX (both dead during ArgMinZ)
  cd* stage = sm.WT;
  const size_t stage_cap = 2 * (size_t)fd.maxm * RL;
  if (fd.nuclear) fast_argmin_z_nuclear<RL, CS>(fd, sm, rank, 1.0, true, sm.xsc + XS_SCAL - 1, stage, stage_cap, nz, &sweeps);
  else fast_argmin_z<RL, CS>(m, rank_one, sm, 1.0, true, false, nz, &sweeps);
  __syncthreads();
  if constexpr (TC && CS > 1) cl_sync<CS>();   // peers have read this CTA's Gram partial: the overlay is free again
  if (prm.need_dual)     // AtY = A' Y  (:289)
    product_ah(sm.Y, [&](int k, int c, cd v) { AtY[k + (size_t)FN * (c0 + c)] = cscale(v, cs); });

  double mu = prm.mu0, opt_obj = INFINITY, last_res = INFINITY, res_comb = 0.0;
  int iters = 0, opt_iter = -1, opt_col = -1, bumps = 0, converged = 0, have_opt = 0;
  const int rout = tk.sbr ? r : 1;
  if (tid == 0) { sm.sc[20] = 0.0; sm.sc[21] = 0.0; sm.sc[22] = 0.0; sm.sc[23] = 0.0; sm.sc[24] = 0.0; }
  // sm.ifl[2] counts eigen-decompositions since the stage began (set to 0 before the :288 call)
  const long long tl0 = clock64();

  for (int it = 1; it <= prm.maxiter; ++it) {
    const double imu = 1.0 / mu;
    const long long tx0 = clock64();
    double* xsc = sm.xsc + (it & 1) * XS_SCAL;
    double* xcol = sm.xcol + (it & 1) * SMALL_DMAX;
    // ---- X update (:304, :380-388) in its two-product Woodbury form.  With Q = Z - N/mu, T = Y - M/mu and
    // S = I + A A':   inv(A'A + I)(A'T + Q) = Q + A' W,   A X = T - W,   W = S^-1 (T - A Q)
    for (int idx = tid; idx < m * RL; idx += NT) {       // T -> AX (becomes A X below)
      const cd y = sm.Y[idx], mm = sm.M[idx];
      sm.AX[idx] = cmk(fma(-mm.x, imu, y.x), fma(-mm.y, imu, y.y));
    }
    for (int idx = tid; idx < FN * RL; idx += NT) {      // Q -> X
      const cd z = sm.Z[idx], nn = sm.N[idx];
      sm.X[idx] = cmk(fma(-nn.x, imu, z.x), fma(-nn.y, imu, z.y));
    }
    __syncthreads();
    product_a(sm.X, [&](int i, int c, cd v) {                      // R = T - A Q -> WT
      const cd t = sm.AX[i + m * c];
      sm.WT[i + m * c] = cmk(fma(-v.x, cs, t.x), fma(-v.y, cs, t.y));
    });
    if constexpr (TC) tc_prefetch<true>(tc, tg, m);                    // the A' W blocks arrive during the S^-1 product
    prod_sinv<RL>(Sinv, m, sm.WT, sm.AX);                           // W -> WT, A X = T - W -> AX
    product_ah(sm.WT, [&](int k, int c, cd v) {                    // X = Q + A' W
      const int p = k + FN * c;
      const cd x = sm.X[p];
      sm.X[p] = cmk(fma(v.x, cs, x.x), fma(v.y, cs, x.y));
    });
    __syncthreads();
    const long long tx1 = clock64();
    if (tid == 0) sm.sc[21] += (double)(tx1 - tx0);
    // ---- Y update (:308), M update (:315-316), objective (:323-340); Y0 - Y kept in WT for A'(Y-Y0)
    double pYd2 = 0.0, pJM2 = 0.0, pY2 = 0.0, pAX2 = 0.0, obj2 = 0.0, nAX2 = 0.0;
    const double i1mu = 1.0 / (1.0 + mu);
    if (tk.sbr) {
      // thread (row i, lane-group member s): columns c = s, s + ks, ...; row sums by shuffle
      const int ks = ksplit_for(m, 8);
      const int i = tid / ks, s = tid - i * ks;
      const bool act = i < m;
      double d2 = 0.0, a2 = 0.0;
      if (act) {
        for (int c = s; c < RL; c += ks) {
          const cd ax = sm.AX[i + m * c], mm = sm.M[i + m * c];
          d2 += cabs2(cmk(fma(mm.x, imu, ax.x), fma(mm.y, imu, ax.y)));
          a2 += cabs2(ax);
        }
      }
      for (int o = 1; o < ks; o <<= 1) {
        d2 += __shfl_xor_sync(0xffffffffu, d2, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
      }
      if (act && s == 0) { sm.xrow[i] = d2; sm.xrow[fd.maxm + i] = a2; }
      cl_sync<CS>();
      if (act) {
        d2 = 0.0; a2 = 0.0;
#pragma unroll
        for (int rk = 0; rk < CS; ++rk) {
          const double* pr = peer_ptr<double, CS>(sm.xrow, rk);
          d2 += pr[i];
          a2 += pr[fd.maxm + i];
        }
        double D = sqrt(d2);
        const bool z = (D == 0.0);
        if (z) D = 1.0;
        const double f = (sm.Bs[i] / D + mu) * i1mu;
        const double isr = 1.0 / sqrt((double)r);
        for (int c = s; c < RL; c += ks) {
          const int p = i + m * c;
          const cd ax = sm.AX[p], mm = sm.M[p], yo = sm.Y[p];
          const cd cc = z ? cmk(isr, 0.0) : cmk(fma(mm.x, imu, ax.x), fma(mm.y, imu, ax.y));
          const cd yn = cscale(cc, f);
          const cd jm = cmk(ax.x - yn.x, ax.y - yn.y);
          const cd dy = cmk(yn.x - yo.x, yn.y - yo.y);
          sm.Y[p] = yn;
          sm.M[p] = cmk(fma(mu, jm.x, mm.x), fma(mu, jm.y, mm.y));
          sm.WT[p] = dy;
          pYd2 += cabs2(dy);
          pJM2 += cabs2(jm);
          pY2 += cabs2(yn);
        }
        if (s == 0) {
          nAX2 += a2;                    // cluster totals: identical in every CTA
          const double dd = sqrt(a2) - sm.Bs[i];
          obj2 += dd * dd;
        }
      }
    } else {
      for (int c = warp; c < RL; c += NW) {
        double oc = 0.0;
        for (int i = lane; i < m; i += 32) {
          const int p = i + m * c;
          const cd ax = sm.AX[p], mm = sm.M[p], yo = sm.Y[p];
          cd cc = cmk(fma(mm.x, imu, ax.x), fma(mm.y, imu, ax.y));
          double D = sqrt(cabs2(cc));
          if (D == 0.0) { cc = cmk(1.0, 0.0); D = 1.0; }
          const double f = (sm.Bs[i] / D + mu) * i1mu;
          const cd yn = cscale(cc, f);
          const cd jm = cmk(ax.x - yn.x, ax.y - yn.y);
          const cd dy = cmk(yn.x - yo.x, yn.y - yo.y);
          sm.Y[p] = yn;
          sm.M[p] = cmk(fma(mu, jm.x, mm.x), fma(mu, jm.y, mm.y));
          sm.WT[p] = dy;
          pYd2 += cabs2(dy);
          pJM2 += cabs2(jm);
          pY2 += cabs2(yn);
          const double a2 = cabs2(ax);
          pAX2 += a2;
          const double dd = sqrt(a2) - sm.Bs[i];
          oc += dd * dd;
        }
        oc = warp_sum(oc);
        if (lane == 0) xcol[c0 + c] = sqrt(oc);
      }
    }
    {
      double v[6] = {pYd2, pJM2, pY2, pAX2, obj2, nAX2};
      block_sum<6>(v, sm.red);
      pYd2 = v[0]; pJM2 = v[1]; pY2 = v[2]; pAX2 = v[3]; obj2 = v[4]; nAX2 = v[5];
    }
    // ---- A'(Y - Y0) (:309): advances AtY in global memory; only feeds res_dual
    double pAtYd2 = 0.0, pAtY2 = 0.0;
    if (prm.need_dual) {
      double v[2] = {0.0, 0.0};
      product_ah(sm.WT, [&](int k, int c, cd acc) {
        const size_t p = k + (size_t)FN * (c0 + c);
        const cd d = cscale(acc, cs);
        cd a = AtY[p];
        a.x += d.x;
        a.y += d.y;
        AtY[p] = a;
        v[0] += cabs2(d);
        v[1] += cabs2(a);
      });
      block_sum<2>(v, sm.red);
      pAtYd2 = v[0]; pAtY2 = v[1];
    }
    const long long tx2 = clock64();
    if (tid == 0) sm.sc[22] += (double)(tx2 - tx1);
    // ---- Z, N update (:312, :319-320)
    // warm-started from the previous eigenvectors; a cold start every 64 iterations re-orthonormalises U
    if (fd.nuclear) {
      // A'(Y-Y0) has consumed WT by now; the screen scalar uses the last slot of this iteration's xsc
      fast_argmin_z_nuclear<RL, CS>(fd, sm, rank, mu, false, xsc + XS_SCAL - 1, stage, stage_cap, nz, &sweeps);
    } else {
      fast_argmin_z<RL, CS>(m, rank_one, sm, mu, false, (sm.ifl[2] & 63) != 0, nz, &sweeps);
    }
    const long long tx3 = clock64();
    if (tid == 0) sm.sc[23] += (double)(tx3 - tx2);
    // ---- cluster-wide scalars
    if (tid == 0) {
      xsc[0] = pYd2; xsc[1] = pJM2; xsc[2] = pY2; xsc[3] = pAX2; xsc[4] = pAtYd2; xsc[5] = pAtY2;
      xsc[6] = nz[0]; xsc[7] = nz[1]; xsc[8] = nz[2]; xsc[9] = nz[3];
    }
    cl_sync<CS>();
    double tot[10];
#pragma unroll
    for (int q = 0; q < 10; ++q) tot[q] = 0.0;
#pragma unroll
    for (int rk = 0; rk < CS; ++rk) {
      const double* pr = peer_ptr<double, CS>(xsc, rk);
#pragma unroll
      for (int q = 0; q < 10; ++q) tot[q] += pr[q];
    }
    const double nYd2 = tot[0], nJM2 = tot[1], nY2 = tot[2];
    if (!tk.sbr) nAX2 = tot[3];
    const double nAtYd2 = tot[4], nAtY2 = tot[5], nJN2 = tot[6], nZd2 = tot[7], nX2 = tot[8], nZ2 = tot[9];

    // ---- best solution so far (:323-340); NaN objectives never win (MATLAB min skips NaN)
    double obj; int jbest = -1;
    if (tk.sbr) {
      obj = sqrt(obj2);
    } else {
      obj = NAN;
      for (int c = 0; c < r; ++c) {
        const double oc = peer_ptr<double, CS>(xcol, c / RL)[c];
        if (oc == oc && (jbest < 0 || oc < obj)) { obj = oc; jbest = c; }
      }
    }
    if (obj < opt_obj) {
      opt_obj = obj; opt_iter = it; opt_col = jbest; have_opt = 1;
      if (tk.sbr) {
        if (tk.Xout) for (int idx = tid; idx < FN * RL; idx += NT) tk.Xout[(size_t)FN * c0 + idx] = sm.X[idx];
        if (tk.Yout) for (int idx = tid; idx < m * RL; idx += NT) tk.Yout[(size_t)m * c0 + idx] = sm.Y[idx];
      } else if (jbest / RL == rank) {
        const int cl = jbest - c0;
        if (tk.Xout) for (int k = tid; k < FN; k += NT) tk.Xout[k] = sm.X[k + FN * cl];
        if (tk.Yout) for (int i = tid; i < m; i += NT) tk.Yout[i] = sm.Y[i + m * cl];
      }
    }
    // ---- residuals and stopping rule (:343-354)
    const double res_prim = sqrt(nJM2 + nJN2);
    const double res_dual = mu * sqrt(nAtYd2 + nZd2);
    res_comb = sqrt(nJM2 + nJN2 + nYd2 + nZd2);
    iters = it;
    if (tk.trace != nullptr && tid == 0 && rank == 0) tk.trace[it - 1] = res_comb;
    if (prm.need_dual) {
      const double mx1 = fmax(sqrt(nAX2), sqrt(nY2)), mx2 = fmax(sqrt(nX2), sqrt(nZ2));
      const double th_prim = prm.tol_abs * sqrt((double)(m + FN) * r) + prm.tol_rel * sqrt(mx1 * mx1 + mx2 * mx2);
      const double th_dual = prm.tol_abs * sqrt((double)FN * r * 2.0) + prm.tol_rel * sqrt(nAtY2 + nZ2);
      const double th_comb = prm.tol_abs * sqrt((double)(m + FN) * r * 2.0) +
                             prm.tol_rel * sqrt(mx1 * mx1 + mx2 * mx2 + nY2 + nZ2);
      if ((res_prim < th_prim && res_dual < th_dual) || (res_comb < th_comb)) { converged = 1; break; }
    }
    if (res_comb > last_res * 0.9) { mu *= prm.rho; ++bumps; }   // :358-361
    last_res = res_comb;
    __syncthreads();
    if (tid == 0) sm.sc[24] += (double)(clock64() - tx3);
  }
  __syncthreads();
  // ---- outputs: the best iterate is already in place; all-NaN objectives give NaN (H4)
  if (!have_opt) {
    if (tk.sbr) {
      if (tk.Xout) for (int idx = tid; idx < FN * RL; idx += NT) tk.Xout[(size_t)FN * c0 + idx] = cmk(NAN, NAN);
      if (tk.Yout) for (int idx = tid; idx < m * RL; idx += NT) tk.Yout[(size_t)m * c0 + idx] = cmk(NAN, NAN);
    } else if (rank == 0) {
      if (tk.Xout) for (int k = tid; k < FN; k += NT) tk.Xout[k] = cmk(NAN, NAN);
      if (tk.Yout) for (int i = tid; i < m; i += NT) tk.Yout[i] = cmk(NAN, NAN);
    }
  }
  (void)rout;
  if (tk.state) {   // [X Z N (n x r) | Y M (m x r)], own columns
    cd* st = tk.state;
    const size_t nr = (size_t)FN * r, mr = (size_t)m * r;
    for (int idx = tid; idx < FN * RL; idx += NT) {
      const size_t g = (size_t)FN * c0 + idx;
      st[g] = sm.X[idx]; st[nr + g] = sm.Z[idx]; st[2 * nr + g] = sm.N[idx];
    }
    for (int idx = tid; idx < m * RL; idx += NT) {
      const size_t g = (size_t)m * c0 + idx;
      st[3 * nr + g] = sm.Y[idx]; st[3 * nr + mr + g] = sm.M[idx];
    }
  }
  if (tk.scal && tid == 0 && rank == 0) {
    tk.scal[SC_MU] = mu; tk.scal[SC_OPT_OBJ] = opt_obj; tk.scal[SC_ITERS] = iters;
    tk.scal[SC_OPT_ITER] = opt_iter; tk.scal[SC_OPT_COL] = opt_col; tk.scal[SC_BUMPS] = bumps;
    tk.scal[SC_CONVERGED] = converged; tk.scal[SC_RES_COMB] = res_comb; tk.scal[SC_SWEEPS] = sweeps;
    tk.scal[9] = sm.sc[20]; tk.scal[10] = sm.sc[21]; tk.scal[11] = (double)(clock64() - tl0);
    tk.scal[12] = sm.sc[22]; tk.scal[13] = sm.sc[23]; tk.scal[14] = sm.sc[24];
    tk.scal[15] = (double)(tl0 - ttask0);      // set-up: codes, S and its inverse, operand blocks, first ArgMinZ
  }
  cl_sync<CS>();   // no CTA leaves (or reuses its exchange buffers) while a peer may still read them
}

template <int RL, int CS, bool TC>
__global__ void __launch_bounds__(NT, (RL == 1) ? 2 : 1)
fast_stage_kernel(const StageTask* __restrict__ tasks, int ntasks, DevParams prm, FastDims fd, cd* wsbase, int* counter) {
  extern __shared__ __align__(1024) unsigned char fast_smem_raw[];
  __shared__ int s_next_task;
  const FastSmem<RL> sm = fast_carve<RL>(fast_smem_raw, fd);
  int rank = 0;
  if constexpr (CS > 1) rank = (int)cg::this_cluster().block_rank();
  const int cid = blockIdx.x / CS, ncl = gridDim.x / CS;
  cd* wsg = wsbase + (size_t)cid * fd.ws_stride;
  TcCtx tc = {};
  if constexpr (TC) {
    // one CTA per SM (the host pads the shared-memory request): it owns all 512 tensor-memory columns
    if (threadIdx.x == 0) {
      for (int s = 0; s < 2 * TC_MAXSLOT + 1; ++s) umma::mbar_init(sm.tc_bars + s, 1);
      umma::mbar_fence_init();
    }
    if (threadIdx.x < 32) umma::tmem_alloc512(sm.tc_tslot);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    tc.Bs = sm.tc_bs; tc.ov = sm.tc_ov; tc.ex = sm.tc_ex; tc.bars = sm.tc_bars; tc.tmem = *sm.tc_tslot;
    tc.nslot_launch = fd.tc.nslot; tc.nslot = fd.tc.nslot; tc.n1 = fd.tc.n1; tc.slot_bytes = fd.tc.slot_bytes;
    tc.wt = (unsigned char*)sm.WT; tc.wt_slot = -1;
  }
  // Tasks are handed out through a device counter (longest first): a cluster that starts late -- the launch may share
  // the GPU with the other kernel groups of its stage -- simply takes fewer of them.  counter == nullptr: round robin.
  for (int t = cid;; t += ncl) {
    if (counter != nullptr) t = next_task<CS>(counter, &s_next_task, rank);
    if (t >= ntasks) break;
    const StageTask tk = tasks[t];
    if (tk.active != nullptr && *tk.active != tk.active_expect) continue;   // cluster-uniform
    run_fast<RL, CS, TC>(tk, prm, fd, sm, wsg, rank, tc);
  }
  if constexpr (TC) {
    umma::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) umma::tmem_free512(tc.tmem);
  }
}

}  // namespace twoace
