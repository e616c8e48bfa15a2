// InferADMM on the cluster kernel for 256 < m <= 1024 rows (16 x 16 antennas, quantised sensing matrix, r = 20):
// the upper half of the reference's measurement sweep M = [4 36 121 225 361 529 784 1024]
// (main/channel_recovery_ADMM_v2_simulation_A2only.m:106-118), where the solver actually recovers the channel.
//
// Same algorithm as fast_stage.cuh (inferLowRankV4.m:260-365), columns of the iterate split over a cluster of 4 CTAs.
// What changes with m > n = 256:
//   * ArgMinX uses the n x n form of :383 directly, X = U (A'T + Q) with U = (A'A + I)^-1 (256 x 256, in L2) and
//     A X as a separate product -- the m x m Woodbury core of the small-m kernel would be the larger matrix here;
//   * the m-sized arrays Y and M (and Y - Y0 when a tolerance is set) do not fit in shared memory: they live in
//     global memory (L2-resident: 3 x m x 20 x 16 B per cluster) and every pass over them runs in chunks of 256 rows
//     through the shared-memory buffers of the small-m kernel;
//   * both sensing-matrix products are tensor-core products (tc_prod.cuh) over row chunks: A'T accumulates all chunks
//     in the same tensor-memory accumulators (the operand digits are rewritten chunk by chunk, with column scales
//     from a pre-pass), A X slices X once and reads the accumulators back chunk by chunk, so that the Y / M update
//     of a chunk (which needs the row norms over all 20 columns: one cluster exchange per chunk) follows its product;
//   * A'A itself is a tensor-core product (the columns of u are exact one-digit operands), built once per trial.
#pragma once
#include "fast_stage.cuh"

namespace twoace {

constexpr int BIG_CH = 256;        // rows per chunk (two 128-row operand tiles)
constexpr int BIG_RL = TC_NC;      // columns per CTA
constexpr int BIG_CS = 4;          // CTAs per cluster
constexpr int BIG_R = BIG_RL * BIG_CS;
constexpr int BIG_MMAX = 1024;

// global workspace of one cluster, in cd units:  U | AtY | operand blocks | Y | M | Y - Y0
__host__ __device__ inline size_t big_ws_elems(int mfull) {
  const size_t mt = (size_t)(mfull + 127) / 128;
  return (size_t)FN * FN + (size_t)FN * BIG_R + mt * 4096 + 3 * (size_t)mfull * BIG_R;
}

// Operand blocks of the whole instance -> global memory (tc.aop), the work split over the CTAs of the cluster.
// Block (p, it, ks) at ((p * mt + it) * 2 + ks) * 16 KB, byte (i, k) at (k/16 % 8) * 2048 + (i % 128) * 16 + k % 16.
__device__ inline void big_build_blocks(const TcCtx& tc, const uint32_t* __restrict__ codes, const int* rows_s, int m, int mt,
                                        int part, int nparts) {
  const int items = mt * 128 * 16;
  for (int idx = threadIdx.x + NT * part; idx < items; idx += NT * nparts) {
    const int il = idx & 127, t2 = idx >> 7, it = t2 % mt, w = t2 / mt;
    const int i = 128 * it + il;
    uint32_t re[4] = {0u, 0u, 0u, 0u}, im[4] = {0u, 0u, 0u, 0u};
    if (i < m) {
      const uint32_t word = codes[(size_t)rows_s[i] * 16 + w];
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const uint32_t c = (word >> (2 * q)) & 3u;
        re[q >> 2] |= ((0x00FF0001u >> (8 * c)) & 0xFFu) << (8 * (q & 3));   // 1, 0, -1, 0
        im[q >> 2] |= ((0xFF000100u >> (8 * c)) & 0xFFu) << (8 * (q & 3));   // 0, 1, 0, -1
      }
    }
    const int ks = w >> 3;
    const size_t inblk = (size_t)(w & 7) * 2048 + (size_t)il * 16;
    unsigned char* d0 = tc.aop + (size_t)tc_block_id(0, it, ks, mt) * 16384 + inblk;
    unsigned char* d1 = tc.aop + (size_t)tc_block_id(1, it, ks, mt) * 16384 + inblk;
    *reinterpret_cast<uint4*>(d0) = make_uint4(re[0], re[1], re[2], re[3]);
    *reinterpret_cast<uint4*>(d1) = make_uint4(im[0], im[1], im[2], im[3]);
  }
  asm volatile("fence.proxy.async;" ::: "memory");
  __threadfence();
}

// out(k, c, sum_i conj(u(i,k)) in(i, c)), k < 256, c < TC_NC, over all m rows in chunks of 256.
// in(i, c) is evaluated twice per row (scale pre-pass and slicing).  Every thread of the CTA calls it.
template <class InF, class OutF>
__device__ __forceinline__ void big_product_ah(TcCtx& tc, int m, int mt, InF in, OutF out, uint32_t* red) {
  const int tid = threadIdx.x;
  uint32_t h[TC_NC];
#pragma unroll
  for (int c = 0; c < TC_NC; ++c) h[c] = 0u;
  for (int i = tid; i < m; i += NT) {
#pragma unroll
    for (int c = 0; c < TC_NC; ++c) h[c] = max(h[c], tc_hi(in(i, c)));
  }
  TcScale S;
  tc_scales(h, red, S);
  const int nch = (mt + 1) >> 1;
  for (int ch = 0; ch < nch; ++ch) {
    TcPass ps;
    ps.R = 128; ps.nt = min(2, mt - 2 * ch); ps.t0 = 2 * ch; ps.mt = mt; ps.rows = m - BIG_CH * ch;
    ps.resident = false; ps.s0 = 0; ps.acc_first = ch > 0;
    tc_first_round<true>(tc, ps, false);       // the ring is free: the previous pass has completed
    const int i = BIG_CH * ch + tid;
    if (i < m) {
      cd x[TC_NC];
#pragma unroll
      for (int c = 0; c < TC_NC; ++c) x[c] = in(i, c);
      tc_slice_row(tc.Bs, tid, x, S);
    }
    umma::fence_async_smem();
    __syncthreads();
    tc_mma_pass<true>(tc, ps);
    __syncthreads();                           // the B operand may be rewritten
  }
  tc_epilogue<true>(tc, S, FN, out);
  __syncthreads();
}

// per chunk ch: out(il, c, sum_k u(256 ch + il, k) in(k, c)) for il < rows of the chunk, then after_chunk(ch, rows).
template <class InF, class OutF, class ChunkF>
__device__ __forceinline__ void big_product_a(TcCtx& tc, int m, int mt, InF in, OutF out, ChunkF after_chunk, uint32_t* red) {
  const int tid = threadIdx.x;
  cd x[TC_NC];
  uint32_t h[TC_NC];
#pragma unroll
  for (int c = 0; c < TC_NC; ++c) { x[c] = in(tid, c); h[c] = tc_hi(x[c]); }
  TcScale S;
  tc_scales(h, red, S);
  tc_slice_row(tc.Bs, tid, x, S);
  umma::fence_async_smem();
  __syncthreads();
  const int nch = (mt + 1) >> 1;
  for (int ch = 0; ch < nch; ++ch) {
    TcPass ps;
    ps.R = 128; ps.nt = min(2, mt - 2 * ch); ps.t0 = 2 * ch; ps.mt = mt; ps.rows = 0;
    ps.resident = false; ps.s0 = 0; ps.acc_first = false;
    tc_first_round<false>(tc, ps, false);
    tc_mma_pass<false>(tc, ps);
    __syncthreads();
    const int rows = min(BIG_CH, m - BIG_CH * ch);
    tc_epilogue<false>(tc, S, rows, out);
    __syncthreads();
    after_chunk(ch, rows);
  }
}

// V <- U V in place (U: d x d in global memory / L2, V: d x RL in shared memory, d = 256): the n x n solve of :383.
template <int RL>
__device__ __forceinline__ void prod_sq_inplace(const cd* __restrict__ U, cd* V) {
  constexpr int d = FN, NR = 4;
  const int tid = threadIdx.x;
  constexpr int mq = d / NR;                // 64 row quads
  constexpr int ks = NT / mq;               // 4 lanes per quad
  const int iq = tid / ks, s = tid - iq * ks;
  cd acc[NR][RL];
#pragma unroll
  for (int u = 0; u < NR; ++u)
#pragma unroll
    for (int c = 0; c < RL; ++c) acc[u][c] = cmk(0.0, 0.0);
  constexpr int nj = d / ks;
  cd cur[2][NR], nxt[2][NR];
  auto fetch = [&](cd (&dst)[2][NR], int q) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = s + ks * min(q + h, nj - 1);
#pragma unroll
      for (int u = 0; u < NR; ++u) dst[h][u] = __ldg(U + NR * iq + u + (size_t)d * j);
    }
  };
  fetch(cur, 0);
  for (int q = 0; q < nj; q += 2) {
    if (q + 2 < nj) fetch(nxt, q + 2);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = s + ks * (q + h);
#pragma unroll
      for (int c = 0; c < RL; ++c) {
        const cd r = V[j + d * c];
#pragma unroll
        for (int u = 0; u < NR; ++u) cfma(acc[u][c], cur[h][u], r);
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int u = 0; u < NR; ++u) cur[h][u] = nxt[h][u];
  }
#pragma unroll
  for (int u = 0; u < NR; ++u) group_reduce<RL>(acc[u], ks);
  __syncthreads();            // every thread has finished reading V
  if (s == 0) {
#pragma unroll
    for (int u = 0; u < NR; ++u)
#pragma unroll
      for (int c = 0; c < RL; ++c) V[NR * iq + u + d * c] = acc[u][c];
  }
  __syncthreads();
}

__device__ inline void run_big(const StageTask& tk, const DevParams& prm, const FastDims& fd, const FastSmem<BIG_RL>& sm,
                               cd* wsg, int rank, TcCtx& tc) {
  constexpr int RL = BIG_RL, CS = BIG_CS, r = BIG_R;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m = tk.m, mt = (m + 127) >> 7;
  const int c0 = rank * RL;
  const double cs = *tk.cscale, bsc = *tk.bscale;
  const int rank_one = tk.rank_one_ptr ? *tk.rank_one_ptr : tk.rank_one;
  const long long ttask0 = clock64();
  // ---- cluster workspace
  const size_t mtl = (size_t)(fd.mfull + 127) / 128;
  cd* Uws = wsg;
  cd* AtY = wsg + (size_t)FN * FN;
  tc.aop = (unsigned char*)(AtY + (size_t)FN * r);
  cd* Yg = (cd*)(tc.aop + mtl * 65536);
  cd* Mg = Yg + (size_t)fd.mfull * r;
  cd* dYg = Mg + (size_t)fd.mfull * r;
  cd* U = tk.sinv != nullptr ? tk.sinv : Uws;               // (A'A + I)^-1, 256 x 256
  const bool u_reuse = tk.sinv != nullptr && tk.sinv_state == 1;
  tc.nslot = min(tc.nslot_launch, NW - 1);
  tc.wt_slot = -1;
  tc.premask = 0;
  uint32_t* red = (uint32_t*)sm.red;

  // ---- stage-local copies
  for (int i = tid; i < m; i += NT) {
    sm.rows_s[i] = tk.A.rows ? tk.A.rows[i] : i;
    sm.Bs[i] = bsc * tk.B[tk.brows ? tk.brows[i] : i];
  }
  jacobi_tables<FTX>(sm.pairs);
  __syncthreads();
  double nb2;
  {
    double v[1] = {0.0};
    for (int i = tid; i < m; i += NT) v[0] += sm.Bs[i] * sm.Bs[i];
    block_sum<1>(v, sm.red);
    nb2 = v[0];
  }
  const double normB = sqrt(nb2);
  big_build_blocks(tc, tk.codes, sm.rows_s, m, mt, rank, CS);
  cl_sync<CS>();
  // u(i, k) as a complex number, from the 2-bit code
  auto ucode = [&](int i, int k) {
    const uint32_t c = (tk.codes[(size_t)sm.rows_s[i] * 16 + (k >> 4)] >> (2 * (k & 15))) & 3u;
    return cmk((c == 0u) - (c == 2u), (c == 1u) - (c == 3u));
  };
  // ---- U = (A'A + I)^-1: columns [64 rank, 64 rank + 64) of A'A by this CTA (exact: one-digit operands), inverse by rank 0
  if (!u_reuse) {
    const double cs2 = cs * cs;
    for (int l0 = 64 * rank; l0 < 64 * rank + 64; l0 += TC_NC) {
      big_product_ah(tc, m, mt, [&](int i, int c) { return ucode(i, min(l0 + c, FN - 1)); },
                     [&](int k, int c, cd v) {
                       const int l = l0 + c;
                       if (l < 64 * rank + 64) U[k + (size_t)FN * l] = cmk(fma(cs2, v.x, k == l ? 1.0 : 0.0), cs2 * v.y);
                     }, red);
    }
    __threadfence();
    cl_sync<CS>();
    // Gauss-Jordan with the columns split over the cluster: one cluster barrier per pivot step
    spd_inverse_part(U, FN, sm.WT, sm.WT + FN, rank, CS, [] { __threadfence(); cl_sync<CS>(); });
  }

  // ---- X = X0 (own columns), M = N = 0
  for (int idx = tid; idx < FN * RL; idx += NT) {
    sm.X[idx] = tk.X0[(size_t)FN * c0 + idx];
    sm.N[idx] = cmk(0.0, 0.0);
  }
  for (int idx = tid; idx < m * RL; idx += NT) Mg[(size_t)m * c0 + idx] = cmk(0.0, 0.0);
  __syncthreads();
  auto x_in = [&](int k, int c) { return sm.X[k + FN * c]; };
  auto ax_out = [&](int il, int c, cd v) { sm.AX[il + BIG_CH * c] = cscale(v, cs); };
  // rescale so |A X| matches |B|  (:278-286): first pass for the norms only
  {
    double acc[RL];
#pragma unroll
    for (int c = 0; c < RL; ++c) acc[c] = 0.0;
    big_product_a(tc, m, mt, x_in, ax_out, [&](int ch, int rows) {
      (void)ch;
      if (tid < rows) {
#pragma unroll
        for (int c = 0; c < RL; ++c) acc[c] += cabs2(sm.AX[tid + BIG_CH * c]);
      }
      __syncthreads();
    }, red);
    block_sum<RL>(acc, sm.red);
    if (tk.sbr) {
      double t = 0.0;
#pragma unroll
      for (int c = 0; c < RL; ++c) t += acc[c];
      if (tid == 0) sm.xsc[0] = t;
      cl_sync<CS>();
      double tot = 0.0;
#pragma unroll
      for (int rk = 0; rk < CS; ++rk) tot += peer_ptr<double, CS>(sm.xsc, rk)[0];
      const double s = normB / sqrt(tot);
      for (int idx = tid; idx < FN * RL; idx += NT) sm.X[idx] = cscale(sm.X[idx], s);
      cl_sync<CS>();   // peers are done with xsc[0] before it is reused
    } else {
      for (int idx = tid; idx < FN * RL; idx += NT) sm.X[idx] = cscale(sm.X[idx], normB / sqrt(acc[idx / FN]));
    }
    __syncthreads();
  }
  // Y = normalize_rows(A X, B)  (:287, :517-538), chunk by chunk
  big_product_a(tc, m, mt, x_in, ax_out, [&](int ch, int rows) {
    const int i = BIG_CH * ch + tid;
    double* xr = (ch & 1) ? sm.rowtot : sm.xrow;
    if (tk.sbr) {
      if (tid < rows) {
        double d2 = 0.0;
#pragma unroll
        for (int c = 0; c < RL; ++c) d2 += cabs2(sm.AX[tid + BIG_CH * c]);
        xr[tid] = d2;
      }
      cl_sync<CS>();
      if (tid < rows) {
        double d2 = 0.0;
#pragma unroll
        for (int rk = 0; rk < CS; ++rk) d2 += peer_ptr<double, CS>(xr, rk)[tid];
        double D = sqrt(d2);
        const bool z = (D == 0.0);
        if (z) D = 1.0;
        const double f = sm.Bs[i] / D, isr = 1.0 / sqrt((double)r);
#pragma unroll
        for (int c = 0; c < RL; ++c) {
          const cd v = z ? cmk(isr, 0.0) : sm.AX[tid + BIG_CH * c];
          Yg[i + (size_t)m * (c0 + c)] = cscale(v, f);
        }
      }
    } else if (tid < rows) {
#pragma unroll
      for (int c = 0; c < RL; ++c) {
        cd v = sm.AX[tid + BIG_CH * c];
        double D = sqrt(cabs2(v));
        if (D == 0.0) { v = cmk(1.0, 0.0); D = 1.0; }
        Yg[i + (size_t)m * (c0 + c)] = cscale(v, sm.Bs[i] / D);
      }
    }
    __syncthreads();
  }, red);
  int sweeps = 0;
  double nz[4];
  if (tid == 0) sm.ifl[2] = 0;
  __syncthreads();
  fast_argmin_z<RL, CS>(m, rank_one, sm, 1.0, true, false, nz, &sweeps);     // Z = ArgMinZ(X, 0, 1)  (:288)
  __syncthreads();
  cl_sync<CS>();   // peers have read this CTA's Gram partial: the overlay is free again
  if (prm.need_dual)     // AtY = A' Y  (:289)
    big_product_ah(tc, m, mt, [&](int i, int c) { return Yg[i + (size_t)m * (c0 + c)]; },
                   [&](int k, int c, cd v) { AtY[k + (size_t)FN * (c0 + c)] = cscale(v, cs); }, red);

  double mu = prm.mu0, opt_obj = INFINITY, last_res = INFINITY, res_comb = 0.0;
  int iters = 0, opt_iter = -1, opt_col = -1, bumps = 0, converged = 0, have_opt = 0;
  if (tid == 0) { sm.sc[20] = 0.0; sm.sc[21] = 0.0; sm.sc[22] = 0.0; sm.sc[23] = 0.0; sm.sc[24] = 0.0; }
  const long long tl0 = clock64();

  for (int it = 1; it <= prm.maxiter; ++it) {
    const double imu = 1.0 / mu, i1mu = 1.0 / (1.0 + mu);
    const long long tx0 = clock64();
    double* xsc = sm.xsc + (it & 1) * XS_SCAL;
    double* xcol = sm.xcol + (it & 1) * SMALL_DMAX;
    // ---- X update (:304, :380-388): X = U (A'(Y - M/mu) + Z - N/mu)
    big_product_ah(tc, m, mt,
                   [&](int i, int c) {
                     const size_t p = i + (size_t)m * (c0 + c);
                     const cd y = Yg[p], mm = Mg[p];
                     return cmk(fma(-mm.x, imu, y.x), fma(-mm.y, imu, y.y));
                   },
                   [&](int k, int c, cd v) {
                     const int p = k + FN * c;
                     const cd z = sm.Z[p], nn = sm.N[p];
                     sm.X[p] = cmk(fma(v.x, cs, fma(-nn.x, imu, z.x)), fma(v.y, cs, fma(-nn.y, imu, z.y)));
                   }, red);
    prod_sq_inplace<RL>(U, sm.X);
    const long long tx1 = clock64();
    if (tid == 0) sm.sc[21] += (double)(tx1 - tx0);
    // ---- A X (:305) chunk by chunk, each followed by its Y update (:308), M update (:315-316), objective (:323-340)
    double pYd2 = 0.0, pJM2 = 0.0, pY2 = 0.0, pAX2 = 0.0, obj2 = 0.0, nAX2 = 0.0;
    double ocol[RL];
#pragma unroll
    for (int c = 0; c < RL; ++c) ocol[c] = 0.0;
    big_product_a(tc, m, mt, x_in, ax_out, [&](int ch, int rows) {
      const int i = BIG_CH * ch + tid;
      double* xr = (ch & 1) ? sm.rowtot : sm.xrow;
      if (tk.sbr) {
        if (tid < rows) {
          double d2 = 0.0, a2 = 0.0;
#pragma unroll
          for (int c = 0; c < RL; ++c) {
            const cd ax = sm.AX[tid + BIG_CH * c], mm = Mg[i + (size_t)m * (c0 + c)];
            d2 += cabs2(cmk(fma(mm.x, imu, ax.x), fma(mm.y, imu, ax.y)));
            a2 += cabs2(ax);
          }
          xr[tid] = d2;
          xr[BIG_CH + tid] = a2;
        }
        cl_sync<CS>();
        if (tid < rows) {
          double d2 = 0.0, a2 = 0.0;
#pragma unroll
          for (int rk = 0; rk < CS; ++rk) {
            const double* pr = peer_ptr<double, CS>(xr, rk);
            d2 += pr[tid];
            a2 += pr[BIG_CH + tid];
          }
          double D = sqrt(d2);
          const bool z = (D == 0.0);
          if (z) D = 1.0;
          const double f = (sm.Bs[i] / D + mu) * i1mu, isr = 1.0 / sqrt((double)r);
#pragma unroll
          for (int c = 0; c < RL; ++c) {
            const size_t p = i + (size_t)m * (c0 + c);
            const cd ax = sm.AX[tid + BIG_CH * c], mm = Mg[p], yo = Yg[p];
            const cd cc = z ? cmk(isr, 0.0) : cmk(fma(mm.x, imu, ax.x), fma(mm.y, imu, ax.y));
            const cd yn = cscale(cc, f);
            const cd jm = cmk(ax.x - yn.x, ax.y - yn.y), dy = cmk(yn.x - yo.x, yn.y - yo.y);
            Yg[p] = yn;
            Mg[p] = cmk(fma(mu, jm.x, mm.x), fma(mu, jm.y, mm.y));
            if (prm.need_dual) dYg[p] = dy;
            pYd2 += cabs2(dy); pJM2 += cabs2(jm); pY2 += cabs2(yn);
          }
          nAX2 += a2;                    // cluster totals: identical in every CTA
          const double dd = sqrt(a2) - sm.Bs[i];
          obj2 += dd * dd;
        }
      } else if (tid < rows) {
#pragma unroll
        for (int c = 0; c < RL; ++c) {
          const size_t p = i + (size_t)m * (c0 + c);
          const cd ax = sm.AX[tid + BIG_CH * c], mm = Mg[p], yo = Yg[p];
          cd cc = cmk(fma(mm.x, imu, ax.x), fma(mm.y, imu, ax.y));
          double D = sqrt(cabs2(cc));
          if (D == 0.0) { cc = cmk(1.0, 0.0); D = 1.0; }
          const double f = (sm.Bs[i] / D + mu) * i1mu;
          const cd yn = cscale(cc, f);
          const cd jm = cmk(ax.x - yn.x, ax.y - yn.y), dy = cmk(yn.x - yo.x, yn.y - yo.y);
          Yg[p] = yn;
          Mg[p] = cmk(fma(mu, jm.x, mm.x), fma(mu, jm.y, mm.y));
          if (prm.need_dual) dYg[p] = dy;
          pYd2 += cabs2(dy); pJM2 += cabs2(jm); pY2 += cabs2(yn);
          const double a2 = cabs2(ax);
          pAX2 += a2;
          const double dd = sqrt(a2) - sm.Bs[i];
          ocol[c] += dd * dd;
        }
      }
      __syncthreads();
    }, red);
    {
      double v[6] = {pYd2, pJM2, pY2, pAX2, obj2, nAX2};
      block_sum<6>(v, sm.red);
      pYd2 = v[0]; pJM2 = v[1]; pY2 = v[2]; pAX2 = v[3]; obj2 = v[4]; nAX2 = v[5];
    }
    if (!tk.sbr) {
      block_sum<RL>(ocol, sm.red);
      if (tid < RL) xcol[c0 + tid] = sqrt(ocol[tid]);
    }
    // ---- A'(Y - Y0) (:309): advances AtY in global memory; only feeds res_dual
    double pAtYd2 = 0.0, pAtY2 = 0.0;
    if (prm.need_dual) {
      __threadfence_block();
      __syncthreads();
      double v[2] = {0.0, 0.0};
      big_product_ah(tc, m, mt, [&](int i, int c) { return dYg[i + (size_t)m * (c0 + c)]; },
                     [&](int k, int c, cd acc) {
                       const size_t p = k + (size_t)FN * (c0 + c);
                       const cd d = cscale(acc, cs);
                       cd a = AtY[p];
                       a.x += d.x;
                       a.y += d.y;
                       AtY[p] = a;
                       v[0] += cabs2(d);
                       v[1] += cabs2(a);
                     }, red);
      block_sum<2>(v, sm.red);
      pAtYd2 = v[0]; pAtY2 = v[1];
    }
    const long long tx2 = clock64();
    if (tid == 0) sm.sc[22] += (double)(tx2 - tx1);
    // ---- Z, N update (:312, :319-320), warm-started eigenbasis with a cold restart every 64 decompositions
    fast_argmin_z<RL, CS>(m, rank_one, sm, mu, false, (sm.ifl[2] & 63) != 0, nz, &sweeps);
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
        if (tk.Yout) for (int idx = tid; idx < m * RL; idx += NT) tk.Yout[(size_t)m * c0 + idx] = Yg[(size_t)m * c0 + idx];
      } else if (jbest / RL == rank) {
        const int cl = jbest - c0;
        if (tk.Xout) for (int k = tid; k < FN; k += NT) tk.Xout[k] = sm.X[k + FN * cl];
        if (tk.Yout) for (int i = tid; i < m; i += NT) tk.Yout[i] = Yg[i + (size_t)m * jbest];
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
  if (tk.state) {   // [X Z N (n x r) | Y M (m x r)], own columns
    cd* st = tk.state;
    const size_t nr = (size_t)FN * r, mr = (size_t)m * r;
    for (int idx = tid; idx < FN * RL; idx += NT) {
      const size_t g = (size_t)FN * c0 + idx;
      st[g] = sm.X[idx]; st[nr + g] = sm.Z[idx]; st[2 * nr + g] = sm.N[idx];
    }
    for (int idx = tid; idx < m * RL; idx += NT) {
      const size_t g = (size_t)m * c0 + idx;
      st[3 * nr + g] = Yg[g]; st[3 * nr + mr + g] = Mg[g];
    }
  }
  if (tk.scal && tid == 0 && rank == 0) {
    tk.scal[SC_MU] = mu; tk.scal[SC_OPT_OBJ] = opt_obj; tk.scal[SC_ITERS] = iters;
    tk.scal[SC_OPT_ITER] = opt_iter; tk.scal[SC_OPT_COL] = opt_col; tk.scal[SC_BUMPS] = bumps;
    tk.scal[SC_CONVERGED] = converged; tk.scal[SC_RES_COMB] = res_comb; tk.scal[SC_SWEEPS] = sweeps;
    tk.scal[9] = sm.sc[20]; tk.scal[10] = sm.sc[21]; tk.scal[11] = (double)(clock64() - tl0);
    tk.scal[12] = sm.sc[22]; tk.scal[13] = sm.sc[23]; tk.scal[14] = sm.sc[24];
    tk.scal[15] = (double)(tl0 - ttask0);
  }
  (void)lane; (void)warp;
  cl_sync<CS>();   // no CTA leaves (or reuses its exchange buffers) while a peer may still read them
}

__global__ void __launch_bounds__(NT, 1)
big_stage_kernel(const StageTask* __restrict__ tasks, int ntasks, DevParams prm, FastDims fd, cd* wsbase, int* counter) {
  extern __shared__ __align__(1024) unsigned char big_smem_raw[];
  __shared__ int s_next_task;
  const FastSmem<BIG_RL> sm = fast_carve<BIG_RL>(big_smem_raw, fd);
  const int rank = (int)cg::this_cluster().block_rank();
  const int cid = blockIdx.x / BIG_CS, ncl = gridDim.x / BIG_CS;
  cd* wsg = wsbase + (size_t)cid * fd.ws_stride;
  TcCtx tc = {};
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
  for (int t = cid;; t += ncl) {      // dynamic task queue, see fast_stage_kernel
    if (counter != nullptr) t = next_task<BIG_CS>(counter, &s_next_task, rank);
    if (t >= ntasks) break;
    const StageTask tk = tasks[t];
    if (tk.active != nullptr && *tk.active != tk.active_expect) continue;   // cluster-uniform
    run_big(tk, prm, fd, sm, wsg, rank, tc);
  }
  umma::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) umma::tmem_free512(tc.tmem);
}


// ------------------------------------------------------------------------------------------------------------------
// r = 1 stages for 256 < m <= 1024: the refinement on all rows (inferLowRankV4.m:68-80).  One CTA per task, two CTAs
// per SM; the iterate (X, Z, N: 256; Y, M, Y - Y0, A X: m) lives in shared memory, U = (A'A + I)^-1 in L2, the 2-bit
// codes are read straight from global memory (64 B per row, L1/L2-resident).  Both sensing-matrix products are
// multiply-free: u in {1, j, -1, -j} turns every term into a swap / negate and two additions.

// acc += j^c v
__device__ __forceinline__ void cadd_rot(cd& acc, uint32_t c, cd v) {
  const bool odd = (c & 1u) != 0u;
  double a = odd ? -v.y : v.x;
  double b = odd ? v.x : v.y;
  if (c & 2u) { a = -a; b = -b; }
  acc.x += a;
  acc.y += b;
}

// out(i, sum_k u(i,k) x(k)) for i < m: thread t owns rows t, t + 256, ... (up to four), every x(k) read from shared
// memory once per thread and used for all of its rows.
template <class OutF>
__device__ __forceinline__ void big1_prod_a(const uint32_t* __restrict__ codes, const int* rows_s, int m, const cd* x, OutF out) {
  const int tid = threadIdx.x;
  const int nq = (m + NT - 1) / NT;
  const uint4* cr[4];
  cd acc[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = tid + NT * q;
    cr[q] = reinterpret_cast<const uint4*>(codes + (size_t)rows_s[i < m ? i : 0] * 16);
    acc[q] = cmk(0.0, 0.0);
  }
  for (int w4 = 0; w4 < 4; ++w4) {
    uint4 cw[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) cw[q] = (q < nq) ? __ldg(cr[q] + w4) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int ww = 0; ww < 4; ++ww) {
      uint32_t wd[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) wd[q] = ww == 0 ? cw[q].x : ww == 1 ? cw[q].y : ww == 2 ? cw[q].z : cw[q].w;
      const cd* xv = x + 64 * w4 + 16 * ww;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const cd v = xv[k];
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (q < nq) cadd_rot(acc[q], (wd[q] >> (2 * k)) & 3u, v);
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = tid + NT * q;
    if (i < m) out(i, acc[q]);
  }
}

// out(k, sum_i conj(u(i,k)) in(i)) for k < 256.  Warp w owns the code words 2w, 2w + 1 (k in [32 w, 32 w + 32)); its
// lanes are 16 row subsets x 2 words, lane subset s takes rows s, s + 16, ...; the 16 partial sums of every k are
// combined by a shuffle reduce-scatter, after which the lane of subset s holds k = 16 word + s.
template <class InF, class OutF>
__device__ __forceinline__ void big1_prod_ah(const uint32_t* __restrict__ codes, const int* rows_s, int m, InF in, OutF out) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int s = lane & 15, w = 2 * warp + (lane >> 4);
  cd acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = cmk(0.0, 0.0);
  for (int i = s; i < m; i += 16) {
    const uint32_t wd = __ldg(codes + (size_t)rows_s[i] * 16 + w);
    const cd t = in(i);
#pragma unroll
    for (int j = 0; j < 16; ++j) cadd_rot(acc[j], (0u - (wd >> (2 * j))) & 3u, t);      // conj(j^c) = j^(-c)
  }
#pragma unroll
  for (int h = 8; h >= 1; h >>= 1) {
    const bool up = (s & h) != 0;
#pragma unroll
    for (int j = 0; j < h; ++j) {
      const cd mine = up ? acc[j + h] : acc[j];
      const cd give = up ? acc[j] : acc[j + h];
      acc[j] = cmk(mine.x + __shfl_xor_sync(0xffffffffu, give.x, h), mine.y + __shfl_xor_sync(0xffffffffu, give.y, h));
    }
  }
  out(16 * w + s, acc[0]);
}

__device__ inline void run_big1(const StageTask& tk, const DevParams& prm, const FastSmem<1>& sm, cd* wsg) {
  const int tid = threadIdx.x;
  const int m = tk.m;
  const double cs = *tk.cscale, bsc = *tk.bscale;
  const int rank_one = tk.rank_one_ptr ? *tk.rank_one_ptr : tk.rank_one;
  const long long ttask0 = clock64();
  cd* U = wsg;                        // (A'A + I)^-1
  cd* AtY = wsg + (size_t)FN * FN;    // [256]
  const uint32_t* codes = tk.codes;
  for (int i = tid; i < m; i += NT) {
    sm.rows_s[i] = tk.A.rows ? tk.A.rows[i] : i;
    sm.Bs[i] = bsc * tk.B[tk.brows ? tk.brows[i] : i];
  }
  jacobi_tables<FTX>(sm.pairs);
  __syncthreads();
  double nb2;
  {
    double v[1] = {0.0};
    for (int i = tid; i < m; i += NT) v[0] += sm.Bs[i] * sm.Bs[i];
    block_sum<1>(v, sm.red);
    nb2 = v[0];
  }
  const double normB = sqrt(nb2);
  // ---- U = (A'A + I)^-1  (:220-222): column l of A'A is A' u(:, l), exact in FP64 (sums of 1, j, -1, -j)
  {
    const double cs2 = cs * cs;
    for (int l = 0; l < FN; ++l) {
      const int wl = l >> 4, sh = 2 * (l & 15);
      big1_prod_ah(codes, sm.rows_s, m,
                   [&](int i) {
                     const uint32_t c = (__ldg(codes + (size_t)sm.rows_s[i] * 16 + wl) >> sh) & 3u;
                     return cmk((double)((c == 0u) - (c == 2u)), (double)((c == 1u) - (c == 3u)));
                   },
                   [&](int k, cd v) { U[k + (size_t)FN * l] = cmk(fma(cs2, v.x, k == l ? 1.0 : 0.0), cs2 * v.y); });
    }
    __threadfence_block();
    __syncthreads();
    spd_inverse(U, FN, sm.WT, sm.WT + FN);
    __syncthreads();
  }
  // ---- X = X0, M = N = 0, rescale so |A X| matches |B| (:278-286), Y = normalize_rows(A X, B) (:287)
  for (int k = tid; k < FN; k += NT) { sm.X[k] = tk.X0[k]; sm.N[k] = cmk(0.0, 0.0); }
  for (int i = tid; i < m; i += NT) sm.M[i] = cmk(0.0, 0.0);
  __syncthreads();
  big1_prod_a(codes, sm.rows_s, m, sm.X, [&](int i, cd v) { sm.AX[i] = cscale(v, cs); });
  __syncthreads();
  {
    double v[1] = {0.0};
    for (int i = tid; i < m; i += NT) v[0] += cabs2(sm.AX[i]);
    block_sum<1>(v, sm.red);
    const double s = normB / sqrt(v[0]);
    for (int k = tid; k < FN; k += NT) sm.X[k] = cscale(sm.X[k], s);
    for (int i = tid; i < m; i += NT) {
      cd a = cscale(sm.AX[i], s);
      double D = sqrt(cabs2(a));
      if (D == 0.0) { a = cmk(1.0, 0.0); D = 1.0; }
      sm.Y[i] = cscale(a, sm.Bs[i] / D);
    }
  }
  __syncthreads();
  int sweeps = 0;
  double nz[4];
  if (tid == 0) sm.ifl[2] = 0;
  __syncthreads();
  fast_argmin_z<1, 1>(m, rank_one, sm, 1.0, true, false, nz, &sweeps);     // Z = ArgMinZ(X, 0, 1)  (:288)
  __syncthreads();
  if (prm.need_dual)     // AtY = A' Y  (:289)
    big1_prod_ah(codes, sm.rows_s, m, [&](int i) { return sm.Y[i]; }, [&](int k, cd v) { AtY[k] = cscale(v, cs); });

  double mu = prm.mu0, opt_obj = INFINITY, last_res = INFINITY, res_comb = 0.0;
  int iters = 0, opt_iter = -1, opt_col = -1, bumps = 0, converged = 0, have_opt = 0;
  if (tid == 0) { sm.sc[20] = 0.0; sm.sc[21] = 0.0; sm.sc[22] = 0.0; sm.sc[23] = 0.0; sm.sc[24] = 0.0; }
  __syncthreads();
  const long long tl0 = clock64();

  for (int it = 1; it <= prm.maxiter; ++it) {
    const double imu = 1.0 / mu, i1mu = 1.0 / (1.0 + mu);
    const long long tx0 = clock64();
    // ---- X update (:304, :380-388): X = U (A'(Y - M/mu) + Z - N/mu)
    big1_prod_ah(codes, sm.rows_s, m,
                 [&](int i) {
                   const cd y = sm.Y[i], mm = sm.M[i];
                   return cmk(fma(-mm.x, imu, y.x), fma(-mm.y, imu, y.y));
                 },
                 [&](int k, cd v) {
                   const cd z = sm.Z[k], nn = sm.N[k];
                   sm.X[k] = cmk(fma(v.x, cs, fma(-nn.x, imu, z.x)), fma(v.y, cs, fma(-nn.y, imu, z.y)));
                 });
    __syncthreads();
    prod_sq_inplace<1>(U, sm.X);
    const long long tx1 = clock64();
    if (tid == 0) sm.sc[21] += (double)(tx1 - tx0);
    // ---- A X (:305), Y update (:308), M update (:315-316), objective (:323-340); Y - Y0 kept in WT for A'(Y - Y0)
    big1_prod_a(codes, sm.rows_s, m, sm.X, [&](int i, cd v) { sm.AX[i] = cscale(v, cs); });
    double pYd2 = 0.0, pJM2 = 0.0, pY2 = 0.0, pAX2 = 0.0, obj2 = 0.0;
    for (int i = tid; i < m; i += NT) {       // every thread touches only the rows it has just written
      const cd ax = sm.AX[i], mm = sm.M[i], yo = sm.Y[i];
      cd cc = cmk(fma(mm.x, imu, ax.x), fma(mm.y, imu, ax.y));
      double D = sqrt(cabs2(cc));
      if (D == 0.0) { cc = cmk(1.0, 0.0); D = 1.0; }
      const double f = (sm.Bs[i] / D + mu) * i1mu;
      const cd yn = cscale(cc, f);
      const cd jm = cmk(ax.x - yn.x, ax.y - yn.y), dy = cmk(yn.x - yo.x, yn.y - yo.y);
      sm.Y[i] = yn;
      sm.M[i] = cmk(fma(mu, jm.x, mm.x), fma(mu, jm.y, mm.y));
      sm.WT[i] = dy;
      pYd2 += cabs2(dy); pJM2 += cabs2(jm); pY2 += cabs2(yn);
      const double a2 = cabs2(ax);
      pAX2 += a2;
      const double dd = sqrt(a2) - sm.Bs[i];
      obj2 += dd * dd;
    }
    {
      double v[5] = {pYd2, pJM2, pY2, pAX2, obj2};
      block_sum<5>(v, sm.red);
      pYd2 = v[0]; pJM2 = v[1]; pY2 = v[2]; pAX2 = v[3]; obj2 = v[4];
    }
    // ---- A'(Y - Y0) (:309): only feeds res_dual
    double pAtYd2 = 0.0, pAtY2 = 0.0;
    if (prm.need_dual) {
      double v[2] = {0.0, 0.0};
      big1_prod_ah(codes, sm.rows_s, m, [&](int i) { return sm.WT[i]; },
                   [&](int k, cd acc) {
                     const cd d = cscale(acc, cs);
                     cd a = AtY[k];
                     a.x += d.x;
                     a.y += d.y;
                     AtY[k] = a;
                     v[0] += cabs2(d);
                     v[1] += cabs2(a);
                   });
      block_sum<2>(v, sm.red);
      pAtYd2 = v[0]; pAtY2 = v[1];
    }
    const long long tx2 = clock64();
    if (tid == 0) sm.sc[22] += (double)(tx2 - tx1);
    // ---- Z, N update (:312, :319-320)
    fast_argmin_z<1, 1>(m, rank_one, sm, mu, false, (sm.ifl[2] & 63) != 0, nz, &sweeps);
    const long long tx3 = clock64();
    if (tid == 0) sm.sc[23] += (double)(tx3 - tx2);
    const double nJN2 = nz[0], nZd2 = nz[1], nX2 = nz[2], nZ2 = nz[3];
    // ---- best solution so far (:323-340); a NaN objective never wins
    const double obj = sqrt(obj2);
    if (obj < opt_obj) {
      opt_obj = obj; opt_iter = it; opt_col = tk.sbr ? -1 : 0; have_opt = 1;
      if (tk.Xout) for (int k = tid; k < FN; k += NT) tk.Xout[k] = sm.X[k];
      if (tk.Yout) for (int i = tid; i < m; i += NT) tk.Yout[i] = sm.Y[i];
    }
    // ---- residuals and stopping rule (:343-354)
    const double res_prim = sqrt(pJM2 + nJN2);
    const double res_dual = mu * sqrt(pAtYd2 + nZd2);
    res_comb = sqrt(pJM2 + nJN2 + pYd2 + nZd2);
    iters = it;
    if (tk.trace != nullptr && tid == 0) tk.trace[it - 1] = res_comb;
    if (prm.need_dual) {
      const double mx1 = fmax(sqrt(pAX2), sqrt(pY2)), mx2 = fmax(sqrt(nX2), sqrt(nZ2));
      const double th_prim = prm.tol_abs * sqrt((double)(m + FN)) + prm.tol_rel * sqrt(mx1 * mx1 + mx2 * mx2);
      const double th_dual = prm.tol_abs * sqrt((double)FN * 2.0) + prm.tol_rel * sqrt(pAtY2 + nZ2);
      const double th_comb = prm.tol_abs * sqrt((double)(m + FN) * 2.0) + prm.tol_rel * sqrt(mx1 * mx1 + mx2 * mx2 + pY2 + nZ2);
      if ((res_prim < th_prim && res_dual < th_dual) || (res_comb < th_comb)) { converged = 1; break; }
    }
    if (res_comb > last_res * 0.9) { mu *= prm.rho; ++bumps; }   // :358-361
    last_res = res_comb;
    __syncthreads();
    if (tid == 0) sm.sc[24] += (double)(clock64() - tx3);
  }
  __syncthreads();
  if (!have_opt) {    // all-NaN objectives give NaN (H4)
    if (tk.Xout) for (int k = tid; k < FN; k += NT) tk.Xout[k] = cmk(NAN, NAN);
    if (tk.Yout) for (int i = tid; i < m; i += NT) tk.Yout[i] = cmk(NAN, NAN);
  }
  if (tk.state) {   // [X Z N (n) | Y M (m)]
    cd* st = tk.state;
    for (int k = tid; k < FN; k += NT) { st[k] = sm.X[k]; st[FN + k] = sm.Z[k]; st[2 * FN + k] = sm.N[k]; }
    for (int i = tid; i < m; i += NT) { st[3 * FN + i] = sm.Y[i]; st[3 * FN + m + i] = sm.M[i]; }
  }
  if (tk.scal && tid == 0) {
    tk.scal[SC_MU] = mu; tk.scal[SC_OPT_OBJ] = opt_obj; tk.scal[SC_ITERS] = iters;
    tk.scal[SC_OPT_ITER] = opt_iter; tk.scal[SC_OPT_COL] = opt_col; tk.scal[SC_BUMPS] = bumps;
    tk.scal[SC_CONVERGED] = converged; tk.scal[SC_RES_COMB] = res_comb; tk.scal[SC_SWEEPS] = sweeps;
    tk.scal[9] = sm.sc[20]; tk.scal[10] = sm.sc[21]; tk.scal[11] = (double)(clock64() - tl0);
    tk.scal[12] = sm.sc[22]; tk.scal[13] = sm.sc[23]; tk.scal[14] = sm.sc[24];
    tk.scal[15] = (double)(tl0 - ttask0);
  }
  __syncthreads();
}

__host__ __device__ inline size_t big1_ws_elems() { return (size_t)FN * FN + FN; }

__global__ void __launch_bounds__(NT, 2)
big1_stage_kernel(const StageTask* __restrict__ tasks, int ntasks, DevParams prm, FastDims fd, cd* wsbase) {
  extern __shared__ __align__(1024) unsigned char big1_smem_raw[];
  const FastSmem<1> sm = fast_carve<1>(big1_smem_raw, fd);
  cd* wsg = wsbase + (size_t)blockIdx.x * fd.ws_stride;
  for (int t = blockIdx.x; t < ntasks; t += gridDim.x) {
    const StageTask tk = tasks[t];
    if (tk.active != nullptr && *tk.active != tk.active_expect) continue;
    run_big1(tk, prm, sm, wsg);
  }
}

}  // namespace twoace
