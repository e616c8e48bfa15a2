// Exact-integer tensor-core form of the two sensing-matrix products of an InferADMM iteration
// (inferLowRankV4.m:304-305: A'(.) inside ArgMinX and A*X), for quantised sensing matrices A = c * u,
// u in {1, j, -1, -j} (every shipped codebook).
//
// Idea (Ozaki-style splitting, the product itself is exact):
//   * A is stored once as two int8 matrices Re u, Im u (entries in {0, +-1}) in 128 x 128-byte operand blocks
//     (umma_i8.cuh) that tcgen05 can read K-major (A X) and MN-major (A' T) -- one copy serves both products;
//   * every column of the FP64 operand is scaled by a power of two so that its largest component is below 64
//     and split into TC_S = 8 signed 7-bit digits (balanced, |digit| <= 64): x * 2^(5-e) = sum_s d_s 128^-s with
//     a remainder below 2^-56 of the column maximum; the digits are the int8 B operand;
//   * tcgen05.mma kind::i8 accumulates the digit products exactly in int32 tensor memory
//     (|sum| <= 256 * 64 < 2^15), four accumulators (Re u, Im u) x (output tile 0, 1);
//   * the epilogue recombines  Re = D_re[x_re] -+ D_im[x_im],  Im = D_re[x_im] +- D_im[x_re]  in int32 and
//     evaluates the digit polynomial in FP64 (one rounding per digit, relative to the result).
// The result equals the exactly rounded product of A with the operand truncated at 2^-56 of its column
// maximum: the error is below that of an FP64 accumulation of the same 256-term sums.
// Non-finite operand columns give NaN results (an FP64 product would: every entry of A is non-zero).
#pragma once
#include "common.cuh"
#include "umma_i8.cuh"

namespace twoace {

constexpr int TC_S = 8;                       // digits per real number
constexpr int TC_NC = 5;                      // operand columns per MMA pass
constexpr int TC_N = 2 * TC_NC * TC_S;        // 80 = N of the MMA: (column, re|im, digit)
constexpr int TC_KB = 256;                    // K extent of the B operand (k or i in [0, 256))
constexpr int TC_BS_BYTES = TC_N * TC_KB;     // 20480
constexpr int TC_MAXSLOT = 12;
constexpr int TC_AOP_BYTES = 8 * 16384;       // operand blocks of one instance (m <= 256): 2 parts x 2 tiles x 2 halves

// launch-wide ring geometry (host computed, see fast_tc_layout in fast_stage.cuh)
struct TcDims {
  int on;           // 1: tensor-core products
  int slot_bytes;   // ring slot = 128 * (rows per block of the largest m in the launch)
  int nslot;        // ring slots: n1 in the overlay region (G | P | xG, dead outside ArgMinZ) + the rest
  int n1;
  int ov_bytes;     // size of the overlay region
};

// per-CTA state that persists over the tasks of a launch
struct TcCtx {
  unsigned char* Bs;      // B operand (aliases the setup-time code copy)
  unsigned char* ov;      // overlay region: ring slots [0, n1)
  unsigned char* ex;      // ring slots [n1, nslot)
  uint64_t* bars;         // full[TC_MAXSLOT], empty[TC_MAXSLOT], done
  unsigned char* aop;     // global operand blocks of this cluster (streaming mode)
  uint32_t tmem;
  uint32_t pf, pe, pd;    // phase bits: full / empty (thread 0), done (all threads)
  int nslot, n1, slot_bytes;   // nslot: slots of the CURRENT task (streaming: at most NW - 1, one producer warp each)
  int nslot_launch;            // ring slots carved for the launch (without the WT alias)
  uint32_t premask;       // slots whose first-round copy of the NEXT product is already in flight (streaming mode)
  unsigned char* wt;      // the WT buffer: one more ring slot while it is dead (wt_slot == nslot - 1), or wt_slot = -1
  int wt_slot;
  __device__ __forceinline__ unsigned char* slot(int s) const {
    if (s == wt_slot) return wt;
    return s < n1 ? ov + (size_t)s * slot_bytes : ex + (size_t)(s - n1) * slot_bytes;
  }
};

// geometry of one instance: rows per block, 128-row tiles, blocks, residency
struct TcGeom {
  int R, mt, nblk, bb;
  bool resident;   // every block has its own ring slot (s0 + block id), filled once per stage
  int s0;
};
__device__ __forceinline__ TcGeom tc_geom(int m, int nslot, int n1) {
  TcGeom g;
  g.mt = (m + 127) >> 7;
  g.R = g.mt > 1 ? 128 : ((m + 31) >> 5) << 5;
  g.nblk = 4 * g.mt;
  g.bb = g.R * 128;
  g.resident = nslot - n1 >= g.nblk;   // the overlay slots are clobbered by ArgMinZ: resident blocks live behind them
  g.s0 = g.resident ? n1 : 0;
  return g;
}
// block (p, it, ks): part p (0 = Re u, 1 = Im u), rows [128 it, 128 it + R), bytes k in [128 ks, 128 ks + 128)
__device__ __forceinline__ int tc_block_id(int p, int it, int ks, int mt) { return (p * mt + it) * 2 + ks; }

// Build the operand blocks of one instance from its 2-bit codes cik[w * m + i] (16 codes of row i, k = 16 w + q).
// resident: every CTA fills its own ring; otherwise the cluster writes them once to tc.aop (global memory).
// Ends with the fences that make the bytes visible to the async proxy; the CALLER synchronises (block / cluster).
__device__ inline void tc_build(const TcCtx& tc, const TcGeom& g, const uint32_t* cik, int m, int part, int nparts) {
  const int items = g.mt * g.R * 16;
  for (int idx = threadIdx.x + NT * part; idx < items; idx += NT * nparts) {
    const int il = idx % g.R, t2 = idx / g.R, it = t2 % g.mt, w = t2 / g.mt;
    const int i = 128 * it + il;
    uint32_t re[4] = {0u, 0u, 0u, 0u}, im[4] = {0u, 0u, 0u, 0u};
    if (i < m) {
      const uint32_t word = cik[w * m + i];
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const uint32_t c = (word >> (2 * q)) & 3u;
        re[q >> 2] |= ((0x00FF0001u >> (8 * c)) & 0xFFu) << (8 * (q & 3));   // 1, 0, -1, 0
        im[q >> 2] |= ((0xFF000100u >> (8 * c)) & 0xFFu) << (8 * (q & 3));   // 0, 1, 0, -1
      }
    }
    const int ks = w >> 3;
    const size_t inblk = (size_t)(w & 7) * (g.R * 16) + (size_t)il * 16;
    const int b0 = tc_block_id(0, it, ks, g.mt), b1 = tc_block_id(1, it, ks, g.mt);
    unsigned char* d0 = (g.resident ? tc.slot(g.s0 + b0) : tc.aop + (size_t)b0 * g.bb) + inblk;
    unsigned char* d1 = (g.resident ? tc.slot(g.s0 + b1) : tc.aop + (size_t)b1 * g.bb) + inblk;
    *reinterpret_cast<uint4*>(d0) = make_uint4(re[0], re[1], re[2], re[3]);
    *reinterpret_cast<uint4*>(d1) = make_uint4(im[0], im[1], im[2], im[3]);
  }
  if (g.resident) {
    umma::fence_async_smem();
  } else {
    asm volatile("fence.proxy.async;" ::: "memory");
    __threadfence();
  }
}

// ---- building blocks of a product ---------------------------------------------------------------------------
// Column scales: x * sc has its largest component below 64 (sc = 2^(5-e)); inv = 1 / sc; bad = non-finite column.
struct TcScale {
  double sc[TC_NC], inv[TC_NC];
  bool bad[TC_NC];
};
// h[c]: this thread's maximum over the hi words (sign cleared) of the components it contributes to column c.
// Every thread of the CTA calls it; contains one block barrier; `red` is shared scratch of >= NW * TC_NC words
// that must not be rewritten before the next block barrier.
__device__ __forceinline__ void tc_scales(const uint32_t (&h)[TC_NC], uint32_t* red, TcScale& S) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < TC_NC; ++c) {
    const uint32_t hm = __reduce_max_sync(0xffffffffu, h[c]);
    if (lane == 0) red[warp * TC_NC + c] = hm;
  }
  __syncthreads();
#pragma unroll
  for (int c = 0; c < TC_NC; ++c) {
    uint32_t hm = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) hm = max(hm, red[w * TC_NC + c]);
    const int ef = (int)(hm >> 20);              // exponent field of the largest |component|
    S.bad[c] = ef == 0x7FF;
    const bool dead = S.bad[c] || ef < 6;        // (columns below 2^-1017 are flushed to zero)
    S.sc[c] = dead ? 0.0 : __hiloint2double((2051 - ef) << 20, 0);   // 2^(5 - e)
    S.inv[c] = dead ? 0.0 : __hiloint2double((ef - 5) << 20, 0);     // 2^(e - 5)
  }
}
__device__ __forceinline__ uint32_t tc_hi(cd v) {
  return max((uint32_t)__double2hiint(v.x) & 0x7FFFFFFFu, (uint32_t)__double2hiint(v.y) & 0x7FFFFFFFu);
}
// The 2 x TC_S digits of operand row `kb` (the K index of the MMA) for every column -> B operand.
__device__ __forceinline__ void tc_slice_row(unsigned char* Bs, int kb, const cd (&x)[TC_NC], const TcScale& S) {
  unsigned char* dst = Bs + (size_t)(kb >> 4) * (TC_N * 16) + (kb & 15);
#pragma unroll
  for (int c = 0; c < TC_NC; ++c) {
#pragma unroll
    for (int part = 0; part < 2; ++part) {
      double y = S.bad[c] ? 0.0 : (part ? x[c].y : x[c].x) * S.sc[c];
#pragma unroll
      for (int s = 0; s < TC_S; ++s) {
        const double t = y + 6755399441055744.0;          // 1.5 * 2^52: rint(y) in the low mantissa bits
        dst[((c * 2 + part) * TC_S + s) * 16] = (unsigned char)(__double2loint(t) & 0xFF);
        y = (y - (t - 6755399441055744.0)) * 128.0;       // exact
      }
    }
  }
}

// Block schedule of one MMA pass over nt (1 or 2) consecutive 128-row tiles of A starting at tile t0 (of mt):
//   A X : step j -> (output tile it = j / 4, part p, k half ks):   block (p, t0 + it, ks),  4 nt steps
//   A' T: step j -> (output tile kt, part p, row tile is):          block (p, t0 + is, kt),  4 nt steps
// nt is 1 or 2, so the index arithmetic needs no division.  Returns the block id; kblk = position along K.
template <bool AH>
__device__ __forceinline__ int tc_seq(int j, int nt, int t0, int mt, int& tile_out, int& p, int& kblk) {
  const int sh = nt - 1;   // log2(nt)
  if (!AH) { tile_out = j >> 2; p = (j >> 1) & 1; kblk = j & 1; return tc_block_id(p, t0 + tile_out, kblk, mt); }
  tile_out = j >> (1 + sh); p = (j >> sh) & 1; kblk = j & sh;
  return tc_block_id(p, t0 + kblk, tile_out, mt);
}

struct TcPass {
  int R;        // rows per slab of a block
  int nt;       // tiles of this pass (1 or 2)
  int t0, mt;   // first tile of the pass, tiles of the instance
  int rows;     // A' T: valid operand rows from tile t0 on (K steps beyond them are skipped)
  bool resident;
  int s0;
  bool acc_first;   // A' T: the accumulators already hold the sum over earlier passes
  __device__ __forceinline__ int nblk() const { return 4 * nt; }
  __device__ __forceinline__ uint32_t bb() const { return (uint32_t)R * 128u; }
};
__device__ __forceinline__ TcPass tc_pass_of(const TcGeom& g, int m) {
  TcPass ps;
  ps.R = g.R; ps.nt = g.mt; ps.t0 = 0; ps.mt = g.mt; ps.rows = m; ps.resident = g.resident; ps.s0 = g.s0;
  ps.acc_first = false;
  return ps;
}

// ---- streaming mode -------------------------------------------------------------------------------------
// Ring slot s is refilled by its own producer warp (warp s + 1, one elected lane issues): issuing a bulk copy costs
// the issuing thread a few hundred cycles, so one thread refilling every slot would serialise the copies.  Warp 0
// issues the MMAs and commits a slot's "empty" barrier after the block in it has been multiplied.
// Step j of a pass uses slot j % nslot.
// first-round copies (step j = slot s) of a pass; `skip_wt`: leave out the slot that aliases WT
template <bool AH>
__device__ __forceinline__ void tc_first_round(TcCtx& tc, const TcPass& ps, bool skip_wt) {
  using namespace umma;
  const int s = uniform_warp_id() - 1;     // warp-uniform: producer warp of slot s
  if (ps.resident || s < 0 || s >= tc.nslot || s >= ps.nblk()) return;
  if ((tc.premask >> s) & 1u) return;
  if (skip_wt && s == tc.wt_slot) return;
  int a, b, c;
  const int blk = tc_seq<AH>(s, ps.nt, ps.t0, ps.mt, a, b, c);
  if (elect_one()) {
    mbar_expect_tx(tc.bars + s, ps.bb());
    bulk_g2s(tc.slot(s), tc.aop + (size_t)blk * ps.bb(), ps.bb(), tc.bars + s);
  }
  tc.premask |= 1u << s;
}

// Start the first-round copies of the NEXT product while other work runs (every thread calls it).  The ring must be
// free (the previous product has returned); the WT slot is left out (WT is live between the products).
template <bool AH>
__device__ __forceinline__ void tc_prefetch(TcCtx& tc, const TcGeom& g, int m) {
  const TcPass ps = tc_pass_of(g, m);
  tc_first_round<AH>(tc, ps, true);
}

// The MMAs of one pass: B operand complete in tc.Bs (fenced, block-synchronised by the caller), accumulators
// (output tile, part) at tensor-memory columns (tile * 2 + part) * TC_N.  Every thread of the CTA calls it; returns
// after the MMAs have completed (warp 0 polls the mbarrier; the caller's next block barrier releases the others).
template <bool AH>
__device__ __forceinline__ void tc_mma_pass(TcCtx& tc, const TcPass& ps) {
  using namespace umma;
  const int uwarp = uniform_warp_id();
  const int nblk = ps.nblk();
  // ---- producers (streaming): whole warps run these loops (warp-uniform control flow, see elect_one())
  if (!ps.resident && uwarp >= 1 && uwarp - 1 < tc.nslot) {
    const int s = uwarp - 1;
    if (s < nblk && !((tc.premask >> s) & 1u)) {   // first round not issued yet (e.g. the slot that aliases WT)
      int a, b, c;
      const int blk = tc_seq<AH>(s, ps.nt, ps.t0, ps.mt, a, b, c);
      if (elect_one()) {
        mbar_expect_tx(tc.bars + s, ps.bb());
        bulk_g2s(tc.slot(s), tc.aop + (size_t)blk * ps.bb(), ps.bb(), tc.bars + s);
      }
    }
    for (int j = s + tc.nslot; j < nblk; j += tc.nslot) {
      mbar_wait(tc.bars + TC_MAXSLOT + s, (tc.pe >> s) & 1u);   // the block that was in the slot has been multiplied
      tc.pe ^= 1u << s;
      int a, b, c;
      const int blk = tc_seq<AH>(j, ps.nt, ps.t0, ps.mt, a, b, c);
      if (elect_one()) {
        mbar_expect_tx(tc.bars + s, ps.bb());
        bulk_g2s(tc.slot(s), tc.aop + (size_t)blk * ps.bb(), ps.bb(), tc.bars + s);
      }
    }
  }
  tc.premask = 0;
  // ---- MMAs (warp 0, one elected lane issues): every block is 4 (or fewer) K = 32 steps
  if (uwarp == 0) {
    tc_fence_after();
    const uint32_t idesc = idesc_i8(TC_N, AH, false);
    const uint32_t slab = (uint32_t)ps.R * 16u;
    const uint64_t bd0 = smem_desc(smem_u32(tc.Bs), TC_N * 16, 128);
    const uint64_t ainc = AH ? 32u : (uint64_t)(slab >> 3);   // per K = 32 step: 32 rows x 16 B  |  2 slabs
    int s = 0;
    for (int j = 0; j < nblk; ++j) {
      int tile_out, p, kblk;
      const int blk = tc_seq<AH>(j, ps.nt, ps.t0, ps.mt, tile_out, p, kblk);
      if (ps.resident) {
        s = ps.s0 + blk;
      } else {
        mbar_wait(tc.bars + s, (tc.pf >> s) & 1u);
        tc.pf ^= 1u << s;
        tc_fence_after();
      }
      // descriptors differ from step to step only in their start-address field (bits 0-13, in 16-byte units)
      const uint32_t sa = smem_u32(tc.slot(s));
      const int nst = AH ? (min(128, ps.rows - 128 * kblk) + 31) >> 5 : 4;
      const uint32_t dcol = tc.tmem + (uint32_t)((tile_out * 2 + p) * TC_N);
      uint64_t ad = AH ? smem_desc(sa, 128, slab) : smem_desc(sa, slab, 128);
      uint64_t bd = bd0 + (uint64_t)(kblk * 8 * TC_N);          // K offset of this block: 128 bytes = 8 slabs of N rows
      const bool acc0 = kblk > 0 || (AH && ps.acc_first);
      if (elect_one()) {
#pragma unroll 4
        for (int st = 0; st < nst; ++st) {
          mma_i8(dcol, ad, bd, idesc, acc0 || st > 0);
          ad += ainc;
          bd += 2 * TC_N;                                         // 32 bytes of K = 2 slabs of N rows
        }
        if (!ps.resident && j + tc.nslot < nblk) mma_commit(tc.bars + TC_MAXSLOT + s);
      }
      __syncwarp();
      if (!ps.resident && ++s == tc.nslot) s = 0;
    }
    if (elect_one()) mma_commit(tc.bars + 2 * TC_MAXSLOT);
    __syncwarp();
    mbar_wait(tc.bars + 2 * TC_MAXSLOT, tc.pd);   // only this warp polls; everybody else parks at the next block barrier
  }
  tc.pd ^= 1u;
}

// Accumulators -> FP64: thread t owns accumulator lane t & 127 of output tile t >> 7 and calls out(t, c, value) for
// t < rows_out.  Call after tc_mma_pass + a block barrier.
template <bool AH, class OutF>
__device__ __forceinline__ void tc_epilogue(const TcCtx& tc, const TcScale& S, int rows_out, OutF out) {
  using namespace umma;
  const int tid = threadIdx.x, warp = tid >> 5;
  tc_fence_after();
  if (32 * warp < rows_out) {
    const int tile = warp >> 2;
    const uint32_t base = tc.tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(tile * 2 * TC_N);
#pragma unroll
    for (int c = 0; c < TC_NC; ++c) {
      int32_t a0[8], a1[8], b0[8], b1[8];   // D_re[x_re], D_re[x_im], D_im[x_re], D_im[x_im] digits
      tmem_ld8(base + (c * 2 + 0) * TC_S, a0);
      tmem_ld8(base + (c * 2 + 1) * TC_S, a1);
      tmem_ld8(base + TC_N + (c * 2 + 0) * TC_S, b0);
      tmem_ld8(base + TC_N + (c * 2 + 1) * TC_S, b1);
      tmem_ld_wait();
      double vr = 0.0, vi = 0.0;
#pragma unroll
      for (int s = TC_S - 1; s >= 0; --s) {
        const int tr = AH ? a0[s] + b1[s] : a0[s] - b1[s];
        const int ti = AH ? a1[s] - b0[s] : a1[s] + b0[s];
        vr = fma(vr, 0.0078125, (double)tr);
        vi = fma(vi, 0.0078125, (double)ti);
      }
      cd v = cmk(vr * S.inv[c], vi * S.inv[c]);
      if (S.bad[c]) v = cmk(NAN, NAN);
      if (tid < rows_out) out(tid, c, v);
    }
  }
  tc_fence_before();
}

// One product over TC_NC operand columns, m <= 256.
//   AH == false:  out(i, c) = sum_k u(i, k)       in(k, c),  k < 256 (rows_in = 256), i < m
//   AH == true :  out(k, c) = sum_i conj(u(i, k)) in(i, c),  i < m   (rows_in = m),   k < 256
// Thread t supplies operand row t and receives output row t.  `red` is shared scratch of >= NW * TC_NC words.
// Contains block barriers; every thread of the CTA must call it.
template <bool AH, class InF, class OutF>
__device__ __forceinline__ void tc_product(TcCtx& tc, const TcGeom& g, int m, InF in, OutF out, uint32_t* red) {
  const int tid = threadIdx.x;
  const int rows_in = AH ? m : 256, rows_out = AH ? 256 : m;
  const TcPass ps = tc_pass_of(g, m);
  // the ring is free (the previous product has completed); in an A' T product WT still holds the operand
  tc_first_round<AH>(tc, ps, AH);
  cd x[TC_NC];
  uint32_t h[TC_NC];
#pragma unroll
  for (int c = 0; c < TC_NC; ++c) {
    x[c] = tid < rows_in ? in(tid, c) : cmk(0.0, 0.0);
    h[c] = tc_hi(x[c]);
  }
  TcScale S;
  tc_scales(h, red, S);
  if (tid < rows_in) tc_slice_row(tc.Bs, tid, x, S);
  umma::fence_async_smem();
  __syncthreads();
  tc_mma_pass<AH>(tc, ps);
  __syncthreads();        // (threads spinning on the mbarrier instead would take issue slots from the issuing warp)
  tc_epilogue<AH>(tc, S, rows_out, out);
  __syncthreads();
}

}  // namespace twoace
