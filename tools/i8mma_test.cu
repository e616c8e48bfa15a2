// Stand-alone check of the tcgen05 kind::i8 building blocks in csrc/umma_i8.cuh against a CPU int32 product:
// operand blocks in the no-swizzle canonical layout, read K-major (A X shaped) and MN-major (A' T shaped),
// bulk-copy pipeline on mbarriers, accumulators read back from tensor memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/i8mma_test tools/i8mma_test.cu && tools/i8mma_test
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../2ace-mmwave-channel-estimation_b200/csrc/umma_i8.cuh"

using namespace twoace::umma;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int NSTAGE = 2;

// mode 0: D[i, j] = sum_k A[i, k] B[j, k]   (blocks along k);  mode 1: D[k, j] = sum_i A[i, k] B[j, i]  (blocks along i)
__global__ void __launch_bounds__(256, 1)
i8_kernel(const int8_t* __restrict__ Ablk, int nblk, const int8_t* __restrict__ Bsrc, int N, int mode, int32_t* D,
          long long* cyc, int reps) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* stage = smem;                              // NSTAGE x 16 KB
  unsigned char* Bs = smem + NSTAGE * BLK_BYTES;            // N x K bytes, canonical K-major
  const int K = nblk * BLK;
  uint64_t* bars = (uint64_t*)(Bs + (size_t)N * K);         // full[NSTAGE], empty[NSTAGE], done
  uint32_t* tslot = (uint32_t*)(bars + 2 * NSTAGE + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < 2 * NSTAGE + 1; ++s) mbar_init(bars + s, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc512(tslot);
  // B operand: byte (j, kb) at (kb / 16) * (N * 16) + j * 16 + kb % 16
  for (int idx = tid; idx < N * K; idx += blockDim.x) {
    const int j = idx / K, kb = idx - j * K;
    Bs[(size_t)(kb >> 4) * (N * 16) + j * 16 + (kb & 15)] = (unsigned char)Bsrc[idx];
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;
  const uint32_t idesc = idesc_i8(N, mode == 1, false);
  uint32_t ph_full[NSTAGE] = {0, 0}, ph_empty[NSTAGE] = {0, 0}, ph_done = 0;
  long long t0 = 0;
  for (int rep = 0; rep < reps; ++rep) {
    if (tid == 0) {
      if (rep == 1) t0 = clock64();
      for (int b = 0; b < nblk && b < NSTAGE; ++b) {
        mbar_expect_tx(bars + b, BLK_BYTES);
        bulk_g2s(stage + b * BLK_BYTES, Ablk + (size_t)b * BLK_BYTES, BLK_BYTES, bars + b);
      }
      for (int b = 0; b < nblk; ++b) {
        const int s = b % NSTAGE;
        mbar_wait(bars + s, ph_full[s]);
        ph_full[s] ^= 1;
        tc_fence_after();
        const uint32_t sa = smem_u32(stage + s * BLK_BYTES);
        const uint32_t sb = smem_u32(Bs);
        for (int st = 0; st < 4; ++st) {
          const uint64_t ad = mode == 0 ? smem_desc(sa + st * 2 * SLAB, SLAB, 128) : smem_desc(sa + st * 32 * 16, 128, SLAB);
          const uint64_t bd = smem_desc(sb + ((b * BLK + st * 32) >> 4) * (N * 16), N * 16, 128);
          mma_i8(tmem, ad, bd, idesc, b > 0 || st > 0);
        }
        mma_commit(bars + NSTAGE + s);
        if (b + NSTAGE < nblk) {
          mbar_wait(bars + NSTAGE + s, ph_empty[s]);
          ph_empty[s] ^= 1;
          mbar_expect_tx(bars + s, BLK_BYTES);
          bulk_g2s(stage + s * BLK_BYTES, Ablk + (size_t)(b + NSTAGE) * BLK_BYTES, BLK_BYTES, bars + s);
        } else {
          // drain this stage's commit so the parity bookkeeping stays aligned across reps
          mbar_wait(bars + NSTAGE + s, ph_empty[s]);
          ph_empty[s] ^= 1;
        }
      }
      mma_commit(bars + 2 * NSTAGE);
    }
    mbar_wait(bars + 2 * NSTAGE, ph_done);
    ph_done ^= 1;
    tc_fence_after();
    __syncthreads();
  }
  if (tid == 0) cyc[0] = reps > 1 ? (clock64() - t0) / (reps - 1) : 0;
  if (warp < 4) {
    for (int j0 = 0; j0 < N; j0 += 8) {
      int32_t v[8];
      tmem_ld8(tmem + ((uint32_t)(32 * warp) << 16) + j0, v);
      tmem_ld_wait();
      for (int q = 0; q < 8; ++q) D[(size_t)(32 * warp + lane) * N + j0 + q] = v[q];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_free512(tmem);
}

int main() {
  const int N = 80;
  int fails = 0;
  for (int mode = 0; mode < 2; ++mode) {
    for (int nblk = 1; nblk <= 4; nblk += 1) {
      const int K = nblk * BLK;
      // logical A: mode 0: [128 rows i][K bytes k]; mode 1: [K rows i][128 bytes k]
      std::vector<int8_t> A((size_t)128 * K), B((size_t)N * K), blk((size_t)nblk * BLK_BYTES);
      srand(1234 + mode * 10 + nblk);
      for (auto& a : A) a = (int8_t)((rand() % 3) - 1);
      for (auto& b : B) b = (int8_t)((rand() % 255) - 127);
      for (int b = 0; b < nblk; ++b)
        for (int i = 0; i < 128; ++i)
          for (int k = 0; k < 128; ++k) {
            // block b: mode 0 -> A[i][128 b + k];  mode 1 -> A[128 b + i][k]
            const int8_t v = mode == 0 ? A[(size_t)i * K + 128 * b + k] : A[(size_t)(128 * b + i) * 128 + k];
            blk[(size_t)b * BLK_BYTES + (k >> 4) * SLAB + i * 16 + (k & 15)] = v;
          }
      std::vector<int32_t> ref((size_t)128 * N, 0);
      for (int o = 0; o < 128; ++o)
        for (int j = 0; j < N; ++j) {
          int32_t s = 0;
          for (int q = 0; q < K; ++q) {
            const int a = mode == 0 ? A[(size_t)o * K + q] : A[(size_t)q * 128 + o];
            s += a * (int)B[(size_t)j * K + q];
          }
          ref[(size_t)o * N + j] = s;
        }
      int8_t *dA, *dB; int32_t* dD; long long* dC;
      CK(cudaMalloc(&dA, blk.size())); CK(cudaMalloc(&dB, B.size())); CK(cudaMalloc(&dD, ref.size() * 4)); CK(cudaMalloc(&dC, 8));
      CK(cudaMemcpy(dA, blk.data(), blk.size(), cudaMemcpyHostToDevice));
      CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
      CK(cudaMemset(dD, 0xff, ref.size() * 4));
      const size_t smem = NSTAGE * BLK_BYTES + (size_t)N * K + 256;
      CK(cudaFuncSetAttribute(i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      i8_kernel<<<1, 256, smem>>>(dA, nblk, dB, N, mode, dD, dC, 21);
      CK(cudaGetLastError());
      CK(cudaDeviceSynchronize());
      std::vector<int32_t> out(ref.size());
      long long cyc = 0;
      CK(cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost));
      size_t bad = 0;
      for (size_t i = 0; i < ref.size(); ++i) bad += out[i] != ref[i];
      printf("mode %d nblk %d (K = %d): %zu / %zu mismatches, %lld cycles per product (copy + mma)\n", mode, nblk, K, bad,
             ref.size(), cyc);
      if (bad) {
        ++fails;
        for (int i = 0; i < 4; ++i) printf("  out[%d] = %d ref %d\n", i, out[i], ref[i]);
      }
      cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC);
    }
  }
  printf(fails ? "FAIL\n" : "PASS\n");
  return fails ? 1 : 0;
}
