// tcgen05 kind::i8 issue / completion rate for the no-swizzle canonical layouts used by tc_prod.cuh.
#include <cstdio>
#include <cstdlib>
#include "../2ace-mmwave-channel-estimation_b200/csrc/umma_i8.cuh"
using namespace twoace::umma;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void __launch_bounds__(256, 1) rate_kernel(int N, int mode, int nmma, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* A = smem;                 // 16 KB block
  unsigned char* B = smem + 16384;         // N x 256 bytes
  uint64_t* bar = (uint64_t*)(B + 256 * 256);
  uint32_t* tslot = (uint32_t*)(bar + 2);
  const int tid = threadIdx.x;
  for (int i = tid; i < 16384 + 256 * 256; i += 256) smem[i] = (unsigned char)(i * 7 + 1);
  if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  if (tid < 32) tmem_alloc512(tslot);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;
  if (tid == 0) {
    const uint32_t idesc = idesc_i8(N, mode == 1, false);
    const uint32_t sa = smem_u32(A), sb = smem_u32(B);
    uint64_t ad0 = mode == 0 ? smem_desc(sa, SLAB, 128) : smem_desc(sa, 128, SLAB);
    uint64_t bd0 = smem_desc(sb, N * 16, 128);
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
      uint64_t ad = ad0, bd = bd0;
      for (int i = 0; i < nmma; ++i) {
        mma_i8(tmem, ad, bd, idesc, i > 0);
        ad += (mode == 0 ? (2 * SLAB) >> 4 : 32) * ((i & 3) == 3 ? -3 : 1);
        bd += (2 * N) * ((i & 3) == 3 ? -3 : 1);
      }
      const long long t1 = clock64();
      mma_commit(bar);
      mbar_wait(bar, rep & 1);
      const long long t2 = clock64();
      out[rep * 2] = t1 - t0; out[rep * 2 + 1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_free512(tmem);
}

int main() {
  long long* d; CK(cudaMalloc(&d, 64));
  CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  for (int mode = 0; mode < 2; ++mode)
    for (int N : {16, 80, 160, 256})
      for (int nmma : {4, 16, 64}) {
        rate_kernel<<<1, 256, 100 * 1024>>>(N, mode, nmma, d);
        CK(cudaDeviceSynchronize());
        long long h[6]; CK(cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost));
        printf("mode %d N %3d nmma %2d: issue %6lld (%.0f/mma)  issue+done %6lld (%.0f/mma)\n", mode, N, nmma, h[4], (double)h[4] / nmma, h[5], (double)h[5] / nmma);
      }
  return 0;
}
