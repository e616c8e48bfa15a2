// L2 -> shared-memory streaming rate of one CTA per SM: cp.async.bulk with D copies of `bytes` in flight, and a plain
// LDG.128 loop for comparison.  Prints bytes/cycle/SM.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 ...
#include <cstdio>
#include <cstdlib>
#include "../2ace-mmwave-channel-estimation_b200/csrc/umma_i8.cuh"
using namespace twoace::umma;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void __launch_bounds__(256, 1) bulk_kernel(const unsigned char* src, size_t per_cta, int bytes, int depth, int total, long long* cyc, int nissue) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars0 = (uint64_t*)smem;            // [64]
  unsigned char* buf = smem + 1024;
  const unsigned char* s = src + (size_t)blockIdx.x * per_cta;
  if (threadIdx.x == 0) { for (int i = 0; i < 64; ++i) mbar_init(bars0 + i, 1); mbar_fence_init(); }
  __syncthreads();
  if ((threadIdx.x & 31) == 0 && (threadIdx.x >> 5) < nissue) {
    const int w = threadIdx.x >> 5;
    uint64_t* bars = bars0 + 8 * w;
    buf += (size_t)w * depth * bytes;
    s += (size_t)w * (per_cta / nissue);
    total /= nissue;
    uint32_t ph = 0;
    const long long t0 = clock64();
    for (int j = 0; j < depth && j < total; ++j) { mbar_expect_tx(bars + j, bytes); bulk_g2s(buf + (size_t)j * bytes, s + ((size_t)j * bytes) % per_cta, bytes, bars + j); }
    for (int j = 0; j < total; ++j) {
      const int sl = j % depth;
      mbar_wait(bars + sl, (ph >> sl) & 1u); ph ^= 1u << sl;
      if (j + depth < total) { mbar_expect_tx(bars + sl, bytes); bulk_g2s(buf + (size_t)sl * bytes, s + ((size_t)(j + depth) * bytes) % per_cta, bytes, bars + sl); }
    }
    if (w == 0) cyc[blockIdx.x] = clock64() - t0;
  }
}

__global__ void __launch_bounds__(256, 1) ldg_kernel(const double2* src, size_t per_cta_elems, int total_elems, double2* out, long long* cyc) {
  const double2* s = src + (size_t)blockIdx.x * per_cta_elems;
  double2 acc = make_double2(0, 0);
  const long long t0 = clock64();
#pragma unroll 8
  for (int i = threadIdx.x; i < total_elems; i += 256) { const double2 v = __ldg(s + (i & (per_cta_elems - 1))); acc.x += v.x; acc.y += v.y; }
  const long long t1 = clock64();
  out[blockIdx.x * 256 + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  const size_t per_cta = 1 << 20;   // 1 MB per CTA (L2 resident: 148 MB > L2, so use 64 CTAs for the L2 test)
  for (int grid : {4}) {
    unsigned char* d; long long* dc; double2* dout;
    CK(cudaMalloc(&d, per_cta * grid)); CK(cudaMemset(d, 1, per_cta * grid)); CK(cudaMalloc(&dc, 8 * grid)); CK(cudaMalloc(&dout, grid * 256 * 16));
    CK(cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int nissue : {1, 2, 4}) for (int bytes : {8192, 16384}) for (int depth : {2}) {
      if ((size_t)bytes * depth * nissue > 190 * 1024) continue;
      const int total = (int)(4 * per_cta / bytes);
      for (int rep = 0; rep < 2; ++rep) bulk_kernel<<<grid, 256, 200 * 1024>>>(d, per_cta, bytes, depth, total, dc, nissue);
      CK(cudaDeviceSynchronize());
      long long c[148]; CK(cudaMemcpy(c, dc, 8 * grid, cudaMemcpyDeviceToHost));
      long long mx = 0; for (int i = 0; i < grid; ++i) mx = c[i] > mx ? c[i] : mx;
      printf("grid %3d issuers %d bulk %5d B x depth %d: %.1f B/clk/SM (%.0f cycles per copy)\n", grid, nissue, bytes, depth, (double)total * bytes / mx, (double)mx / total);
    }
    const int te = (int)(4 * per_cta / 16);
    for (int rep = 0; rep < 2; ++rep) ldg_kernel<<<grid, 256>>>((const double2*)d, per_cta / 16, te, dout, dc);
    CK(cudaDeviceSynchronize());
    long long c[148]; CK(cudaMemcpy(c, dc, 8 * grid, cudaMemcpyDeviceToHost));
    long long mx = 0; for (int i = 0; i < grid; ++i) mx = c[i] > mx ? c[i] : mx;
    printf("grid %3d LDG.128 unroll 8: %.1f B/clk/SM\n", grid, (double)te * 16 / mx);
    cudaFree(d); cudaFree(dc); cudaFree(dout);
  }
  return 0;
}
