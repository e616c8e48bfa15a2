// Dependent-chain latency of FP64 primitives on this GPU (cycles per op), single warp.
#include <cstdio>
#include <cuda_runtime.h>
#define N 256
template <int OP> __global__ void k(double* out, long long* cyc, double x0) {
  double x = x0 + threadIdx.x * 1e-3;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) {
    if (OP == 0) x = fma(x, 1.0000001, 1e-9);
    if (OP == 1) x = rsqrt(x) + 1.5;
    if (OP == 2) x = sqrt(x) + 1.5;
    if (OP == 3) x = 1.0 / x + 1.5;
    if (OP == 4) x = __drcp_rn(x) + 1.5;
    if (OP == 5) x = (double)rsqrtf((float)x) + 1.5;
    if (OP == 6) x = (double)(1.0f / (float)x) + 1.5;
    if (OP == 7) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31) + 1e-9;
    if (OP == 8) { __syncthreads(); x += 1e-9; }
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  double* d; long long* c; cudaMalloc(&d, 8192); cudaMalloc(&c, 8);
  const char* names[] = {"dfma", "rsqrt(double)", "sqrt(double)", "1.0/x", "__drcp_rn", "rsqrtf+cvt", "rcpf+cvt", "shfl64", "__syncthreads(256thr)"};
  for (int op = 0; op < 9; ++op) {
    long long h = 0;
    for (int rep = 0; rep < 2; ++rep) {
      int nt = (op == 8) ? 256 : 32;
      switch (op) {
        case 0: k<0><<<1, nt>>>(d, c, 1.3); break; case 1: k<1><<<1, nt>>>(d, c, 1.3); break;
        case 2: k<2><<<1, nt>>>(d, c, 1.3); break; case 3: k<3><<<1, nt>>>(d, c, 1.3); break;
        case 4: k<4><<<1, nt>>>(d, c, 1.3); break; case 5: k<5><<<1, nt>>>(d, c, 1.3); break;
        case 6: k<6><<<1, nt>>>(d, c, 1.3); break; case 7: k<7><<<1, nt>>>(d, c, 1.3); break;
        case 8: k<8><<<1, nt>>>(d, c, 1.3); break;
      }
      cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    }
    printf("%-24s %.1f cycles/op (incl. the +1.5 dadd where present)\n", names[op], (double)h / N);
  }
  return 0;
}
