// jacobi_small<D> vs the software-pipelined jacobi_small_p<D>: cycles per round and bitwise equality.
#include <cstdio>
#include <vector>
#include <complex>
#include <random>
#include "../2ace-mmwave-channel-estimation_b200/csrc/common.cuh"
using namespace twoace;
template <int D, bool PIPE>
__global__ void __launch_bounds__(256) kern(const cd* Gin, cd* Vout, double* evals, long long* cyc, int* sweeps, int reps) {
  extern __shared__ __align__(16) unsigned char raw[];
  cd* G = (cd*)raw; cd* Hh = G + D * D; cd* V = Hh + D * D;
  unsigned char* tab = (unsigned char*)(V + D * D);
  jacobi_tables<D>(tab);
  __syncthreads();
  long long tot = 0; int sw = 0;
  for (int r = 0; r < reps; ++r) {
    for (int e = threadIdx.x; e < D * D; e += 256) G[e] = Gin[blockIdx.x * D * D + e];
    __syncthreads();
    long long t0 = clock64();
    if (PIPE) sw = jacobi_small_p<D>(G, Hh, V, tab, true); else sw = jacobi_small<D>(G, Hh, V, tab, true);
    tot += clock64() - t0;
    __syncthreads();
  }
  for (int e = threadIdx.x; e < D * D; e += 256) Vout[blockIdx.x * D * D + e] = V[e];
  if (threadIdx.x < D) evals[blockIdx.x * D + threadIdx.x] = G[(D + 1) * threadIdx.x].x;
  if (threadIdx.x == 0) { cyc[blockIdx.x] = tot / reps; sweeps[blockIdx.x] = sw; }
}
template <int D> void run() {
  const int nb = 148, reps = 10;
  std::mt19937 rng(1); std::normal_distribution<double> nd;
  std::vector<cd> G(nb * D * D);
  for (int b = 0; b < nb; ++b) {
    std::vector<std::complex<double>> Em(D * 40);
    for (auto& e : Em) e = {nd(rng), nd(rng)};
    for (int i = 0; i < D; ++i) for (int j = 0; j < D; ++j) {
      std::complex<double> s = 0; for (int k = 0; k < 40; ++k) s += Em[i * 40 + k] * std::conj(Em[j * 40 + k]);
      if (i == j) s = s.real();
      G[b * D * D + i + D * j] = make_double2(s.real(), s.imag());
    }
  }
  cd *dG, *dV; double* dE; long long* dC; int* dS;
  cudaMalloc(&dG, nb * D * D * 16); cudaMalloc(&dV, nb * D * D * 16); cudaMalloc(&dE, nb * D * 8); cudaMalloc(&dC, nb * 8); cudaMalloc(&dS, nb * 4);
  cudaMemcpy(dG, G.data(), nb * D * D * 16, cudaMemcpyHostToDevice);
  std::vector<cd> V[2]; std::vector<double> ev[2];
  for (int v = 0; v < 2; ++v) {
    const size_t sm = 3 * D * D * 16 + JacobiTab<D>::BYTES;
    cudaFuncSetAttribute(kern<D, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    cudaFuncSetAttribute(kern<D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (v) kern<D, true><<<nb, 256, sm>>>(dG, dV, dE, dC, dS, reps); else kern<D, false><<<nb, 256, sm>>>(dG, dV, dE, dC, dS, reps);
    cudaDeviceSynchronize();
    V[v].resize(nb * D * D); ev[v].resize(nb * D);
    std::vector<long long> cyc(nb); std::vector<int> sw(nb);
    cudaMemcpy(V[v].data(), dV, nb * D * D * 16, cudaMemcpyDeviceToHost); cudaMemcpy(ev[v].data(), dE, nb * D * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(cyc.data(), dC, nb * 8, cudaMemcpyDeviceToHost); cudaMemcpy(sw.data(), dS, nb * 4, cudaMemcpyDeviceToHost);
    long long c = 0; int s = 0;
    for (int b = 0; b < nb; ++b) { c += cyc[b]; s += sw[b]; }
    printf("D=%d %s: %s, %.0f cycles per call, %.2f sweeps, %.0f cycles/round\n", D, v ? "pipelined" : "baseline ",
           cudaGetErrorString(cudaGetLastError()), (double)c / nb, (double)s / nb, (double)c / s / (D - 1));
  }
  size_t diffV = 0, diffE = 0;
  for (size_t i = 0; i < V[0].size(); ++i) diffV += (V[0][i].x != V[1][i].x) || (V[0][i].y != V[1][i].y);
  for (size_t i = 0; i < ev[0].size(); ++i) diffE += ev[0][i] != ev[1][i];
  printf("D=%d bitwise: %zu differing eigenvector entries, %zu differing eigenvalues\n", D, diffV, diffE);
}
int main() { run<16>(); run<20>(); run<32>(); return 0; }
