// Isolated timing + accuracy of jacobi16 (cycles per sweep) on this GPU.
#include <cstdio>
#include <vector>
#include <complex>
#include <random>
#include "../2ace-mmwave-channel-estimation_b200/csrc/common.cuh"
using namespace twoace;
__global__ void __launch_bounds__(256) kern(const cd* Gin, cd* Vout, double* evals, long long* cyc, int* sweeps, int reps) {
  __shared__ cd G[256], H[256], V[256];
  __shared__ unsigned char pairs[1280];
  jacobi_tables<16>(pairs);
  __syncthreads();
  long long tot = 0; int sw = 0;
  for (int r = 0; r < reps; ++r) {
    G[threadIdx.x] = Gin[blockIdx.x * 256 + threadIdx.x];
    __syncthreads();
    long long t0 = clock64();
    sw = jacobi_small<16>(G, H, V, pairs, true);
    tot += clock64() - t0;
    __syncthreads();
  }
  Vout[blockIdx.x * 256 + threadIdx.x] = V[threadIdx.x];
  if (threadIdx.x < 16) evals[blockIdx.x * 16 + threadIdx.x] = G[17 * threadIdx.x].x;
  if (threadIdx.x == 0) { cyc[blockIdx.x] = tot / reps; sweeps[blockIdx.x] = sw; }
}
int main() {
  const int nb = 148, reps = 20;
  std::mt19937 rng(1); std::normal_distribution<double> nd;
  std::vector<cd> G(nb * 256);
  for (int b = 0; b < nb; ++b) {
    std::complex<double> E[16][40];
    for (auto& row : E) for (auto& e : row) e = {nd(rng), nd(rng)};
    for (int i = 0; i < 16; ++i) for (int j = 0; j < 16; ++j) {
      std::complex<double> s = 0; for (int k = 0; k < 40; ++k) s += E[i][k] * std::conj(E[j][k]);
      if (i == j) s = s.real();
      G[b * 256 + i + 16 * j] = make_double2(s.real(), s.imag());
    }
  }
  cd *dG, *dV; double* dE; long long* dC; int* dS;
  cudaMalloc(&dG, nb * 256 * 16); cudaMalloc(&dV, nb * 256 * 16); cudaMalloc(&dE, nb * 16 * 8); cudaMalloc(&dC, nb * 8); cudaMalloc(&dS, nb * 4);
  cudaMemcpy(dG, G.data(), nb * 256 * 16, cudaMemcpyHostToDevice);
  kern<<<nb, 256>>>(dG, dV, dE, dC, dS, reps);
  std::vector<cd> V(nb * 256); std::vector<double> ev(nb * 16); std::vector<long long> cyc(nb); std::vector<int> sw(nb);
  cudaMemcpy(V.data(), dV, nb * 256 * 16, cudaMemcpyDeviceToHost); cudaMemcpy(ev.data(), dE, nb * 16 * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(cyc.data(), dC, nb * 8, cudaMemcpyDeviceToHost); cudaMemcpy(sw.data(), dS, nb * 4, cudaMemcpyDeviceToHost);
  printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
  double maxres = 0, maxorth = 0; long long c = 0; int s = 0;
  for (int b = 0; b < nb; ++b) {
    c += cyc[b]; s += sw[b];
    for (int i = 0; i < 16; ++i) for (int j = 0; j < 16; ++j) {   // residual G V - V diag(ev), orthogonality
      std::complex<double> r = 0, o = 0;
      for (int k = 0; k < 16; ++k) {
        std::complex<double> g(G[b * 256 + i + 16 * k].x, G[b * 256 + i + 16 * k].y), v(V[b * 256 + k + 16 * j].x, V[b * 256 + k + 16 * j].y);
        r += g * v;
        std::complex<double> vi(V[b * 256 + k + 16 * i].x, V[b * 256 + k + 16 * i].y);
        o += std::conj(vi) * v;
      }
      std::complex<double> vij(V[b * 256 + i + 16 * j].x, V[b * 256 + i + 16 * j].y);
      maxres = fmax(maxres, std::abs(r - vij * ev[b * 16 + j]));
      maxorth = fmax(maxorth, std::abs(o - (i == j ? 1.0 : 0.0)));
    }
  }
  printf("jacobi16 cold: %.0f cycles per call, %.2f sweeps -> %.0f cycles/sweep, %.0f cycles/round; max |GV - V L| = %.2e (|G|~80), max |V'V - I| = %.2e\n",
         (double)c / nb, (double)s / nb, (double)c / s, (double)c / s / 15, maxres, maxorth);
  return 0;
}
