// On-device instance synthesis: one CTA builds one (trial, M, SNR) instance
//   * Eq. 23 sparse multipath channel   Numerical_Simulation/src/generate_channel/Generate_Channel.m:76-139
//     (L > 1: no Rician tail, :98-106; vecH = vec(H_Matrix), H_Matrix Nr x Nt, :139)
//   * probe selection: M rows without replacement from a row range of the registered codebook
//     (randperm, main/channel_recovery_ADMM_v2_simulation_A2only.m:137; resolution stage of
//     ..._multiresolution.m:137-143 as the row range)
//   * RSS amplitudes |FW vecH + noise| with signal power 1, noise CN(0, 10^(-SNR/10))
//     Numerical_Simulation/src/generate_measurement/Generate_Measurement.m:84-101
//   * the randsample draws of inferLowRankV4.m:37 (one per trial of inferLowRankV4_multi)
// MATLAB's generator streams cannot be reproduced outside MATLAB; the stream here is a counter-based
// Philox4x32-10 (Salmon et al., SC'11) so that an instance depends only on (seed, global trial id) -- independent
// of batch composition, rank count and GPU count -- and is restated bit for bit by oracle/synth.py:
//   key = (seed lo, seed hi), counter = (index, stream, trial lo, trial hi)
//   stream 0, index l: AoD_l, AoA_l = (u0 - 0.5, u1 - 0.5) * searching_area          (degrees)
//   stream 1, index l: path gain g_l = Box-Muller(u0, u1) / sqrt(2)                  (normalised to unit norm)
//   stream 2, index j: 32-bit sort key of candidate row j; the M smallest (key, j) in ascending order are the probes
//   stream 3, index i: noise of measurement i = sqrt(noise_power / 2) * Box-Muller(u0, u1)
//   stream 4 + t, index j: sort key of measurement j for train draw t
//   u0 = ((w0 >> 5) * 2^26 + (w1 >> 6)) * 2^-53, u1 likewise from (w2, w3);
//   Box-Muller(u0, u1) = sqrt(-2 ln(1 - u0)) * (cos 2 pi u1 + j sin 2 pi u1)
#pragma once
#include "common.cuh"

namespace twoace {

struct SynthTask {
  int m;                 // probes
  int row_lo, row_hi;    // candidate codebook rows [row_lo, row_hi)
  double snr_db;
  unsigned long long trial;
  int* rows_out;         // [m] codebook row ids
  int* train_out;        // [ntrain][floor(m * cc_frac)] 0-based measurement ids
  double* B_out;         // [m]
  cd* vecH_out;          // [n]
  double* angles_out;    // [2 L] AoD then AoA (degrees) or nullptr
};

struct SynthDims {
  int nt, nr, L, ntrain;
  double area, k_phase;             // searching area (degrees), 2 pi d / lambda
  double cc_frac, row_scale;
  unsigned int key0, key1;
  int pmax;                         // sort capacity (power of two >= max(row_hi - row_lo, m))
};

__host__ __device__ inline void philox4x32_10(unsigned int c0, unsigned int c1, unsigned int c2, unsigned int c3,
                                              unsigned int k0, unsigned int k1, unsigned int (&out)[4]) {
  for (int r = 0; r < 10; ++r) {
    const unsigned long long p0 = 0xD2511F53ull * c0, p1 = 0xCD9E8D57ull * c2;
    const unsigned int hi0 = (unsigned int)(p0 >> 32), lo0 = (unsigned int)p0;
    const unsigned int hi1 = (unsigned int)(p1 >> 32), lo1 = (unsigned int)p1;
    const unsigned int n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ double u53(unsigned int a, unsigned int b) {
  return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

__device__ __forceinline__ cd box_muller(double u0, double u1) {
  const double rr = sqrt(-2.0 * log(1.0 - u0));
  double s, c;
  sincospi(2.0 * u1, &s, &c);
  return cmk(rr * c, rr * s);
}

// ascending bitonic sort of P (power of two) 64-bit keys in shared memory
__device__ inline void bitonic_sort_u64(unsigned long long* a, int P) {
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < P; i += NT) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned long long x = a[i], y = a[l];
          const bool up = (i & k) == 0;
          if ((x > y) == up) { a[i] = y; a[l] = x; }
        }
      }
      __syncthreads();
    }
  }
}

__host__ __device__ inline size_t synth_smem_bytes(const SynthDims& d) {
  return (size_t)d.pmax * 8 + (size_t)d.nt * d.nr * sizeof(cd) + 4 * 32 * sizeof(double) + 64;
}

__global__ void __launch_bounds__(NT) synth_kernel(const SynthTask* __restrict__ tasks, int ntasks, SynthDims dm,
                                                   const cd* __restrict__ cb_rm, int n) {
  extern __shared__ __align__(16) unsigned char synth_smem[];
  unsigned long long* keys = (unsigned long long*)synth_smem;
  cd* vh = (cd*)(synth_smem + (size_t)dm.pmax * 8);
  double* sc = (double*)(vh + n);          // [0..L) phi_t, [32..) phi_r, [64..) g re, [96..) g im
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int t = blockIdx.x; t < ntasks; t += gridDim.x) {
    const SynthTask tk = tasks[t];
    const unsigned int t0 = (unsigned int)tk.trial, t1 = (unsigned int)(tk.trial >> 32);
    const int L = dm.L, m = tk.m;
    // ---- angles and path gains (Generate_Channel.m:76-106)
    if (tid < L) {
      unsigned int w[4];
      philox4x32_10((unsigned int)tid, 0u, t0, t1, dm.key0, dm.key1, w);
      const double aod = (u53(w[0], w[1]) - 0.5) * dm.area, aoa = (u53(w[2], w[3]) - 0.5) * dm.area;
      sc[tid] = dm.k_phase * sin(aod * (3.14159265358979323846 / 180.0));
      sc[32 + tid] = dm.k_phase * sin(aoa * (3.14159265358979323846 / 180.0));
      philox4x32_10((unsigned int)tid, 1u, t0, t1, dm.key0, dm.key1, w);
      const cd g = box_muller(u53(w[0], w[1]), u53(w[2], w[3]));
      sc[64 + tid] = g.x * 0.70710678118654752440;
      sc[96 + tid] = g.y * 0.70710678118654752440;
      if (tk.angles_out) { tk.angles_out[tid] = aod; tk.angles_out[L + tid] = aoa; }
    }
    __syncthreads();
    if (tid == 0) {
      double s = 0.0;
      for (int l = 0; l < L; ++l) s += sc[64 + l] * sc[64 + l] + sc[96 + l] * sc[96 + l];
      s = 1.0 / sqrt(s);
      for (int l = 0; l < L; ++l) { sc[64 + l] *= s; sc[96 + l] *= s; }
    }
    __syncthreads();
    // ---- vecH[kr + Nr kt] = sum_l g_l exp(-j phi_r,l kr) exp(+j phi_t,l kt)   (:124-139)
    for (int idx = tid; idx < n; idx += NT) {
      const int kr = idx % dm.nr, kt = idx / dm.nr;
      cd acc = cmk(0.0, 0.0);
      for (int l = 0; l < L; ++l) {
        double s, c;
        sincos(sc[l] * kt - sc[32 + l] * kr, &s, &c);
        acc.x += sc[64 + l] * c - sc[96 + l] * s;
        acc.y += sc[64 + l] * s + sc[96 + l] * c;
      }
      vh[idx] = acc;
      tk.vecH_out[idx] = acc;
    }
    // ---- probes: the m smallest (key, j) of the candidate rows
    const int R = tk.row_hi - tk.row_lo;
    int P = 1;
    while (P < R) P <<= 1;
    for (int j = tid; j < P; j += NT) {
      unsigned long long kk = ~0ull;
      if (j < R) {
        unsigned int w[4];
        philox4x32_10((unsigned int)j, 2u, t0, t1, dm.key0, dm.key1, w);
        kk = ((unsigned long long)w[0] << 32) | (unsigned int)j;
      }
      keys[j] = kk;
    }
    __syncthreads();
    bitonic_sort_u64(keys, P);
    for (int i = tid; i < m; i += NT) tk.rows_out[i] = tk.row_lo + (int)(unsigned int)keys[i];
    __syncthreads();
    // ---- B_i = | row_scale * cb[row_i, :] vecH + noise_i |   (Generate_Measurement.m:84-101)
    const double sig = sqrt(pow(10.0, -tk.snr_db / 10.0) * 0.5);
    for (int i = warp; i < m; i += NW) {
      const cd* row = cb_rm + (size_t)(tk.row_lo + (int)(unsigned int)keys[i]) * n;
      cd acc = cmk(0.0, 0.0);
      for (int k = lane; k < n; k += 32) cfma(acc, row[k], vh[k]);
      acc.x = warp_sum(acc.x);
      acc.y = warp_sum(acc.y);
      if (lane == 0) {
        unsigned int w[4];
        philox4x32_10((unsigned int)i, 3u, t0, t1, dm.key0, dm.key1, w);
        const cd z = box_muller(u53(w[0], w[1]), u53(w[2], w[3]));
        const double yr = fma(sig, z.x, dm.row_scale * acc.x), yi = fma(sig, z.y, dm.row_scale * acc.y);
        tk.B_out[i] = sqrt(yr * yr + yi * yi);
      }
    }
    __syncthreads();
    // ---- train draws (inferLowRankV4.m:36-37: randsample(m, floor(m * cc_frac)))
    const int mtr = (int)floor((double)m * dm.cc_frac);
    int P2 = 1;
    while (P2 < m) P2 <<= 1;
    for (int d = 0; d < dm.ntrain; ++d) {
      for (int j = tid; j < P2; j += NT) {
        unsigned long long kk = ~0ull;
        if (j < m) {
          unsigned int w[4];
          philox4x32_10((unsigned int)j, 4u + (unsigned int)d, t0, t1, dm.key0, dm.key1, w);
          kk = ((unsigned long long)w[0] << 32) | (unsigned int)j;
        }
        keys[j] = kk;
      }
      __syncthreads();
      bitonic_sort_u64(keys, P2);
      for (int i = tid; i < mtr; i += NT) tk.train_out[(size_t)d * mtr + i] = (int)(unsigned int)keys[i];
      __syncthreads();
    }
  }
}

}  // namespace twoace
