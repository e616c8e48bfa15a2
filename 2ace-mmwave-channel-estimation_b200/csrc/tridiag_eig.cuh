// Leading eigenpairs of a dense Hermitian matrix: Householder tridiagonalisation, bisection on the Sturm sequence,
// inverse iteration on the tridiagonal matrix, back-transformation.  Replaces MATLAB's eig() at
// inferLowRankV4.m:549 (SpectralInitialize keeps only the r largest eigenpairs, :550-552) for 96 < d <= 512, where a
// full Jacobi decomposition of the d x d Gram matrix was 29 % of a multiresolution (config 4) step:
// ~d^3 complex multiply-adds here against ~40 d^3 for eight sweeps of two-sided block Jacobi.
//   * tridiagonalisation  A = Q T Q', Q = H(0) ... H(d-2), H(k) = I - tau v v'   (the unblocked scheme of LAPACK zhetd2,
//     reflectors as in zlarfg so that the sub-diagonal of T is real)
//   * eigenvalue j of T by multi-section: S points per pass and eigenvalue, counted with the Sturm sequence
//   * eigenvectors of T by inverse iteration with partial pivoting (the scheme of LAPACK dstein / dlagtf), vectors whose
//     eigenvalues lie within 1e-3 ||T|| of each other re-orthogonalised (classical Gram-Schmidt, applied twice)
//   * V = Q Z, one warp per group of columns.
// Validated against numpy.linalg.eigh on the Gram matrices of the benchmark instances (projector on the leading 20
// eigenvectors equal to 4e-15, tests/test_gpu_parity.py::test_spectral_init_*).
#pragma once
#include "common.cuh"

namespace twoace {

constexpr int TRI_DMAX = 512;     // largest d (shared scratch: 64 B per row)
constexpr int TRI_RMAX = 32;      // most eigenpairs

__host__ __device__ inline size_t tri_smem_bytes(int d) { return (size_t)d * 64 + 4 * TRI_RMAX * sizeof(double) + 64 * sizeof(double); }
// global scratch behind the d x r complex output V, in cd units: Z (d r doubles) + three factor arrays
__host__ __device__ inline size_t tri_ws_elems(int d, int r) { return (size_t)d * r + ((size_t)4 * d * r + 1) / 2; }

// G: d x d Hermitian, FULL storage, column-major, leading dimension d, global memory (destroyed).
// V: out, d x r (leading dimension d) followed by tri_ws_elems(d, r) - d r elements of scratch.
// lam: out (shared), r eigenvalues in descending order.  smem: tri_smem_bytes(d) bytes, 16-byte aligned.
// red: block_sum scratch.  Every thread of the CTA calls it.
__device__ inline void top_eig_tridiag(cd* __restrict__ G, int d, int r, cd* __restrict__ V, double* lam, unsigned char* smem,
                                       double* red) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* dd = (double*)smem;            // [d] diagonal of T
  double* ee = dd + d;                   // [d] sub-diagonal of T (ee[d-1] unused)
  cd* tau = (cd*)(ee + d);               // [d]
  cd* vs = tau + d;                      // [d] current Householder vector
  cd* ws = vs + d;                       // [d] p, then w
  double* blo = (double*)(ws + d);       // [TRI_RMAX] bisection brackets
  double* bhi = blo + TRI_RMAX;
  double* sdot = bhi + TRI_RMAX;         // [TRI_RMAX] Gram-Schmidt coefficients
  double* sc = sdot + 2 * TRI_RMAX;      // scalars
  int* cnts = (int*)(sc + 16);           // [256] Sturm counts of a pass (aliases 32 doubles... sized below)
  // (cnts needs NT ints = 1 KB: it aliases vs, which is dead during the bisection)
  cnts = (int*)vs;

  // ---------------------------------------------------------------- tridiagonalisation
  for (int k = 0; k + 1 < d; ++k) {
    const int nk = d - k - 1;                       // order of the trailing block; [alpha; x] has nk entries
    cd* col = G + (size_t)(k + 1) + (size_t)d * k;  // alpha = col[0], x = col[1 .. nk)
    double v1[1] = {0.0};
    for (int i = 1 + tid; i < nk; i += NT) v1[0] += cabs2(col[i]);
    block_sum<1>(v1, red);
    if (tid == 0) {
      const cd alpha = col[0];
      const double xn2 = v1[0];
      if (xn2 == 0.0 && alpha.y == 0.0) {
        tau[k] = cmk(0.0, 0.0);
        ee[k] = alpha.x;
        sc[0] = 0.0; sc[1] = 0.0;
      } else {
        const double nrm = sqrt(cabs2(alpha) + xn2);
        const double beta = alpha.x >= 0.0 ? -nrm : nrm;
        tau[k] = cmk((beta - alpha.x) / beta, -alpha.y / beta);
        ee[k] = beta;
        const double ar = alpha.x - beta, ai = alpha.y, den = ar * ar + ai * ai;     // 1 / (alpha - beta)
        sc[0] = ar / den; sc[1] = -ai / den;
      }
      dd[k] = G[(size_t)k + (size_t)d * k].x;
    }
    __syncthreads();
    const cd tk = tau[k];
    const cd scl = cmk(sc[0], sc[1]);
    for (int i = tid; i < nk; i += NT) {
      const cd v = (i == 0) ? cmk(1.0, 0.0) : cmul(col[i], scl);
      vs[i] = v;
      col[i] = v;                                   // kept for the back-transformation
    }
    __syncthreads();
    if (tk.x != 0.0 || tk.y != 0.0) {
      cd* A22 = G + (size_t)(k + 1) + (size_t)d * (k + 1);
      // p = tau A22 v   (rows over threads: consecutive threads read consecutive rows of a column)
      for (int i = tid; i < nk; i += NT) {
        cd a0 = cmk(0.0, 0.0), a1 = a0, a2 = a0, a3 = a0;
        const cd* rowp = A22 + i;
        int j = 0;
        for (; j + 3 < nk; j += 4) {
          cfma(a0, rowp[(size_t)d * j], vs[j]);
          cfma(a1, rowp[(size_t)d * (j + 1)], vs[j + 1]);
          cfma(a2, rowp[(size_t)d * (j + 2)], vs[j + 2]);
          cfma(a3, rowp[(size_t)d * (j + 3)], vs[j + 3]);
        }
        for (; j < nk; ++j) cfma(a0, rowp[(size_t)d * j], vs[j]);
        const cd acc = cmk((a0.x + a1.x) + (a2.x + a3.x), (a0.y + a1.y) + (a2.y + a3.y));
        ws[i] = cmul(tk, acc);
      }
      __syncthreads();
      // alpha2 = -1/2 tau (p' v);  w = p + alpha2 v
      double v2[2] = {0.0, 0.0};
      for (int i = tid; i < nk; i += NT) {
        const cd p = ws[i], v = vs[i];
        v2[0] += p.x * v.x + p.y * v.y;             // conj(p) v
        v2[1] += p.x * v.y - p.y * v.x;
      }
      block_sum<2>(v2, red);
      const cd al = cmul(cmk(-0.5 * tk.x, -0.5 * tk.y), cmk(v2[0], v2[1]));
      for (int i = tid; i < nk; i += NT) {
        const cd v = vs[i];
        cd w = ws[i];
        w.x += al.x * v.x - al.y * v.y;
        w.y += al.x * v.y + al.y * v.x;
        ws[i] = w;
      }
      __syncthreads();
      // A22 <- A22 - v w' - w v'
      for (int i = tid; i < nk; i += NT) {
        const cd vi = vs[i], wi = ws[i];
        cd* rowp = A22 + i;
#pragma unroll 4
        for (int j = 0; j < nk; ++j) {
          const cd vj = vs[j], wj = ws[j];
          cd a = rowp[(size_t)d * j];
          // v_i conj(w_j) + w_i conj(v_j)
          a.x -= (vi.x * wj.x + vi.y * wj.y) + (wi.x * vj.x + wi.y * vj.y);
          a.y -= (vi.y * wj.x - vi.x * wj.y) + (wi.y * vj.x - wi.x * vj.y);
          rowp[(size_t)d * j] = a;
        }
      }
    }
    __syncthreads();
  }
  if (tid == 0) { dd[d - 1] = G[(size_t)(d - 1) + (size_t)d * (d - 1)].x; ee[d - 1] = 0.0; }
  __syncthreads();

  // ---------------------------------------------------------------- eigenvalues: multi-section on the Sturm count
  {
    double v3[3] = {0.0, 0.0, 0.0};     // max |d|, max |e|, (unused)
    double mx_d = 0.0, mx_e = 0.0, glo = INFINITY, ghi = -INFINITY;
    for (int i = tid; i < d; i += NT) {
      const double el = i > 0 ? fabs(ee[i - 1]) : 0.0, er = i + 1 < d ? fabs(ee[i]) : 0.0;
      mx_d = fmax(mx_d, fabs(dd[i]));
      mx_e = fmax(mx_e, er);
      glo = fmin(glo, dd[i] - el - er);
      ghi = fmax(ghi, dd[i] + el + er);
    }
    // block max / min through shared memory (no atomics on doubles): use red as [4][NW]
    for (int o = 16; o > 0; o >>= 1) {
      mx_d = fmax(mx_d, __shfl_xor_sync(0xffffffffu, mx_d, o));
      mx_e = fmax(mx_e, __shfl_xor_sync(0xffffffffu, mx_e, o));
      glo = fmin(glo, __shfl_xor_sync(0xffffffffu, glo, o));
      ghi = fmax(ghi, __shfl_xor_sync(0xffffffffu, ghi, o));
    }
    if (lane == 0) { red[warp] = mx_d; red[NW + warp] = mx_e; red[2 * NW + warp] = glo; red[3 * NW + warp] = ghi; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < NW; ++w) {
        red[0] = fmax(red[0], red[w]); red[NW] = fmax(red[NW], red[NW + w]);
        red[2 * NW] = fmin(red[2 * NW], red[2 * NW + w]); red[3 * NW] = fmax(red[3 * NW], red[3 * NW + w]);
      }
      const double tn = fmax(red[0] + 2.0 * red[NW], 1e-300);
      const double pivmin = 2.2250738585072014e-308 * fmax(1.0, red[NW] * red[NW]);
      const double pad = 2.0 * tn * 1e-15 * d + 2.0 * pivmin;
      sc[2] = tn; sc[3] = pivmin; sc[4] = red[2 * NW] - pad; sc[5] = red[3 * NW] + pad;
    }
    __syncthreads();
    (void)v3;
  }
  const double tn = sc[2], pivmin = sc[3];
  {
    const int S = NT / r;                       // section points per eigenvalue and pass (8 for r = 32, 12 for r = 20)
    int passes = 1;
    for (double w = 1.0; w < 4e16; w *= (double)(S + 1)) ++passes;
    const int g = tid / S, s = tid - g * S;     // eigenvalue (descending index) and point of this thread
    const bool act = g < r;
    if (tid < r) { blo[tid] = sc[4]; bhi[tid] = sc[5]; }
    __syncthreads();
    for (int ps = 0; ps < passes; ++ps) {
      int cnt = 0;
      double x = 0.0;
      if (act) {
        const double lo = blo[g], hi = bhi[g];
        x = lo + (hi - lo) * (double)(s + 1) / (double)(S + 1);
        double q = dd[0] - x;
        if (fabs(q) < pivmin) q = -pivmin;
        cnt = q < 0.0;
        for (int i = 1; i < d; ++i) {
          const double e = ee[i - 1];
          q = dd[i] - x - e * e / q;
          if (fabs(q) < pivmin) q = -pivmin;
          cnt += q < 0.0;
        }
      }
      cnts[tid] = cnt;
      __syncthreads();
      if (act && s == 0) {
        const int j = d - 1 - g;                // ascending index of the wanted eigenvalue: count(x) <= j below it
        const double lo = blo[g], hi = bhi[g];
        double nlo = lo, nhi = hi;
        for (int q = 0; q < S; ++q) {           // counts are monotone in the point index
          const double xq = lo + (hi - lo) * (double)(q + 1) / (double)(S + 1);
          if (cnts[g * S + q] <= j) nlo = xq; else { nhi = xq; break; }
        }
        blo[g] = nlo; bhi[g] = nhi;
      }
      __syncthreads();
    }
    if (tid < r) lam[tid] = 0.5 * (blo[tid] + bhi[tid]);
    __syncthreads();
  }

  // ---------------------------------------------------------------- eigenvectors of T: inverse iteration
  double* Z = (double*)(V + (size_t)d * r);     // [d x r], column c contiguous
  double* Fa = Z + (size_t)d * r;               // pivots
  double* Fu = Fa + (size_t)d * r;              // first super-diagonal of U
  double* Fu2 = Fu + (size_t)d * r;             // second super-diagonal of U
  for (size_t idx = tid; idx < (size_t)d * r; idx += NT) {
    const unsigned int i = (unsigned int)(idx % d), c = (unsigned int)(idx / d);
    Z[idx] = (double)((i * 2654435761u + 40503u * c) % 1000003u) / 1000003.0 - 0.5;   // deterministic start vectors
  }
  __syncthreads();
  const double ortol = 1e-3 * tn, tiny = 2.220446049250313e-16 * tn;
  for (int it = 0; it < 3; ++it) {
    if (tid < r) {          // thread c: factor T - lam_c I with partial pivoting and solve for column c in place
      const int c = tid;
      const double lm = lam[c];
      double* x = Z + (size_t)d * c;
      double* fa = Fa + (size_t)d * c; double* fu = Fu + (size_t)d * c; double* fu2 = Fu2 + (size_t)d * c;
      double a_i = dd[0] - lm, du_i = ee[0], x_i = x[0];
      for (int i = 0; i + 1 < d; ++i) {
        const double dl = ee[i];
        double a_n = dd[i + 1] - lm, du_n = (i + 2 < d) ? ee[i + 1] : 0.0, x_n = x[i + 1];
        if (fabs(a_i) >= fabs(dl)) {
          if (fabs(a_i) < tiny) a_i = tiny;
          const double f = dl / a_i;
          fa[i] = a_i; fu[i] = du_i; fu2[i] = 0.0; x[i] = x_i;
          a_n -= f * du_i;
          x_n -= f * x_i;
        } else {            // swap rows i and i + 1
          const double f = a_i / dl;
          fa[i] = dl; fu[i] = a_n; fu2[i] = du_n; x[i] = x_n;
          const double t = a_n;
          a_n = du_i - f * t;
          du_n = -f * du_n;
          x_n = x_i - f * x_n;
        }
        a_i = a_n; du_i = du_n; x_i = x_n;
      }
      if (fabs(a_i) < tiny) a_i = tiny;
      double x1 = x_i / a_i, x2 = 0.0;          // x[d-1]
      x[d - 1] = x1;
      for (int i = d - 2; i >= 0; --i) {
        const double xi = (x[i] - fu[i] * x1 - fu2[i] * x2) / fa[i];
        x[i] = xi;
        x2 = x1; x1 = xi;
      }
    }
    __syncthreads();
    // Gram-Schmidt inside clusters of close eigenvalues (twice), then normalisation; columns in order
    for (int c = 0; c < r; ++c) {
      double* zc = Z + (size_t)d * c;
      const double lc = lam[c];
      for (int rep = 0; rep < 2; ++rep) {
        bool any = false;
        for (int i = 0; i < c; ++i) any = any || (fabs(lam[i] - lc) <= ortol);
        if (!any) break;                        // uniform: depends on shared data only
        for (int i = warp; i < c; i += NW) {
          if (fabs(lam[i] - lc) <= ortol) {
            const double* zi = Z + (size_t)d * i;
            double sdt = 0.0;
            for (int q = lane; q < d; q += 32) sdt += zi[q] * zc[q];
            sdt = warp_sum(sdt);
            if (lane == 0) sdot[i] = sdt;
          }
        }
        __syncthreads();
        for (int q = tid; q < d; q += NT) {
          double z = zc[q];
          for (int i = 0; i < c; ++i)
            if (fabs(lam[i] - lc) <= ortol) z -= sdot[i] * Z[(size_t)d * i + q];
          zc[q] = z;
        }
        __syncthreads();
      }
      double v1[1] = {0.0};
      for (int q = tid; q < d; q += NT) v1[0] += zc[q] * zc[q];
      block_sum<1>(v1, red);
      const double inv = v1[0] > 0.0 ? 1.0 / sqrt(v1[0]) : 0.0;
      for (int q = tid; q < d; q += NT) zc[q] *= inv;
      __syncthreads();
    }
  }

  // ---------------------------------------------------------------- V = Q Z
  for (size_t idx = tid; idx < (size_t)d * r; idx += NT) V[idx] = cmk(Z[idx], 0.0);
  __syncthreads();
  {
    constexpr int CW = (TRI_RMAX + NW - 1) / NW;     // columns per warp: warp, warp + NW, ...
    for (int k = d - 2; k >= 0; --k) {
      const cd tk = tau[k];
      if (tk.x == 0.0 && tk.y == 0.0) continue;
      const int nk = d - k - 1;
      const cd* v = G + (size_t)(k + 1) + (size_t)d * k;
      cd sdt[CW];
#pragma unroll
      for (int u = 0; u < CW; ++u) sdt[u] = cmk(0.0, 0.0);
      for (int i = lane; i < nk; i += 32) {
        const cd vi = v[i];
#pragma unroll
        for (int u = 0; u < CW; ++u) {
          const int c = warp + NW * u;
          if (c < r) cfmac(sdt[u], vi, V[(size_t)d * c + k + 1 + i]);       // conj(v_i) V(i, c)
        }
      }
#pragma unroll
      for (int u = 0; u < CW; ++u) {
        sdt[u].x = warp_sum(sdt[u].x);
        sdt[u].y = warp_sum(sdt[u].y);
        sdt[u] = cmul(tk, sdt[u]);
      }
      for (int i = lane; i < nk; i += 32) {
        const cd vi = v[i];
#pragma unroll
        for (int u = 0; u < CW; ++u) {
          const int c = warp + NW * u;
          if (c < r) {
            cd x = V[(size_t)d * c + k + 1 + i];
            x.x -= vi.x * sdt[u].x - vi.y * sdt[u].y;
            x.y -= vi.x * sdt[u].y + vi.y * sdt[u].x;
            V[(size_t)d * c + k + 1 + i] = x;
          }
        }
      }
      __syncwarp();
    }
  }
  __syncthreads();
}

}  // namespace twoace
