// PhaseLift: trace-regularised least squares over PSD matrices solved with TFOCS'
// Auslender-Teboulle accelerated proximal gradient, one CTA per instance, whole solve in one launch.
//
//   MyPhaseLift.m:69-107            -> run_phaselift (options, x0 = 0, leading eigenvector of the result)
//   initializeLinopPR.m:61,65       -> lifted_forward / the adjoint GEMM in prox_step
//   tfocs_AT.m:20-94, tfocs_backtrack.m:4-46, tfocs_iterate.m:8-33,326-362 -> the loop below, statement by statement
//   prox_trace.m:69-158             -> warm-started block-Jacobi eigendecomposition + positive part of (D - lambda t)
//
// Formulation differences (exact in exact arithmetic, parity-tested against oracle/phaselift.py):
//  * Row-space reduction.  With x0 = 0 every iterate lies in range(A') (x) range(A'): the gradient
//    A' diag(v) A does, and the prox keeps eigenvectors of eigenvalues > lambda*t > 0 only.  For m < n we factor
//    A A' = L L' (Cholesky), so that A = L Qh with orthonormal rows Qh, and iterate on the m x m matrix
//    Xr = Qh X Qh' with the operator L in place of A (all Frobenius norms and inner products are preserved).
//    The eigenproblem per iteration shrinks from n x n to m x m.  Rank-deficient A A' or m >= n: no reduction.
//  * The lifted measurements of a Hermitian matrix are real; the reference carries rounding-level imaginary
//    parts through (they enter f at the 1e-34 level).  We keep A(x) as real vectors.
//  * x - y = theta (z - z_old) (tfocs_AT.m:38,69), so |x - y|^2 and <x - y, g_y> = sum_i g_i A(x - y)_i are
//    evaluated from z, z_old and the measurement vectors without forming y or storing g_y.
//  * A(z) = sum_k s_k |A v_k|^2 over the kept eigenpairs instead of diag(A z A').
//  * eig(): two-sided block Jacobi warm-started from the previous eigenbasis U (G = U' W U), stopped when the
//    largest rotation of a sweep has |sin| <= 1e-6 (off-diagonals then <= 1e-12 |W|: the prox is non-expansive,
//    errors do not accumulate beyond that level).  Only eigenpairs above lambda*t enter the prox, so after the
//    first (full) sweep the later sweeps are restricted to block pairs holding a diagonal entry above that
//    threshold; the kept eigenvectors are moved to the leading columns so that this is one block row.
#pragma once
#include "common.cuh"

namespace twoace {

struct PlOpts {
  int maxIts;
  double tol;
  int restart;
  double lam, alpha, beta, L0;
  int cntr_reset;
  double backtrack_tol;
  int reduce;      // 1: use the row-space reduction when m < n
};

struct PlTask {
  const cd* A_cm;      // dense m x n column-major, or nullptr
  const cd* cb;        // row-major codebook (rows x n), used when A_cm == nullptr
  const int* rows;     // codebook row ids [m]
  double scale;        // A(i, k) = scale * cb[rows[i], k]
  const double* y;     // intensities [m]
  int m;
  cd* sig;             // out [n]
  double* info;        // out [PL_INFO] or nullptr
};

constexpr int PL_INFO = 16;  // niter, n_prox, n_backtracks, status, rank, L, d, lambda_max, Jacobi sweeps,
                             // cycles: gradient GEMM, warm transform, Jacobi, z / A_z, x update + tests, subproblem solves
                             // (part of Jacobi); 15 reserved
enum { PL_ST_TOL = 1, PL_ST_MAXIT = 2, PL_ST_DX0 = 3, PL_ST_NAN = 4, PL_ST_SMALLSTEP = 5 };

constexpr int PG_TM = 64, PG_TK = 8, PG_LD = PG_TM + 1;
constexpr int PG_TILE = PG_TK * PG_LD;   // cd elements per operand tile

// C(i, j) = sum_k a(i, k) * b(k, j), i < M, j < N.  256 threads, 64 x 64 output tile, 4 x 4 per thread, operand
// chunks of 8 staged through shared memory with a register prefetch of the next chunk.  AKF / BKF: the operand
// functor is contiguous in k (else in i / j) -- decides which index runs fastest over the loading threads.
// HERM: the result is Hermitian (M == N): tiles strictly above the diagonal are skipped and the caller's store
// functor mirrors the tiles below it (store is then called with i0 >= j0 tiles only).
template <bool AKF, bool BKF, bool HERM = false, class FA, class FB, class FS>
__device__ __forceinline__ void cta_gemm(int M, int N, int K, FA a, FB b, FS store, cd* sA, cd* sB) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  for (int j0 = 0; j0 < N; j0 += PG_TM) {
    for (int i0 = 0; i0 < M; i0 += PG_TM) {
      if (HERM && i0 < j0) continue;
      cd acc[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = cmk(0.0, 0.0);
      cd ra[2], rb[2];
      auto fetch = [&](int k0) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int idx = tid + u * NT;
          const int ka = AKF ? (idx & 7) : (idx >> 6), ia = AKF ? (idx >> 3) : (idx & 63);
          const int kb = BKF ? (idx & 7) : (idx >> 6), jb = BKF ? (idx >> 3) : (idx & 63);
          ra[u] = (i0 + ia < M && k0 + ka < K) ? a(i0 + ia, k0 + ka) : cmk(0.0, 0.0);
          rb[u] = (j0 + jb < N && k0 + kb < K) ? b(k0 + kb, j0 + jb) : cmk(0.0, 0.0);
        }
      };
      fetch(0);
      for (int k0 = 0; k0 < K; k0 += PG_TK) {
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int idx = tid + u * NT;
          const int ka = AKF ? (idx & 7) : (idx >> 6), ia = AKF ? (idx >> 3) : (idx & 63);
          const int kb = BKF ? (idx & 7) : (idx >> 6), jb = BKF ? (idx >> 3) : (idx & 63);
          sA[ka * PG_LD + ia] = ra[u];
          sB[kb * PG_LD + jb] = rb[u];
        }
        __syncthreads();
        if (k0 + PG_TK < K) fetch(k0 + PG_TK);
#pragma unroll
        for (int kk = 0; kk < PG_TK; ++kk) {
          cd av[4], bv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) av[u] = sA[kk * PG_LD + tx + 16 * u];
#pragma unroll
          for (int v = 0; v < 4; ++v) bv[v] = sB[kk * PG_LD + ty + 16 * v];
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int v = 0; v < 4; ++v) cfma(acc[u][v], av[u], bv[v]);
        }
      }
#pragma unroll
      for (int v = 0; v < 4; ++v)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + tx + 16 * u, j = j0 + ty + 16 * v;
          if (i < M && j < N) store(i, j, acc[u][v]);
        }
    }
  }
  __syncthreads();
}

struct PlSmem {
  cd *sA, *sB;             // GEMM operand tiles (aliases of Sb / Q)
  cd *S, *Sb, *Q;          // 32 x 32 Jacobi subproblem buffers
  unsigned char* tab;      // JacobiTab<32>
  double *b, *Ax, *Az, *Ay, *Axo, *Azo, *g;   // measurement-space vectors [maxm]
  double* sv;              // kept shifted eigenvalues [n]
  double* sacc;            // [n] diagonal of U' z_old U (shifted eigenvalues of the accepted prox, by column)
  int* idx;                // their column indices [n]
  double* red;             // reduction scratch [4 * NW]
  int* ib;                 // small int scratch [16]
};

__host__ __device__ inline size_t pl_smem_bytes(int n, int maxm) {
  size_t b = 0;
  b += 3 * 32 * 32 * sizeof(cd);
  b += 2 * JacobiTab<32>::BYTES + 64;
  b += 7 * (size_t)((maxm + 1) / 2 * 2) * sizeof(double);
  b += 2 * (size_t)n * sizeof(double) + (size_t)n * sizeof(int);
  b += 4 * NW * sizeof(double) + 16 * sizeof(int);
  return b + 64;
}

__device__ inline PlSmem pl_carve(unsigned char* p, int n, int maxm) {
  PlSmem s;
  const int mv = (maxm + 1) / 2 * 2;
  s.S = reinterpret_cast<cd*>(p);  p += 32 * 32 * sizeof(cd);
  s.Sb = reinterpret_cast<cd*>(p); p += 32 * 32 * sizeof(cd);
  s.Q = reinterpret_cast<cd*>(p);  p += 32 * 32 * sizeof(cd);
  s.sA = s.Sb;   // the GEMM tiles (520 elements each) alias two Jacobi buffers: never live at the same time
  s.sB = s.Q;
  s.tab = p;                       p += 2 * JacobiTab<32>::BYTES + 64;
  double* dp = reinterpret_cast<double*>(p);
  s.b = dp; s.Ax = dp + mv; s.Az = dp + 2 * mv; s.Ay = dp + 3 * mv; s.Axo = dp + 4 * mv; s.Azo = dp + 5 * mv;
  s.g = dp + 6 * mv;
  dp += 7 * mv;
  s.sv = dp; dp += n;
  s.sacc = dp; dp += n;
  s.red = dp; dp += 4 * NW;
  s.idx = reinterpret_cast<int*>(dp);
  s.ib = s.idx + n;
  return s;
}

// per-CTA global workspace (cd elements): At [maxm x dcap], X[2], Z[2], W, U[2] (dcap x dcap each),
// T [max(maxm,dcap) x dcap].  dcap = the largest dimension the iteration can run in: n for n <= PL_DMAX; for wider
// problems the iteration only runs in the row space of A (m <= PL_DMAX rows) and dcap = maxm.
constexpr int PL_DMAX = 256;     // block Jacobi: 16 blocks of 16
constexpr int PL_NMAX = 2000;    // MyPhaseLift.m:87-89 switches to opts.largescale above this
__host__ __device__ inline int pl_dcap(int n, int maxm) { return n <= PL_DMAX ? n : (maxm < PL_DMAX ? maxm : PL_DMAX); }
__host__ __device__ inline size_t pl_ws_elems(int dcap, int maxm) {
  const size_t nn = (size_t)dcap * dcap;
  return (size_t)maxm * dcap + 7 * nn + (size_t)(maxm > dcap ? maxm : dcap) * dcap;
}

struct PlWs {
  cd *At, *X[2], *Z[2], *W, *U[2], *T;
};

__device__ inline PlWs pl_ws(cd* base, int dcap, int maxm) {
  PlWs w;
  const size_t nn = (size_t)dcap * dcap;
  w.At = base; base += (size_t)maxm * dcap;
  w.X[0] = base; base += nn;
  w.X[1] = base; base += nn;
  w.Z[0] = base; base += nn;
  w.Z[1] = base; base += nn;
  w.W = base; base += nn;
  w.U[0] = base; base += nn;
  w.U[1] = base; base += nn;
  w.T = base;
  return w;
}

// out[i] = Re( a_i M a_i' ), i < m, for a Hermitian d x d matrix M (initializeLinopPR.m:61); T: m x d scratch.
__device__ inline void lifted_forward(const cd* At, int m, int d, const cd* M, cd* T, double* out, const PlSmem& sm) {
  cta_gemm<false, true>(m, d, d, [&](int i, int k) { return At[i + (size_t)m * k]; },
                        [&](int k, int j) { return M[k + (size_t)d * j]; },
                        [&](int i, int j, cd v) { T[i + (size_t)m * j] = v; }, sm.sA, sm.sB);
  for (int i = threadIdx.x; i < m; i += NT) {
    double a = 0.0;
    for (int j = 0; j < d; ++j) {
      const cd t = T[i + (size_t)m * j], c = At[i + (size_t)m * j];
      a += t.x * c.x + t.y * c.y;
    }
    out[i] = a;
  }
  __syncthreads();
}

// Hermitian eigendecomposition W = V diag(lam) V' warm-started from the basis U.  W is overwritten by the
// rotated matrix (eigenvalues on its diagonal); V receives the eigenvectors.  T: d x d scratch.
__device__ inline int warm_eig(cd* W, const cd* U, cd* V, cd* T, int d, const PlSmem& sm, long long* tc = nullptr,
                               double act_thr = -INFINITY, bool rotated = false) {
  const int tid = threadIdx.x;
  const long long t0 = clock64();
  // T = W U   (skipped when the caller already formed U' W U in W)
  if (!rotated) cta_gemm<false, true>(d, d, d, [&](int i, int k) { return W[i + (size_t)d * k]; },
                        [&](int k, int j) { return U[k + (size_t)d * j]; },
                        [&](int i, int j, cd v) { T[i + (size_t)d * j] = v; }, sm.sA, sm.sB);
  // W <- U' T (Hermitian up to rounding; the diagonal is made real)
  if (!rotated) cta_gemm<true, true, true>(d, d, d, [&](int i, int k) { return cconj(U[k + (size_t)d * i]); },
                             [&](int k, int j) { return T[k + (size_t)d * j]; },
                             [&](int i, int j, cd v) {
                               if (i == j) v.y = 0.0;
                               W[i + (size_t)d * j] = v;
                               if ((i / PG_TM) != (j / PG_TM)) W[j + (size_t)d * i] = cconj(v);   // mirrored tile
                             }, sm.sA, sm.sB);
  double gm = 0.0;
  for (int i = tid; i < d; i += NT) gm = fmax(gm, fabs(W[i + (size_t)d * i].x));
  for (size_t e = tid; e < (size_t)d * d; e += NT) V[e] = U[e];
  // block max
  gm = fmax(gm, __shfl_xor_sync(0xffffffffu, gm, 16));
  gm = fmax(gm, __shfl_xor_sync(0xffffffffu, gm, 8));
  gm = fmax(gm, __shfl_xor_sync(0xffffffffu, gm, 4));
  gm = fmax(gm, __shfl_xor_sync(0xffffffffu, gm, 2));
  gm = fmax(gm, __shfl_xor_sync(0xffffffffu, gm, 1));
  if ((tid & 31) == 0) sm.red[tid >> 5] = gm;
  __syncthreads();
  gm = 0.0;
  for (int w = 0; w < NW; ++w) gm = fmax(gm, sm.red[w]);
  __syncthreads();
  const double skip = 1.0e-13 * gm;
  const long long t1 = clock64();
  const int sw = block_jacobi_heig(W, d, V, d, d, sm.S, sm.Sb, sm.Q, sm.tab, 30, false, 1.0e-12, skip * skip,
                                   tc ? tc + 4 : nullptr, act_thr > -INFINITY ? act_thr - 1.0e-6 * gm : act_thr,
                                   (1.0e-4 * gm) * (1.0e-4 * gm));
  if (tc) { tc[0] += t1 - t0; tc[1] += clock64() - t1; }
  return sw;
}

__device__ inline void run_phaselift(const PlTask& tk, int n, int maxm, const PlOpts& o, cd* wsbase, unsigned char* smraw) {
  const int tid = threadIdx.x;
  const int m = tk.m;
  PlSmem sm = pl_carve(smraw, n, maxm);
  const int dcap = pl_dcap(n, maxm);
  PlWs ws = pl_ws(wsbase, dcap, maxm);
  auto Aget = [&](int i, int k) -> cd {
    if (tk.A_cm) return tk.A_cm[i + (size_t)m * k];
    return cscale(tk.cb[(size_t)tk.rows[i] * n + k], tk.scale);
  };
  for (int i = tid; i < m; i += NT) sm.b[i] = tk.y[i];
  __syncthreads();

  // ---- operator: row-space reduction (m < n) or A itself
  int d = n;
  bool reduced = false;
  if (o.reduce && m < n) {
    cd* S = ws.W;   // m x m
    if (tk.A_cm) {
      cta_gemm<false, false>(m, m, n, [&](int i, int k) { return tk.A_cm[i + (size_t)m * k]; },
                             [&](int k, int j) { return cconj(tk.A_cm[j + (size_t)m * k]); },
                             [&](int i, int j, cd v) { if (i == j) v.y = 0.0; S[i + (size_t)m * j] = v; }, sm.sA, sm.sB);
    } else {
      cta_gemm<true, true>(m, m, n, [&](int i, int k) { return cscale(tk.cb[(size_t)tk.rows[i] * n + k], tk.scale); },
                           [&](int k, int j) { return cscale(cconj(tk.cb[(size_t)tk.rows[j] * n + k]), tk.scale); },
                           [&](int i, int j, cd v) { if (i == j) v.y = 0.0; S[i + (size_t)m * j] = v; }, sm.sA, sm.sB);
    }
    double dmax = 0.0;
    for (int i = 0; i < m; ++i) dmax = fmax(dmax, S[i + (size_t)m * i].x);
    // right-looking Cholesky, lower triangle, in place
    bool ok = dmax > 0.0;
    for (int k = 0; k < m && ok; ++k) {
      const double pv = S[k + (size_t)m * k].x;
      if (!(pv > 1.0e-8 * dmax)) { ok = false; break; }   // (uniform: every thread reads the same value)
      const double rp = 1.0 / sqrt(pv);
      __syncthreads();
      for (int i = k + tid; i < m; i += NT) S[i + (size_t)m * k] = cscale(S[i + (size_t)m * k], rp);
      __syncthreads();
      const int rem = m - k - 1;
      for (int e = tid; e < rem * rem; e += NT) {
        const int i = k + 1 + e % rem, j = k + 1 + e / rem;
        if (i >= j) {
          const cd li = S[i + (size_t)m * k], lj = S[j + (size_t)m * k];
          cd v = S[i + (size_t)m * j];
          v.x -= li.x * lj.x + li.y * lj.y;
          v.y -= li.y * lj.x - li.x * lj.y;
          if (i == j) v.y = 0.0;
          S[i + (size_t)m * j] = v;
        }
      }
      __syncthreads();
    }
    __syncthreads();
    if (ok) {
      reduced = true;
      d = m;
      for (int e = tid; e < m * m; e += NT) {
        const int i = e % m, j = e / m;
        ws.At[e] = (i >= j) ? S[e] : cmk(0.0, 0.0);
      }
    }
    __syncthreads();
  }
  if (!reduced && n > dcap) {
    // n > PL_DMAX and no row-space factor (rank-deficient rows): the n x n iteration is outside this kernel
    for (int k = tid; k < n; k += NT) tk.sig[k] = cmk(NAN, NAN);
    if (tk.info && tid == 0) {
      for (int q = 0; q < PL_INFO; ++q) tk.info[q] = 0.0;
      tk.info[3] = -2.0; tk.info[6] = n;
    }
    __syncthreads();
    return;
  }
  if (!reduced) {
    for (size_t e = tid; e < (size_t)m * n; e += NT) ws.At[e] = Aget((int)(e % m), (int)(e / m));
    __syncthreads();
  }
  const size_t dd = (size_t)d * d;

  // ---- tfocs_initialize.m:418-478, 521, 582-592 with x0 = zeros(n) (MyPhaseLift.m:95)
  for (size_t e = tid; e < dd; e += NT) {
    const int i = (int)(e % d), j = (int)(e / d);
    ws.X[0][e] = cmk(0.0, 0.0);
    ws.Z[0][e] = cmk(0.0, 0.0);
    ws.U[0][e] = cmk(i == j ? 1.0 : 0.0, 0.0);
  }
  for (int i = tid; i < m; i += NT) { sm.Ax[i] = 0.0; sm.Az[i] = 0.0; sm.Ay[i] = 0.0; }
  __syncthreads();
  int xc = 0, zc = 0, uc = 0;      // current buffers of x, z, U
  double L = o.L0, theta = INFINITY;
  double f_x = 0.0;
  {
    double v[1] = {0.0};
    for (int i = tid; i < m; i += NT) v[0] += sm.b[i] * sm.b[i];
    block_sum<1>(v, sm.red);
    f_x = 0.5 * v[0];
  }
  double f_y = f_x;
  bool y_is_fresh = true;          // A_y / f_y valid for the current y (tfocs_AT.m:49 clears them when theta < 1)
  bool zold_diag = false;          // U[uc]' z_old U[uc] = diag(sacc): z_old is the prox output whose basis U[uc] is
  int cntr_Ay = 0, cntr_Ax = 0;
  bool force_Ax = false;           // cntr_Ax = Inf (tfocs_backtrack.m:18)
  bool backtrack_simple = true;
  int backtrack_steps = 0, restart_iter = 0, n_iter = 0, status = 0;
  int n_prox = 0, n_bt = 0, rank = 0;
  long long tc[6] = {0, 0, 0, 0, 0, 0}, n_sweeps = 0;   // [5]: cycles inside the 32 x 32 subproblem solves
  double xy_sq = 0.0;

  while (true) {                                                     // tfocs_AT.m:20
    // x_old / z_old are X[xc] / Z[zc]; A_x_old / A_z_old:
    for (int i = tid; i < m; i += NT) { sm.Axo[i] = sm.Ax[i]; sm.Azo[i] = sm.Az[i]; }
    __syncthreads();
    const double L_old = L;                                          // :28
    L = L * o.alpha;                                                 // :29
    const double theta_old = theta;                                  // :30
    double norm_x2 = 0.0, norm_dx2 = 0.0;
    const cd* XO = ws.X[xc];
    const cd* ZO = ws.Z[zc];
    cd* XN = ws.X[xc ^ 1];
    cd* ZN = ws.Z[zc ^ 1];
    while (true) {                                                   // :31 backtracking loop
      theta = isinf(theta_old) ? 1.0 : 2.0 / (1.0 + sqrt(1.0 + 4.0 * (L / L_old) / (theta_old * theta_old)));   // :34
      if (theta < 1.0) {                                             // :37-50
        if (cntr_Ay >= o.cntr_reset) {
          // explicit A(y) = (1-theta) A(x_old) + theta A(z_old) from the matrices themselves
          lifted_forward(ws.At, m, d, XO, ws.T, sm.Ay, sm);
          lifted_forward(ws.At, m, d, ZO, ws.T, sm.g, sm);
          for (int i = tid; i < m; i += NT) sm.Ay[i] = (1.0 - theta) * sm.Ay[i] + theta * sm.g[i];
          cntr_Ay = 0;
        } else {
          cntr_Ay++;
          for (int i = tid; i < m; i += NT) sm.Ay[i] = (1.0 - theta) * sm.Axo[i] + theta * sm.Azo[i];
        }
        __syncthreads();
        y_is_fresh = false;
      }
      // g_Ay = A_y - b, f_y (smooth_quad.m:149-155 through tfocs_initialize.m:342)
      {
        double v[1] = {0.0};
        for (int i = tid; i < m; i += NT) {
          const double gi = sm.Ay[i] - sm.b[i];
          sm.g[i] = gi;
          v[0] += gi * gi;
        }
        block_sum<1>(v, sm.red);
        if (!y_is_fresh) f_y = 0.5 * v[0];
        y_is_fresh = true;
      }
      const double step = 1.0 / (theta * L);                         // :59
      const double tau = o.lam * step;                               // prox_trace.m:77
      long long t0 = clock64();
      cd* V = ws.U[uc ^ 1];
      long long t1;
      if (zold_diag) {
        // z_old = U diag(sacc) U' exactly (it was built from these columns), so
        //   U' W U = diag(sacc) - step * (A U)' diag(g) (A U):  one m x d x d and one Hermitian d x d x m product
        // instead of forming W and transforming it (three products)
        const cd* U = ws.U[uc];
        cta_gemm<false, true>(m, d, d, [&](int i, int k) { return ws.At[i + (size_t)m * k]; },
                              [&](int k, int j) { return U[k + (size_t)d * j]; },
                              [&](int i, int j, cd v) { ws.T[i + (size_t)m * j] = v; }, sm.sA, sm.sB);
        cta_gemm<true, true, true>(d, d, m,
                                   [&](int i, int k) { return cscale(cconj(ws.T[k + (size_t)m * i]), sm.g[k]); },
                                   [&](int k, int j) { return ws.T[k + (size_t)m * j]; },
                                   [&](int i, int j, cd v) {
                                     cd w = cmk(-step * v.x, -step * v.y);
                                     if (i == j) { w.x += sm.sacc[i]; w.y = 0.0; }
                                     ws.W[i + (size_t)d * j] = w;
                                     if ((i / PG_TM) != (j / PG_TM)) ws.W[j + (size_t)d * i] = cconj(w);
                                   }, sm.sA, sm.sB);
        t1 = clock64();
        tc[0] += t1 - t0;
        n_sweeps += warm_eig(ws.W, ws.U[uc], V, ws.T, d, sm, &tc[1], tau, true);
      } else {
      // W = z_old - step * A' diag(g) A                               (:60 argument, initializeLinopPR.m:65)
      cta_gemm<true, true, true>(d, d, m,
                                 [&](int i, int k) { return cscale(cconj(ws.At[k + (size_t)m * i]), sm.g[k]); },
                                 [&](int k, int j) { return ws.At[k + (size_t)m * j]; },
                                 [&](int i, int j, cd v) {
                                   const cd z0 = ZO[i + (size_t)d * j];
                                   cd w = cmk(z0.x - step * v.x, z0.y - step * v.y);
                                   if (i == j) w.y = 0.0;
                                   ws.W[i + (size_t)d * j] = w;
                                   if ((i / PG_TM) != (j / PG_TM)) ws.W[j + (size_t)d * i] = cconj(w);   // mirrored tile
                                 }, sm.sA, sm.sB);
      t1 = clock64();
      tc[0] += t1 - t0;
      n_sweeps += warm_eig(ws.W, ws.U[uc], V, ws.T, d, sm, &tc[1], tau);   // prox_trace.m:92
      }
      zold_diag = false;   // U[uc ^ 1] now belongs to this attempt; set again when the attempt is accepted
      uc ^= 1;
      t0 = clock64();
      n_prox++;
      // kept eigenpairs: s = D - tau > 0                              (prox_trace.m:140-142)
      if (tid == 0) sm.ib[0] = 0;
      __syncthreads();
      for (int k = tid; k < d; k += NT) {
        const double s = ws.W[k + (size_t)d * k].x - tau;
        if (s > 0.0) {
          const int slot = atomicAdd(&sm.ib[0], 1);
          sm.idx[slot] = k;
        }
      }
      __syncthreads();
      const int kact = sm.ib[0];
      // (order the kept list by column index so the summation order is deterministic)
      if (tid == 0) {
        for (int a = 1; a < kact; ++a) {
          const int key = sm.idx[a];
          int c = a - 1;
          while (c >= 0 && sm.idx[c] > key) { sm.idx[c + 1] = sm.idx[c]; --c; }
          sm.idx[c + 1] = key;
        }
      }
      __syncthreads();
      for (int k = tid; k < kact; k += NT) sm.sv[k] = ws.W[sm.idx[k] + (size_t)d * sm.idx[k]].x - tau;
      __syncthreads();
      rank = kact;
      // keep the kept eigenvectors in the leading columns of the basis (column order is free): the next warm
      // start then finds every entry above the threshold in block 0 and its later sweeps touch 7 of 28 block pairs
      if (kact <= 16) {
        for (int a = 0; a < kact; ++a) {
          const int c = sm.idx[a];                     // (ascending, so c >= a and column c has not moved yet)
          if (c != a) {
            for (int i = tid; i < d; i += NT) {
              const cd t = V[i + (size_t)d * a];
              V[i + (size_t)d * a] = V[i + (size_t)d * c];
              V[i + (size_t)d * c] = t;
            }
          }
        }
        __syncthreads();
        for (int a = tid; a < kact; a += NT) sm.idx[a] = a;
        __syncthreads();
      }
      // z = V_+ diag(s) V_+'                                          (prox_trace.m:147-149)
      if (kact == 0) {
        for (size_t e = tid; e < dd; e += NT) ZN[e] = cmk(0.0, 0.0);
        for (int i = tid; i < m; i += NT) sm.Az[i] = 0.0;
        __syncthreads();
      } else {
        cta_gemm<false, false, true>(d, d, kact,
                                     [&](int i, int k) { return cscale(V[i + (size_t)d * sm.idx[k]], sm.sv[k]); },
                                     [&](int k, int j) { return cconj(V[j + (size_t)d * sm.idx[k]]); },
                                     [&](int i, int j, cd v) {
                                       if (i == j) v.y = 0.0;
                                       ZN[i + (size_t)d * j] = v;
                                       if ((i / PG_TM) != (j / PG_TM)) ZN[j + (size_t)d * i] = cconj(v);
                                     }, sm.sA, sm.sB);
        // A_z = sum_k s_k |A v_k|^2                                   (:61)
        cta_gemm<false, true>(m, kact, d, [&](int i, int k) { return ws.At[i + (size_t)m * k]; },
                              [&](int k, int j) { return V[k + (size_t)d * sm.idx[j]]; },
                              [&](int i, int j, cd v) { ws.T[i + (size_t)m * j] = v; }, sm.sA, sm.sB);
        for (int i = tid; i < m; i += NT) {
          double a = 0.0;
          for (int k = 0; k < kact; ++k) a += sm.sv[k] * cabs2(ws.T[i + (size_t)m * k]);
          sm.Az[i] = a;
        }
        __syncthreads();
      }
      t1 = clock64();
      tc[3] += t1 - t0;
      // x = (1-theta) x_old + theta z and the norms of the iterate tests      (:64-78)
      double v3[3] = {0.0, 0.0, 0.0};
      if (theta == 1.0) {
        for (size_t e = tid; e < dd; e += NT) {
          const cd z = ZN[e], xo = XO[e], zo = ZO[e];
          XN[e] = z;
          v3[0] += cabs2(z);
          v3[1] += cabs2(csub(z, xo));
          v3[2] += cabs2(csub(z, zo));
        }
        for (int i = tid; i < m; i += NT) sm.Ax[i] = sm.Az[i];
      } else {
        const double w0 = 1.0 - theta;
        for (size_t e = tid; e < dd; e += NT) {
          const cd z = ZN[e], xo = XO[e], zo = ZO[e];
          const cd x = cmk(w0 * xo.x + theta * z.x, w0 * xo.y + theta * z.y);
          XN[e] = x;
          v3[0] += cabs2(x);
          v3[1] += cabs2(csub(x, xo));
          v3[2] += cabs2(csub(z, zo));
        }
      }
      block_sum<3>(v3, sm.red);
      norm_x2 = v3[0];
      norm_dx2 = v3[1];
      if (theta != 1.0) {
        if (force_Ax || cntr_Ax >= o.cntr_reset) {
          cntr_Ax = 0;
          force_Ax = false;
          lifted_forward(ws.At, m, d, XN, ws.T, sm.Ax, sm);
        } else {
          cntr_Ax++;
          for (int i = tid; i < m; i += NT) sm.Ax[i] = (1.0 - theta) * sm.Axo[i] + theta * sm.Az[i];
          __syncthreads();
        }
      } else {
        __syncthreads();
      }
      tc[4] += clock64() - t1;
      // ---- tfocs_backtrack.m:4-46
      if (o.beta >= 1.0) break;
      xy_sq = theta * theta * v3[2];                                 // |x - y|^2, x - y = theta (z - z_old)
      if (xy_sq == 0.0) break;                                       // :16
      if (xy_sq / norm_x2 < 2.220446049250313e-16) force_Ax = true;  // :18
      double localL;
      {
        double v4[4] = {0.0, 0.0, 0.0, 0.0};   // |A_x - b|^2, <g, A_z - A_z_old>, |A_x - A_y|^2
        for (int i = tid; i < m; i += NT) {
          const double rx = sm.Ax[i] - sm.b[i];
          v4[0] += rx * rx;
          v4[1] += sm.g[i] * (sm.Az[i] - sm.Azo[i]);
          const double dxy = sm.Ax[i] - sm.Ay[i];
          v4[2] += dxy * dxy;
        }
        block_sum<4>(v4, sm.red);
        if (backtrack_simple) {                                      // :21-27
          f_x = 0.5 * v4[0];
          const double q_x = f_y + theta * v4[1] + 0.5 * L * xy_sq;
          localL = L + 2.0 * fmax(f_x - q_x, 0.0) / xy_sq;
          backtrack_simple = fabs(f_y - f_x) >= o.backtrack_tol * fmax(fabs(f_x), fabs(f_y));
        } else {                                                     // :29-32
          f_x = 0.5 * v4[0];
          localL = 2.0 * v4[2] / xy_sq;
        }
      }
      backtrack_steps++;
      if (localL <= L) break;                                        // :37 (Lexact = Inf)
      if (!isinf(localL)) L = localL; else localL = L;               // :38-42
      L = fmax(localL, L / o.beta);                                  // :43
      n_bt++;
    }
    xc ^= 1;
    zc ^= 1;
    // the accepted z is V_+ diag(s) V_+' with V = U[uc]: remember its diagonal form for the next prox
    for (int i = tid; i < d; i += NT) sm.sacc[i] = 0.0;
    __syncthreads();
    for (int k = tid; k < rank; k += NT) sm.sacc[sm.idx[k]] = sm.sv[k];
    __syncthreads();
    zold_diag = true;
    // ---- tfocs_iterate.m:8-33
    n_iter++;
    const double norm_x = sqrt(norm_x2), norm_dx = sqrt(norm_dx2);
    if (f_y != f_y) status = PL_ST_NAN;
    else if (norm_dx == 0.0) { if (n_iter > 1) status = PL_ST_DX0; }
    else if (norm_dx < o.tol * fmax(norm_x, 1.0)) status = PL_ST_TOL;
    else if (n_iter == o.maxIts) status = PL_ST_MAXIT;
    else if (backtrack_steps > 0 && xy_sq == 0.0) status = PL_ST_SMALLSTEP;
    if (status) break;
    // ---- tfocs_iterate.m:331-362: fixed-period restart
    backtrack_steps = 0;
    if (n_iter - restart_iter == o.restart) {
      restart_iter = n_iter;
      backtrack_simple = true;
      theta = INFINITY;
      // y = x, z = x
      for (size_t e = tid; e < dd; e += NT) ws.Z[zc][e] = ws.X[xc][e];
      for (int i = tid; i < m; i += NT) { sm.Ay[i] = sm.Ax[i]; sm.Az[i] = sm.Ax[i]; }
      __syncthreads();
      y_is_fresh = false;   // f_y = f_x: recomputed from A_y = A_x
      zold_diag = false;    // z = x is not diagonal in the basis U
    }
  }

  // ---- MyPhaseLift.m:106-107: leading eigenpair of the x sequence (tfocs_cleanup.m:28-52 returns x)
  for (size_t e = tid; e < dd; e += NT) ws.W[e] = ws.X[xc][e];
  __syncthreads();
  cd* V = ws.U[uc ^ 1];
  warm_eig(ws.W, ws.U[uc], V, ws.T, d, sm);
  int kbest = 0;
  double lbest = -INFINITY;
  for (int k = 0; k < d; ++k) {
    const double l = ws.W[k + (size_t)d * k].x;
    if (l > lbest) { lbest = l; kbest = k; }
  }
  const double amp = sqrt(fmax(lbest, 0.0));
  if (reduced) {
    // sig = amp * A' w with L' w = u (back substitution; w kept in the first column of T)
    cd* w = ws.T;
    for (int i = tid; i < m; i += NT) w[i] = V[i + (size_t)d * kbest];
    __syncthreads();
    for (int i = m - 1; i >= 0; --i) {
      const cd wi = cscale(w[i], 1.0 / ws.At[i + (size_t)m * i].x);
      __syncthreads();
      if (tid == 0) w[i] = wi;
      for (int j = tid; j < i; j += NT) {
        const cd l = ws.At[i + (size_t)m * j];      // L(i, j); (L')(j, i) = conj(L(i, j))
        cd v = w[j];
        v.x -= l.x * wi.x + l.y * wi.y;
        v.y -= l.x * wi.y - l.y * wi.x;
        w[j] = v;
      }
      __syncthreads();
    }
    for (int k = tid; k < n; k += NT) {
      cd a = cmk(0.0, 0.0);
      for (int i = 0; i < m; ++i) cfmac(a, Aget(i, k), w[i]);
      tk.sig[k] = cscale(a, amp);
    }
  } else {
    for (int k = tid; k < n; k += NT) tk.sig[k] = cscale(V[k + (size_t)d * kbest], amp);
  }
  if (tk.info && tid == 0) {
    tk.info[0] = n_iter; tk.info[1] = n_prox; tk.info[2] = n_bt; tk.info[3] = status;
    tk.info[4] = rank; tk.info[5] = L; tk.info[6] = d; tk.info[7] = lbest;
    tk.info[8] = (double)n_sweeps;
    for (int q = 0; q < 5; ++q) tk.info[9 + q] = (double)tc[q];
    tk.info[14] = (double)tc[5]; tk.info[15] = 0.0;
  }
  __syncthreads();
}

// The kernel lives in its own translation unit (pl_kernel.cu): compiled together with the ADMM kernels in one module
// the inliner outlined parts of the prox (stack 720 -> 1232 bytes per thread) and config 3 lost 20 % (r02 measurement).
// api.cu reaches it through these three host functions.
cudaError_t pl_kernel_set_smem(size_t smem);
cudaError_t pl_kernel_occupancy(int* per_sm, size_t smem);
cudaError_t pl_kernel_launch(int grid, size_t smem, cudaStream_t stream, const PlTask* tasks, int ntasks, int n, int maxm,
                             PlOpts o, cd* wsbase, size_t ws_stride, int* counter);

#ifdef TWOACE_PL_KERNEL_TU
__global__ void __launch_bounds__(NT, 2) phaselift_kernel(const PlTask* tasks, int ntasks, int n, int maxm, PlOpts o,
                                                       cd* wsbase, size_t ws_stride, int* counter) {
  extern __shared__ __align__(16) unsigned char pl_smem_raw[];
  __shared__ int s_task;
  while (true) {
    __syncthreads();
    if (threadIdx.x == 0) s_task = atomicAdd(counter, 1);
    __syncthreads();
    const int t = s_task;
    if (t >= ntasks) break;
    run_phaselift(tasks[t], n, maxm, o, wsbase + (size_t)blockIdx.x * ws_stride, pl_smem_raw);
  }
}

cudaError_t pl_kernel_set_smem(size_t smem) {
  return cudaFuncSetAttribute(phaselift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}
cudaError_t pl_kernel_occupancy(int* per_sm, size_t smem) {
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, phaselift_kernel, NT, smem);
}
cudaError_t pl_kernel_launch(int grid, size_t smem, cudaStream_t stream, const PlTask* tasks, int ntasks, int n, int maxm,
                             PlOpts o, cd* wsbase, size_t ws_stride, int* counter) {
  phaselift_kernel<<<grid, NT, smem, stream>>>(tasks, ntasks, n, maxm, o, wsbase, ws_stride, counter);
  return cudaGetLastError();
}
#endif

}  // namespace twoace
