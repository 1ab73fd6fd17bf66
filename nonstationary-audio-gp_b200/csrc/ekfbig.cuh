// Dense RTS smoother of the iterated-EKF variant for state dimensions 80 < n <= 160 (BASELINE config C4 with
// matern32 subbands: D = 32, N = 3, n = 137; matlab/gf_giekf_modulator_nmf_constraints.m:221-253 with
// matlab/unifying_prob_tf/cf_matern32_to_ss.m:93-117).
//
// Same parallel-in-time scan as ekfscan.cuh -- elements (one CTA per step), chunk aggregates, carry, apply; same
// DsArgs and scratch arrays -- but an n x n FP64 matrix no longer fits in shared memory three or four at a time
// (137^2 doubles = 150 KB), so the operands stay in HBM / L2 (a CTA's working set is four matrices, < 1 MB) and are
// read through L1 by register-tiled products on the FP64 FMA pipe.  A CTA is slow on one step (~0.3 ms), the pass is
// not: all steps of a segment are in flight at once.  Blocked Cholesky (8-wide panels, diagonal blocks inverted in
// registers by one thread) and column-parallel triangular solves; only the association of the products differs
// from the reference (tolerance class 1e-6).
#pragma once
#include "common.cuh"
#include "ekf.cuh"
#include "ekfscan.cuh"
#include "fastmath.cuh"

namespace nsagp {

constexpr int kDgThreads = 640;
constexpr int kDgMaxNP = 160;
constexpr int kDgWsMats = 4;        // n x n (padded to a multiple of 8) matrices of global workspace per CTA

// C = op(X) op(Y), n x n, column-major with leading dimensions ldx / ldy, 4 x 4 register tiles, operands in global
// memory.  epi(r, c, v) for r, c < n.  TX: op(X) = X'.  TY: op(Y) = Y'.
template <bool TX, bool TY, class Epi>
__device__ __forceinline__ void dg_gemm(const double* X, int ldx, const double* Y, int ldy, int n, Epi epi) {
  // (no __restrict__: the operands were written earlier in this kernel by other threads of the CTA; they must be read
  // with coherent loads, not through the read-only path)
  const int nt = (n + 3) >> 2;
  for (int t = threadIdx.x; t < nt * nt; t += blockDim.x) {
    const int r0 = (t % nt) * 4, c0 = (t / nt) * 4;
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    // clamped indices: rows / columns beyond n repeat the last one, their results are never emitted
    int ri[4], ci[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { ri[i] = min(r0 + i, n - 1); ci[i] = min(c0 + i, n - 1); }
#pragma unroll 2
    for (int k = 0; k < n; ++k) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = TX ? X[k + (size_t)ri[i] * ldx] : X[ri[i] + (size_t)k * ldx];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = TY ? Y[ci[j] + (size_t)k * ldy] : Y[k + (size_t)ci[j] * ldy];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (r0 + i < n && c0 + j < n) epi(r0 + i, c0 + j, acc[i][j]);
  }
}

// Cholesky factor of the 8 x 8 diagonal tile of Cm at (8p, 8p) (leading dimension ld) and the inverse of that factor
// (lower triangular) -> dinv[r + 8 c].  One thread, registers.  Returns false if a pivot is not positive.
__device__ __forceinline__ bool dg_chol8_inv(const double* Cm, int ld, int p, double* dinv) {
  double a[8][8];
  const double* t = Cm + (8 * p) + (size_t)(8 * p) * ld;
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int r = 0; r < 8; ++r) a[r][c] = (r >= c) ? t[r + (size_t)c * ld] : 0.0;
  bool ok = true;
  double rs[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    double d = a[j][j];
    if (!(d > 0.0) || !(d < 1e300)) { ok = false; d = 1.0; }
    rs[j] = rsqrt_fast(d);
#pragma unroll
    for (int r = j + 1; r < 8; ++r) a[r][j] *= rs[j];
#pragma unroll
    for (int c = j + 1; c < 8; ++c)
#pragma unroll
      for (int r = c; r < 8; ++r) a[r][c] = fma(-a[r][j], a[c][j], a[r][c]);
  }
  double li[8][8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (r < c) li[r][c] = 0.0;
      else if (r == c) li[r][c] = rs[c];
      else {
        double s = 0.0;
#pragma unroll
        for (int k = c; k < r; ++k) s = fma(a[r][k], li[k][c], s);
        li[r][c] = -rs[r] * s;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int r = 0; r < 8; ++r) dinv[r + 8 * c] = li[r][c];
  return ok;
}

// In-place blocked Cholesky of the NP x NP matrix Cm (lower triangle; NP a multiple of 8, padding = identity).
// Afterwards the strictly-lower tiles hold the factor and dinv[p] the inverse of the p-th diagonal block of it.
__device__ __forceinline__ void dg_cholesky(double* Cm, int NP, double* dinv, int* fail) {
  const int tid = threadIdx.x, nth = blockDim.x, NT = NP >> 3;
  for (int p = 0; p < NT; ++p) {
    __syncthreads();
    if (tid == 0 && !dg_chol8_inv(Cm, NP, p, dinv + p * 64)) *fail = 1;
    __syncthreads();
    const double* Dp = dinv + p * 64;
    // panel: rows below the diagonal tile, L(r, 8p..8p+7) = A(r, 8p..8p+7) inv(Lpp)'
    for (int r = 8 * (p + 1) + tid; r < NP; r += nth) {
      double c[8], o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) c[k] = Cm[r + (size_t)(8 * p + k) * NP];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) if (k <= j) s = fma(c[k], Dp[j + 8 * k], s);
        o[j] = s;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) Cm[r + (size_t)(8 * p + j) * NP] = o[j];
    }
    __syncthreads();
    // trailing update of the lower triangle
    const int t0 = 8 * (p + 1), nt = NP - t0;
    for (int e = tid; e < nt * nt; e += nth) {
      const int rr = e % nt, cc = e / nt;
      if (rr < cc) continue;
      const int r = t0 + rr, c = t0 + cc;
      double s = Cm[r + (size_t)c * NP];
#pragma unroll
      for (int k = 0; k < 8; ++k) s = fma(-Cm[r + (size_t)(8 * p + k) * NP], Cm[c + (size_t)(8 * p + k) * NP], s);
      Cm[r + (size_t)c * NP] = s;
    }
  }
  __syncthreads();
}

// X <- C^-1 X (forward) then X <- C^-T X (backward), C = the factor left by dg_cholesky; one thread per column of X.
__device__ __forceinline__ void dg_solve_columns(const double* Cm, const double* dinv, double* Xm, int NP, bool backward) {
  const int NT = NP >> 3;
  for (int c = threadIdx.x; c < NP; c += blockDim.x) {
    double* x = Xm + (size_t)c * NP;
    if (!backward) {
      for (int q = 0; q < NT; ++q) {
        double acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = x[8 * q + i];
        for (int k = 0; k < 8 * q; ++k) {
          const double xk = x[k];
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] = fma(-Cm[(8 * q + i) + (size_t)k * NP], xk, acc[i]);
        }
        const double* Dq = dinv + q * 64;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          double s = 0.0;
#pragma unroll
          for (int k = 0; k < 8; ++k) if (k <= i) s = fma(Dq[i + 8 * k], acc[k], s);
          x[8 * q + i] = s;
        }
      }
    } else {
      for (int q = NT - 1; q >= 0; --q) {
        double acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = x[8 * q + i];
        for (int k = 8 * (q + 1); k < NP; ++k) {
          const double xk = x[k];
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] = fma(-Cm[k + (size_t)(8 * q + i) * NP], xk, acc[i]);
        }
        const double* Dq = dinv + q * 64;
#pragma unroll
        for (int i = 0; i < 8; ++i) {                 // inv(Lqq)' acc
          double s = 0.0;
#pragma unroll
          for (int k = 0; k < 8; ++k) if (k >= i) s = fma(Dq[k + 8 * i], acc[k], s);
          x[8 * q + i] = s;
        }
      }
    }
  }
}

__device__ __forceinline__ double* dg_ws(double* ws, int NP, int slot) {
  return ws + ((size_t)blockIdx.x * kDgWsMats + slot) * (size_t)NP * NP;
}

// ---------------------------------------------------------------------------------------------
// 1. elements: G_k' -> Gt, L_k over PS_k, g_k -> gv   (as ds_elements_kernel)
__global__ void __launch_bounds__(kDgThreads, 1)
dg_elements_kernel(const __grid_constant__ DsArgs g, double* __restrict__ ws, int NP) {
  const EkfArgs& a = g.ekf;
  const int tid = threadIdx.x, nth = blockDim.x;
  const int M = a.M, n = a.n, BM = a.BM;
  double* Pm = dg_ws(ws, NP, 0);
  double* Xm = dg_ws(ws, NP, 1);
  double* Cm = dg_ws(ws, NP, 2);
  extern __shared__ double sm[];
  double* sA = sm;
  double* sQ = sA + M * BM * BM;
  double* dinv = sQ + M * BM * BM;              // [NP/8][64]
  double* mf = dinv + (NP >> 3) * 64;
  double* Amf = mf + kDgMaxNP;
  __shared__ int s_blk[kDgMaxNP];
  __shared__ int s_fail;
  for (int i = tid; i < M * BM * BM; i += nth) { sA[i] = a.A[i]; sQ[i] = a.Q[i]; }
  for (int b = tid; b < M; b += nth)
    for (int i = a.off[b]; i < a.off[b + 1]; ++i) s_blk[i] = b;
  if (tid == 0) s_fail = 0;
  const size_t nn = (size_t)n * n;
  for (long long k = g.seg_k0 + blockIdx.x; k < g.seg_k1; k += gridDim.x) {
    __syncthreads();
    const double* Pg = a.PS + (size_t)k * a.ps_stride;
    for (int i = tid; i < NP * NP; i += nth) {
      const int r = i % NP, c = i / NP;
      Pm[i] = (r < n && c < n) ? Pg[r + (size_t)c * n] : 0.0;
    }
    for (int i = tid; i < NP; i += nth) mf[i] = (i < n) ? a.MS[k * n + i] : 0.0;
    __syncthreads();
    // X = A PS_k (A block diagonal), zero padding
    for (int i = tid; i < NP * NP; i += nth) {
      const int r = i % NP, c = i / NP;
      double s = 0.0;
      if (r < n && c < n) {
        const int b = s_blk[r], o = a.off[b], nb = a.off[b + 1] - o;
        for (int l = 0; l < nb; ++l) s = fma(sA[b * BM * BM + (r - o) + l * BM], Pm[(o + l) + (size_t)c * NP], s);
      }
      Xm[i] = s;
    }
    if (tid < n) {                                   // A * MS(:,k)
      const int b = s_blk[tid], o = a.off[b], nb = a.off[b + 1] - o;
      double s = 0.0;
      for (int c = 0; c < nb; ++c) s = fma(sA[b * BM * BM + (tid - o) + c * BM], mf[o + c], s);
      Amf[tid] = s;
    } else if (tid < NP) Amf[tid] = 0.0;
    __syncthreads();
    // PSkp = X A' + Q (:229), identity padding
    for (int i = tid; i < NP * NP; i += nth) {
      const int r = i % NP, c = i / NP;
      double s;
      if (r < n && c < n) {
        const int b = s_blk[c], o = a.off[b], nb = a.off[b + 1] - o;
        s = (s_blk[r] == b) ? sQ[b * BM * BM + (r - o) + (c - o) * BM] : 0.0;
        for (int l = 0; l < nb; ++l) s = fma(Xm[r + (size_t)(o + l) * NP], sA[b * BM * BM + (c - o) + l * BM], s);
      } else {
        s = (r == c) ? 1.0 : 0.0;
      }
      Cm[i] = s;
    }
    dg_cholesky(Cm, NP, dinv, &s_fail);              // (:232)
    dg_solve_columns(Cm, dinv, Xm, NP, false);       // Y = C^-1 A PS_k
    __syncthreads();
    // L = PS_k - Y'Y (= PS_k - G PSkp G'), written over PS_k in HBM
    {
      double* Lg = a.PS + (size_t)k * a.ps_stride;
      dg_gemm<true, false>(Xm, NP, Xm, NP, n, [&](int r, int c, double v) { Lg[r + (size_t)c * n] = Pm[r + (size_t)c * NP] - v; });
    }
    __syncthreads();
    dg_solve_columns(Cm, dinv, Xm, NP, true);        // G' = C^-T Y
    __syncthreads();
    const long long ks = k - g.seg_k0;
    double* Gg = g.Gt + (size_t)ks * nn;
    for (int i = tid; i < n * n; i += nth) { const int r = i % n, c = i / n; Gg[i] = Xm[r + (size_t)c * NP]; }
    if (tid < n) {                                   // g = MS_k - G A MS_k,  G(r, c) = G'(c, r)
      double s = 0.0;
      for (int c = 0; c < n; ++c) s = fma(Xm[c + (size_t)tid * NP], Amf[c], s);
      g.gv[ks * n + tid] = mf[tid] - s;
    }
  }
  __syncthreads();
  if (tid == 0 && s_fail) atomicCAS(a.status, 0, 1);
}

// ---------------------------------------------------------------------------------------------
// 2. chunk aggregates (as ds_compose_kernel).  Workspace: Ea | La | T1 | T2 (leading dimension n)
__global__ void __launch_bounds__(kDgThreads, 1)
dg_compose_kernel(const __grid_constant__ DsArgs g, double* __restrict__ ws, int NP) {
  const int tid = threadIdx.x, nth = blockDim.x;
  const int n = g.ekf.n;
  const size_t nn = (size_t)n * n;
  double* Ea = dg_ws(ws, NP, 0);
  double* La = dg_ws(ws, NP, 1);
  double* T1 = dg_ws(ws, NP, 2);
  double* T2 = dg_ws(ws, NP, 3);
  __shared__ double ga[kDgMaxNP], gt[kDgMaxNP];
  const long long nseg = g.seg_k1 - g.seg_k0;
  const long long nchunks = (nseg + g.chunk_len - 1) / g.chunk_len;
  for (long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const long long s0 = ch * g.chunk_len, s1 = min(s0 + (long long)g.chunk_len, nseg);
    __syncthreads();
    {
      // the last step of the chunk initialises the aggregate: Ea = G, La = L, ga = g
      const long long s = s1 - 1;
      const double* Gt = g.Gt + (size_t)s * nn;
      const double* Lg = g.ekf.PS + (size_t)(g.seg_k0 + s) * g.ekf.ps_stride;
      for (int i = tid; i < n * n; i += nth) { const int r = i % n, c = i / n; Ea[i] = Gt[c + (size_t)r * n]; La[i] = Lg[i]; }
      for (int i = tid; i < n; i += nth) ga[i] = g.gv[s * n + i];
    }
    for (long long s = s1 - 2; s >= s0; --s) {
      __syncthreads();
      const double* Gt = g.Gt + (size_t)s * nn;
      const double* Lg = g.ekf.PS + (size_t)(g.seg_k0 + s) * g.ekf.ps_stride;
      dg_gemm<true, false>(Gt, n, La, n, n, [&](int r, int c, double v) { T1[r + (size_t)c * n] = v; });      // G La
      dg_gemm<true, false>(Gt, n, Ea, n, n, [&](int r, int c, double v) { T2[r + (size_t)c * n] = v; });      // G Ea
      if (tid < n) {
        double s_ = g.gv[s * n + tid];
        for (int c = 0; c < n; ++c) s_ = fma(Gt[c + (size_t)tid * n], ga[c], s_);
        gt[tid] = s_;
      }
      __syncthreads();
      dg_gemm<false, false>(T1, n, Gt, n, n, [&](int r, int c, double v) { La[r + (size_t)c * n] = v + Lg[r + (size_t)c * n]; });   // T1 G' + L_k
      if (tid < n) ga[tid] = gt[tid];
      double* tmp = Ea; Ea = T2; T2 = tmp;
    }
    __syncthreads();
    double* aE = g.aggE + (size_t)ch * nn;
    double* aL = g.aggL + (size_t)ch * nn;
    for (int i = tid; i < n * n; i += nth) { aE[i] = Ea[i]; aL[i] = La[i]; }
    for (int i = tid; i < n; i += nth) g.aggg[ch * n + i] = ga[i];
  }
}

// H m and diag(H P H') of the smoothed estimate at step k (Ps: leading dimension n); maxDiffP (:250)
__device__ __forceinline__ void dg_emit(const DsArgs& g, long long k, const double* ms, const double* Ps, double& md) {
  const EkfArgs& a = g.ekf;
  const int tid = threadIdx.x, n = a.n;
  if (tid < a.M) {
    const int o = a.off[tid], nb = a.off[tid + 1] - o;
    const double* h = a.h + tid * a.BM;
    double e = 0.0, v = 0.0;
    for (int c = 0; c < nb; ++c) {
      e = fma(h[c], ms[o + c], e);
      double hp = 0.0;
      for (int l = 0; l < nb; ++l) hp = fma(h[l], Ps[(o + l) + (size_t)(o + c) * n], hp);
      v = fma(hp, h[c], v);
    }
    double* ev = g.EV + (size_t)k * 2 * a.M;
    md = fmax(md, fabs(ev[a.M + tid] - v));
    ev[tid] = e; ev[a.M + tid] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// 3. carry across the chunks of the segment (one CTA).  Workspace: Ps | T1
__global__ void __launch_bounds__(kDgThreads, 1)
dg_carry_kernel(const __grid_constant__ DsArgs g, double* __restrict__ ws, int NP, int emit_last) {
  const int tid = threadIdx.x, nth = blockDim.x;
  const int n = g.ekf.n;
  const size_t nn = (size_t)n * n;
  double* Ps = dg_ws(ws, NP, 0);
  double* T1 = dg_ws(ws, NP, 1);
  __shared__ double ms[kDgMaxNP], mt[kDgMaxNP];
  for (int i = tid; i < n * n; i += nth) Ps[i] = g.carryP[i];
  for (int i = tid; i < n; i += nth) ms[i] = g.carrym[i];
  __syncthreads();
  double md = 0.0;
  if (emit_last) dg_emit(g, g.ekf.T - 1, ms, Ps, md);          // step T is its own smoothed estimate
  const long long nseg = g.seg_k1 - g.seg_k0;
  const long long nchunks = (nseg + g.chunk_len - 1) / g.chunk_len;
  for (long long ch = nchunks - 1; ch >= 0; --ch) {
    __syncthreads();
    double* eP = g.entP + (size_t)ch * nn;
    for (int i = tid; i < n * n; i += nth) eP[i] = Ps[i];
    for (int i = tid; i < n; i += nth) g.entm[ch * n + i] = ms[i];
    const double* Em = g.aggE + (size_t)ch * nn;
    const double* Lg = g.aggL + (size_t)ch * nn;
    dg_gemm<false, false>(Em, n, Ps, n, n, [&](int r, int c, double v) { T1[r + (size_t)c * n] = v; });
    if (tid < n) {
      double s = g.aggg[ch * n + tid];
      for (int c = 0; c < n; ++c) s = fma(Em[tid + (size_t)c * n], ms[c], s);
      mt[tid] = s;
    }
    __syncthreads();
    dg_gemm<false, true>(T1, n, Em, n, n, [&](int r, int c, double v) { Ps[r + (size_t)c * n] = v + Lg[r + (size_t)c * n]; });
    if (tid < n) ms[tid] = mt[tid];
  }
  __syncthreads();
  for (int i = tid; i < n * n; i += nth) g.carryP[i] = Ps[i];
  for (int i = tid; i < n; i += nth) g.carrym[i] = ms[i];
  if (tid < g.ekf.M) atomic_max_nonneg(g.maxdiff, md);
}

// ---------------------------------------------------------------------------------------------
// 4. apply (as ds_apply_kernel): the running covariance lives in PS itself.  Workspace: T1
__global__ void __launch_bounds__(kDgThreads, 1)
dg_apply_kernel(const __grid_constant__ DsArgs g, double* __restrict__ ws, int NP) {
  const int tid = threadIdx.x, nth = blockDim.x;
  const int n = g.ekf.n;
  const size_t nn = (size_t)n * n;
  double* T1 = dg_ws(ws, NP, 0);
  __shared__ double ms[kDgMaxNP], mt[kDgMaxNP];
  const long long nseg = g.seg_k1 - g.seg_k0;
  const long long nchunks = (nseg + g.chunk_len - 1) / g.chunk_len;
  double md = 0.0;
  for (long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const long long s0 = ch * g.chunk_len, s1 = min(s0 + (long long)g.chunk_len, nseg);
    __syncthreads();
    const double* Pcur = g.entP + (size_t)ch * nn;
    for (int i = tid; i < n; i += nth) ms[i] = g.entm[ch * n + i];
    for (long long s = s1 - 1; s >= s0; --s) {
      const long long k = g.seg_k0 + s;
      __syncthreads();
      const double* Gt = g.Gt + (size_t)s * nn;
      dg_gemm<true, false>(Gt, n, Pcur, n, n, [&](int r, int c, double v) { T1[r + (size_t)c * n] = v; });     // G P
      if (tid < n) {
        double s_ = g.gv[s * n + tid];
        for (int c = 0; c < n; ++c) s_ = fma(Gt[c + (size_t)tid * n], ms[c], s_);
        mt[tid] = s_;
      }
      __syncthreads();
      double* Pg = g.ekf.PS + (size_t)k * g.ekf.ps_stride;        // holds L_k, receives the smoothed covariance (:249)
      dg_gemm<false, false>(T1, n, Gt, n, n, [&](int r, int c, double v) { Pg[r + (size_t)c * n] += v; });       // T1 G' + L_k
      if (tid < n) { ms[tid] = mt[tid]; g.ekf.MS[k * n + tid] = mt[tid]; }
      __syncthreads();
      dg_emit(g, k, ms, Pg, md);
      Pcur = Pg;
    }
  }
  if (tid < g.ekf.M) atomic_max_nonneg(g.maxdiff, md);
}

}  // namespace nsagp
