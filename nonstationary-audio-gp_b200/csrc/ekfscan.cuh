// Dense Rauch-Tung-Striebel smoother of the iterated-EKF variant as a parallel scan over time
// (gf_giekf_modulator_nmf_constraints.m:221-253), with the dense matrix products on the FP64
// tensor cores (mma.sync m8n8k4 f64, "DMMA").
//
// The reference's backward step  m <- MS_k + G_k (m - A MS_k),  P <- PS_k + G_k (P - PSkp) G_k'
// is the affine / congruence map of the Sarkka & Garcia-Fernandez (2021) smoothing element
//     (E_k, g_k, L_k) = (G_k,  MS_k - G_k A MS_k,  PS_k - G_k PSkp G_k'),
//     x_k = E_k x_{k+1} + g_k,      P_k = E_k P_{k+1} E_k' + L_k,
// whose composition (E_i E_j, E_i g_j + g_i, E_i L_j E_i' + L_i) is associative.  Four kernels
// per segment of time steps, processed from the end of the signal towards its start:
//   1. ds_elements_kernel : one CTA per time step, all steps in parallel: PSkp = A PS_k A' + Q,
//      blocked Cholesky (8-wide panels, look-ahead), Y = C^-1 (A PS_k), L = PS_k - Y'Y,
//      G' = C^-T Y, g.  L_k overwrites PS_k in HBM, G_k' goes to a scratch buffer.
//   2. ds_compose_kernel  : one CTA per chunk of `chunk_len` steps composes the chunk's elements.
//   3. ds_carry_kernel    : one CTA walks the chunk aggregates of the segment (the only sequential
//      part: n_chunks applications) and records the state entering every chunk.
//   4. ds_apply_kernel    : one CTA per chunk re-applies the steps from the entering state and
//      writes MS_k, PS_k, H m, diag(H P H').
// Only the association of the products differs from the reference (tolerance class 1e-6;
// measured ~1e-12).  Matrices live in shared memory column-major with leading dimension
// NP + 4 (== 4 mod 16), which makes every DMMA fragment load (plain or transposed) bank-conflict free.
#pragma once
#include "common.cuh"
#include "ekf.cuh"
#include "fastmath.cuh"

namespace nsagp {

constexpr int kDsThreads = 640;     // 20 warps = the 20 (5 tile rows x 1 tile column) units of an 80 x 80 product: 5 per scheduler
constexpr int kDsCT = 1;            // tile columns per warp unit
constexpr int kDsDinvLd = 12;       // leading dimension of the 8 x 8 inverse diagonal blocks (conflict-free fragments)

template <int NP>
struct Ds {
  static constexpr int LD = NP + 4;
  static constexpr int NT = NP / 8;          // 8 x 8 tiles per dimension
  static constexpr int RT = NT / 2;          // tile rows per warp unit
  static constexpr int MAT = LD * NP;        // doubles per matrix
};

struct DsArgs {
  EkfArgs ekf;                 // model, MS, PS
  double* Gt;                  // [seg steps][n*n]   G_k' (column-major)
  double* gv;                  // [seg steps][n]
  double* aggE;                // [seg chunks][n*n]
  double* aggL;                // [seg chunks][n*n]
  double* aggg;                // [seg chunks][n]
  double* entP;                // [seg chunks][n*n]  smoothed covariance entering the chunk (state at its end + 1)
  double* entm;                // [seg chunks][n]
  double* carryP;              // [n*n] state at the right boundary of the segment being processed
  double* carrym;              // [n]
  double* EV;                  // [T][2][M]
  unsigned long long* maxdiff;
  long long seg_k0, seg_k1;    // element indices [seg_k0, seg_k1) of this segment (element k maps state k+1 -> k)
  int chunk_len;
};

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// C = op(X) * op(Y) over k tiles [kt0, kt1) (units of 4), all in shared memory; epi(r, c, v0, v1) receives
// C(r, c) and C(r, c+1).  TX: op(X) = X'.  TY: op(Y) = Y'.  Warp w < 2*NT/kDsCT owns tile rows (w&1)*RT.., tile columns
// kDsCT*(w>>1)..  (DMMA issues at one per 16 cycles per scheduler: the warps must spread evenly over the four schedulers).
template <int NP, bool TX, bool TY, class Epi>
__device__ __forceinline__ void ds_gemm(const double* __restrict__ X, const double* __restrict__ Y, int kt0, int kt1, Epi epi) {
  constexpr int LD = Ds<NP>::LD, NT = Ds<NP>::NT, RT = Ds<NP>::RT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gr = lane >> 2, q = lane & 3;
  constexpr int CT = kDsCT;
  if (warp >= 2 * NT / CT) return;
  const int tr0 = (warp & 1) * RT, tc0 = (warp >> 1) * CT;
  double acc[RT][CT][2];
#pragma unroll
  for (int i = 0; i < RT; ++i)
#pragma unroll
    for (int j = 0; j < CT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll 4
  for (int kt = kt0; kt < kt1; ++kt) {
    const int k = kt * 4 + q;
    double a[RT], b[CT];
#pragma unroll
    for (int i = 0; i < RT; ++i) {
      const int r = (tr0 + i) * 8 + gr;
      a[i] = TX ? X[k + r * LD] : X[r + k * LD];
    }
#pragma unroll
    for (int j = 0; j < CT; ++j) {
      const int c = (tc0 + j) * 8 + gr;
      b[j] = TY ? Y[c + k * LD] : Y[k + c * LD];
    }
#pragma unroll
    for (int i = 0; i < RT; ++i)
#pragma unroll
      for (int j = 0; j < CT; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
  }
#pragma unroll
  for (int i = 0; i < RT; ++i)
#pragma unroll
    for (int j = 0; j < CT; ++j) epi((tr0 + i) * 8 + gr, (tc0 + j) * 8 + 2 * q, acc[i][j][0], acc[i][j][1]);
}

// Pull the n*n doubles at g towards L2 (one 128-byte line per thread) ahead of the step that reads them.
__device__ __forceinline__ void ds_prefetch_l2(const double* g, int n) {
  const size_t bytes = (size_t)n * n * 8;
  for (size_t o = (size_t)threadIdx.x * 128; o < bytes; o += (size_t)blockDim.x * 128)
    asm volatile("prefetch.global.L2 [%0];" ::"l"((const char*)g + o));
}

template <int NP>
__device__ __forceinline__ void ds_load(double* S, const double* __restrict__ g, int n) {
  constexpr int LD = Ds<NP>::LD;
  for (int i = threadIdx.x; i < NP * NP; i += blockDim.x) {
    const int r = i % NP, c = i / NP;
    S[r + c * LD] = (r < n && c < n) ? g[r + (size_t)c * n] : 0.0;
  }
}
template <int NP>
__device__ __forceinline__ void ds_store(double* __restrict__ g, const double* S, int n) {
  constexpr int LD = Ds<NP>::LD;
  for (int i = threadIdx.x; i < n * n; i += blockDim.x) {
    const int r = i % n, c = i / n;
    g[i] = S[r + c * LD];
  }
}

// v_out[r] = base[r] +/- sum_c op(Mx)(r, c) v[c],  TR: op = transpose (Mx holds G', the product is with G).
// 4 threads per row (kDsThreads = 4 * 80); v_out may not alias v.
template <int NP, bool TR>
__device__ __forceinline__ void ds_matvec(const double* Mx, const double* v, const double* base, double* v_out, int n,
                                          bool subtract = false) {
  constexpr int LD = Ds<NP>::LD;
  const int r = threadIdx.x >> 2, part = threadIdx.x & 3;
  double s = 0.0;
  if (r < NP)
    for (int c = part; c < NP; c += 4) s = fma(TR ? Mx[c + r * LD] : Mx[r + c * LD], v[c], s);
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  if (part == 0 && r < n) v_out[r] = subtract ? base[r] - s : base[r] + s;
}

// Cholesky factor of the 8 x 8 diagonal tile at Cm(8p.., 8p..) and its inverse (lower triangular) -> dinv[8][12].
// One thread, everything in registers.  Returns false if a pivot is not positive.
template <int NP>
__device__ __forceinline__ bool ds_chol8_inv(const double* Cm, int p, double* dinv) {
  constexpr int LD = Ds<NP>::LD;
  double a[8][8];
  const double* t = Cm + (8 * p) + (size_t)(8 * p) * LD;
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int r = 0; r < 8; ++r) a[r][c] = (r >= c) ? t[r + c * LD] : 0.0;
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
  // inverse of the unit-scaled factor: Linv(r,c) = -rs[r] * sum_{k=c}^{r-1} L(r,k) Linv(k,c),  Linv(c,c) = rs[c]
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
    for (int r = 0; r < 8; ++r) dinv[r + c * kDsDinvLd] = li[r][c];
  return ok;
}

// ---------------------------------------------------------------------------------------------
// 1. elements
// shared: Pm | Xm | Cm | Dinv[NT][8*12] | sA | sQ | mf[NP] | Amf[NP] | gk[NP]
template <int NP>
__global__ void __launch_bounds__(kDsThreads, 1)
ds_elements_kernel(const __grid_constant__ DsArgs g) {
  constexpr int LD = Ds<NP>::LD, NT = Ds<NP>::NT, MAT = Ds<NP>::MAT;
  const EkfArgs& a = g.ekf;
  const int tid = threadIdx.x, nth = blockDim.x, warp = tid >> 5, lane = tid & 31, gr = lane >> 2, q = lane & 3;
  const int nwarps = nth >> 5;
  const int M = a.M, n = a.n, BM = a.BM;
  extern __shared__ double sm[];
  double* Pm = sm;
  double* Xm = Pm + MAT;
  double* Cm = Xm + MAT;
  double* Dinv = Cm + MAT;
  double* sA = Dinv + NT * 8 * kDsDinvLd;
  double* sQ = sA + M * BM * BM;
  double* mf = sQ + M * BM * BM;
  double* Amf = mf + NP;
  double* gk = Amf + NP;
  __shared__ int s_blk[96];
  __shared__ int s_fail;
  for (int i = tid; i < M * BM * BM; i += nth) { sA[i] = a.A[i]; sQ[i] = a.Q[i]; }
  for (int b = tid; b < M; b += nth)
    for (int i = a.off[b]; i < a.off[b + 1]; ++i) s_blk[i] = b;
  if (tid == 0) s_fail = 0;
  const size_t nn = (size_t)n * n;

  for (long long k = g.seg_k0 + blockIdx.x; k < g.seg_k1; k += gridDim.x) {
    __syncthreads();
    ds_load<NP>(Pm, a.PS + (size_t)k * a.ps_stride, n);
    if (k + gridDim.x < g.seg_k1) ds_prefetch_l2(a.PS + (size_t)(k + gridDim.x) * a.ps_stride, n);
    for (int i = tid; i < NP; i += nth) mf[i] = (i < n) ? a.MS[k * n + i] : 0.0;
    // padding of X (zero) and of PSkp (identity)
    for (int i = tid; i < NP * NP; i += nth) {
      const int r = i % NP, c = i / NP;
      if (r >= n || c >= n) { Xm[r + c * LD] = 0.0; Cm[r + c * LD] = (r == c) ? 1.0 : 0.0; }
    }
    __syncthreads();
    // X = A PS_k  (A block diagonal: row r of X mixes the rows of PS_k in r's block)
    for (int i = tid; i < n * n; i += nth) {
      const int r = i % n, c = i / n, b = s_blk[r], o = a.off[b], nb = a.off[b + 1] - o;
      double s = 0.0;
      for (int l = 0; l < nb; ++l) s = fma(sA[b * BM * BM + (r - o) + l * BM], Pm[(o + l) + c * LD], s);
      Xm[r + c * LD] = s;
    }
    __syncthreads();
    // PSkp = X A' + Q  (:229)
    for (int i = tid; i < n * n; i += nth) {
      const int r = i % n, c = i / n, b = s_blk[c], o = a.off[b], nb = a.off[b + 1] - o;
      double s = (s_blk[r] == b) ? sQ[b * BM * BM + (r - o) + (c - o) * BM] : 0.0;
      for (int l = 0; l < nb; ++l) s = fma(Xm[r + (o + l) * LD], sA[b * BM * BM + (c - o) + l * BM], s);
      Cm[r + c * LD] = s;
    }
    if (tid < n) {                                   // A * MS(:,k)
      const int b = s_blk[tid], o = a.off[b], nb = a.off[b + 1] - o;
      double s = 0.0;
      for (int c = 0; c < nb; ++c) s = fma(sA[b * BM * BM + (tid - o) + c * BM], mf[o + c], s);
      Amf[tid] = s;
    } else if (tid < NP) Amf[tid] = 0.0;
    __syncthreads();

    // Blocked Cholesky of PSkp (lower, :232) with the forward solve Y = C^-1 X folded into the trailing updates.
    if (tid == 0 && !ds_chol8_inv<NP>(Cm, 0, Dinv)) s_fail = 1;
    for (int p = 0; p < NT; ++p) {
      __syncthreads();
      const double* Dp = Dinv + p * 8 * kDsDinvLd;
      // panel: C_ip = P_ip Dp' (i > p);  Y_pi = Dp X_pi (all i)
      for (int t = warp; t < (NT - 1 - p) + NT; t += nwarps) {
        double c0 = 0.0, c1 = 0.0;
        if (t < NT - 1 - p) {
          const int i = p + 1 + t;
#pragma unroll
          for (int kt = 0; kt < 2; ++kt) {
            const int kk = kt * 4 + q;
            dmma884(c0, c1, Cm[(8 * i + gr) + (8 * p + kk) * LD], Dp[gr + kk * kDsDinvLd]);
          }
          Cm[(8 * i + gr) + (8 * p + 2 * q) * LD] = c0;
          Cm[(8 * i + gr) + (8 * p + 2 * q + 1) * LD] = c1;
        } else {
          const int i = t - (NT - 1 - p);
#pragma unroll
          for (int kt = 0; kt < 2; ++kt) {
            const int kk = kt * 4 + q;
            dmma884(c0, c1, Dp[gr + kk * kDsDinvLd], Xm[(8 * p + kk) + (8 * i + gr) * LD]);
          }
          Xm[(8 * p + gr) + (8 * i + 2 * q) * LD] = c0;
          Xm[(8 * p + gr) + (8 * i + 2 * q + 1) * LD] = c1;
        }
      }
      __syncthreads();
      if (p + 1 == NT) break;
      // trailing update.  Tasks: lower tiles (i, j), p < j <= i, then X tiles (qq, i), qq > p.
      // Warp 0 takes tile (p+1, p+1) first and factors it while the others finish (look-ahead).
      const int nrow = NT - 1 - p;
      const int nlow = nrow * (nrow + 1) / 2, ntask = nlow + nrow * NT;
      auto do_task = [&](int t) {
        if (t < nlow) {
          // t -> (i, j) in the lower triangle, row-major enumeration; t = 0 is (p+1, p+1)
          int ii = 0;
          while ((ii + 1) * (ii + 2) / 2 <= t) ++ii;
          const int jj = t - ii * (ii + 1) / 2;
          const int i = p + 1 + ii, j = p + 1 + jj;
          double c0 = Cm[(8 * i + gr) + (8 * j + 2 * q) * LD], c1 = Cm[(8 * i + gr) + (8 * j + 2 * q + 1) * LD];
#pragma unroll
          for (int kt = 0; kt < 2; ++kt) {
            const int kk = 8 * p + kt * 4 + q;
            dmma884(c0, c1, -Cm[(8 * i + gr) + kk * LD], Cm[(8 * j + gr) + kk * LD]);
          }
          Cm[(8 * i + gr) + (8 * j + 2 * q) * LD] = c0;
          Cm[(8 * i + gr) + (8 * j + 2 * q + 1) * LD] = c1;
        } else {
          const int u = t - nlow;
          const int qq = p + 1 + u / NT, i = u % NT;
          double c0 = Xm[(8 * qq + gr) + (8 * i + 2 * q) * LD], c1 = Xm[(8 * qq + gr) + (8 * i + 2 * q + 1) * LD];
#pragma unroll
          for (int kt = 0; kt < 2; ++kt) {
            const int kk = 8 * p + kt * 4 + q;
            dmma884(c0, c1, -Cm[(8 * qq + gr) + kk * LD], Xm[kk + (8 * i + gr) * LD]);
          }
          Xm[(8 * qq + gr) + (8 * i + 2 * q) * LD] = c0;
          Xm[(8 * qq + gr) + (8 * i + 2 * q + 1) * LD] = c1;
        }
      };
      if (warp == 0) {
        do_task(0);
        __syncwarp();
        if (lane == 0 && !ds_chol8_inv<NP>(Cm, p + 1, Dinv + (p + 1) * 8 * kDsDinvLd)) s_fail = 1;
      } else {
        for (int t = warp; t < ntask; t += nwarps - 1) do_task(t);
      }
    }
    // here: Cm lower tiles = C (diagonal tiles represented by Dinv), Xm = Y = C^-1 A PS_k.
    // L = PS_k - Y'Y  (= PS_k - G PSkp G'), written over PS_k in HBM.
    {
      double* Lg = a.PS + (size_t)k * a.ps_stride;
      ds_gemm<NP, true, false>(Xm, Xm, 0, NP / 4, [&](int r, int c, double v0, double v1) {
        if (r < n) {
          if (c < n) Lg[r + (size_t)c * n] = Pm[r + c * LD] - v0;
          if (c + 1 < n) Lg[r + (size_t)(c + 1) * n] = Pm[r + (c + 1) * LD] - v1;
        }
      });
    }
    __syncthreads();
    // G' = C^-T Y, left-looking, in place; warp i owns the 8 columns of tile column i (no CTA barrier needed).
    if (warp < NT) {
      const int i = warp;
      for (int qq = NT - 1; qq >= 0; --qq) {
        double c0 = Xm[(8 * qq + gr) + (8 * i + 2 * q) * LD], c1 = Xm[(8 * qq + gr) + (8 * i + 2 * q + 1) * LD];
        for (int p = qq + 1; p < NT; ++p) {
#pragma unroll
          for (int kt = 0; kt < 2; ++kt) {
            const int kk = 8 * p + kt * 4 + q;
            dmma884(c0, c1, -Cm[kk + (8 * qq + gr) * LD], Xm[kk + (8 * i + gr) * LD]);
          }
        }
        __syncwarp();
        Xm[(8 * qq + gr) + (8 * i + 2 * q) * LD] = c0;
        Xm[(8 * qq + gr) + (8 * i + 2 * q + 1) * LD] = c1;
        __syncwarp();
        const double* Dq = Dinv + qq * 8 * kDsDinvLd;
        double d0 = 0.0, d1 = 0.0;
#pragma unroll
        for (int kt = 0; kt < 2; ++kt) {
          const int kk = kt * 4 + q;
          dmma884(d0, d1, Dq[kk + gr * kDsDinvLd], Xm[(8 * qq + kk) + (8 * i + gr) * LD]);
        }
        __syncwarp();
        Xm[(8 * qq + gr) + (8 * i + 2 * q) * LD] = d0;
        Xm[(8 * qq + gr) + (8 * i + 2 * q + 1) * LD] = d1;
        __syncwarp();
      }
    }
    __syncthreads();
    const long long ks = k - g.seg_k0;
    ds_store<NP>(g.Gt + (size_t)ks * nn, Xm, n);
    ds_matvec<NP, true>(Xm, Amf, mf, gk, n, true);  // g = MS_k - G A MS_k
    __syncthreads();
    if (tid < n) g.gv[ks * n + tid] = gk[tid];
  }
  __syncthreads();
  if (tid == 0 && s_fail) atomicCAS(a.status, 0, 1);
}

// ---------------------------------------------------------------------------------------------
// 2. chunk aggregates.  shared: Gm (G_k') | Ea | La | T1 | ga[NP] | gt[NP] | gk[NP]
template <int NP>
__global__ void __launch_bounds__(kDsThreads, 1)
ds_compose_kernel(const __grid_constant__ DsArgs g) {
  constexpr int LD = Ds<NP>::LD, MAT = Ds<NP>::MAT;
  const int tid = threadIdx.x, nth = blockDim.x;
  const int n = g.ekf.n;
  const size_t nn = (size_t)n * n;
  extern __shared__ double sm[];
  double* Gm = sm;
  double* Ea = Gm + MAT;
  double* La = Ea + MAT;
  double* T1 = La + MAT;
  double* ga = T1 + MAT;
  double* gt = ga + NP;
  double* gk = gt + NP;
  const long long nseg = g.seg_k1 - g.seg_k0;
  const long long nchunks = (nseg + g.chunk_len - 1) / g.chunk_len;
  for (long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const long long s0 = ch * g.chunk_len, s1 = min(s0 + (long long)g.chunk_len, nseg);    // segment-relative steps
    __syncthreads();
    // last step of the chunk initialises the aggregate: Ea = G, La = L, ga = g
    {
      const long long s = s1 - 1;
      ds_load<NP>(Gm, g.Gt + (size_t)s * nn, n);
      ds_load<NP>(La, g.ekf.PS + (size_t)(g.seg_k0 + s) * g.ekf.ps_stride, n);
      for (int i = tid; i < NP; i += nth) ga[i] = (i < n) ? g.gv[s * n + i] : 0.0;
      __syncthreads();
      for (int i = tid; i < NP * NP; i += nth) { const int r = i % NP, c = i / NP; Ea[r + c * LD] = Gm[c + r * LD]; }
    }
    for (long long s = s1 - 2; s >= s0; --s) {
      __syncthreads();
      ds_load<NP>(Gm, g.Gt + (size_t)s * nn, n);
      for (int i = tid; i < NP; i += nth) gk[i] = (i < n) ? g.gv[s * n + i] : 0.0;
      if (s > s0) {
        ds_prefetch_l2(g.Gt + (size_t)(s - 1) * nn, n);
        ds_prefetch_l2(g.ekf.PS + (size_t)(g.seg_k0 + s - 1) * g.ekf.ps_stride, n);
      }
      __syncthreads();
      // T1 = G La
      ds_gemm<NP, true, false>(Gm, La, 0, NP / 4, [&](int r, int c, double v0, double v1) {
        T1[r + c * LD] = v0; T1[r + (c + 1) * LD] = v1;
      });
      ds_matvec<NP, true>(Gm, ga, gk, gt, n);
      __syncthreads();
      // La = T1 G' + L_k
      {
        const double* Lg = g.ekf.PS + (size_t)(g.seg_k0 + s) * g.ekf.ps_stride;
        ds_gemm<NP, false, false>(T1, Gm, 0, NP / 4, [&](int r, int c, double v0, double v1) {
          const bool rr = r < n;
          La[r + c * LD] = (rr && c < n) ? v0 + Lg[r + (size_t)c * n] : 0.0;
          La[r + (c + 1) * LD] = (rr && c + 1 < n) ? v1 + Lg[r + (size_t)(c + 1) * n] : 0.0;
        });
      }
      for (int i = tid; i < NP; i += nth) ga[i] = (i < n) ? gt[i] : 0.0;
      __syncthreads();
      // Ea = G Ea  (through T1)
      ds_gemm<NP, true, false>(Gm, Ea, 0, NP / 4, [&](int r, int c, double v0, double v1) {
        T1[r + c * LD] = v0; T1[r + (c + 1) * LD] = v1;
      });
      __syncthreads();
      double* tmp = Ea; Ea = T1; T1 = tmp;
    }
    __syncthreads();
    ds_store<NP>(g.aggE + (size_t)ch * nn, Ea, n);
    ds_store<NP>(g.aggL + (size_t)ch * nn, La, n);
    for (int i = tid; i < n; i += nth) g.aggg[ch * n + i] = ga[i];
  }
}

// H m and diag(H P H') of the smoothed estimate at step k; maxDiffP against the previous global iteration (:250).
template <int NP>
__device__ __forceinline__ void ds_emit(const DsArgs& g, long long k, const double* ms, const double* Ps, double& md) {
  constexpr int LD = Ds<NP>::LD;
  const EkfArgs& a = g.ekf;
  const int tid = threadIdx.x;
  if (tid < a.M) {
    const int o = a.off[tid], nb = a.off[tid + 1] - o;
    const double* h = a.h + tid * a.BM;
    double e = 0.0, v = 0.0;
    for (int c = 0; c < nb; ++c) {
      e = fma(h[c], ms[o + c], e);
      double hp = 0.0;
      for (int l = 0; l < nb; ++l) hp = fma(h[l], Ps[(o + l) + (o + c) * LD], hp);
      v = fma(hp, h[c], v);
    }
    double* ev = g.EV + (size_t)k * 2 * a.M;
    md = fmax(md, fabs(ev[a.M + tid] - v));
    ev[tid] = e; ev[a.M + tid] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// 3. carry across the chunks of the segment (one CTA).  shared: Em | Ps | T1 | ms[NP] | mt[NP] | gk[NP]
template <int NP>
__global__ void __launch_bounds__(kDsThreads, 1)
ds_carry_kernel(const __grid_constant__ DsArgs g, int emit_last) {
  constexpr int LD = Ds<NP>::LD, MAT = Ds<NP>::MAT;
  const int tid = threadIdx.x, nth = blockDim.x;
  const int n = g.ekf.n;
  const size_t nn = (size_t)n * n;
  extern __shared__ double sm[];
  double* Em = sm;
  double* Ps = Em + MAT;
  double* T1 = Ps + MAT;
  double* ms = T1 + MAT;
  double* mt = ms + NP;
  double* gk = mt + NP;
  ds_load<NP>(Ps, g.carryP, n);
  for (int i = tid; i < NP; i += nth) ms[i] = (i < n) ? g.carrym[i] : 0.0;
  __syncthreads();
  double md = 0.0;
  if (emit_last) ds_emit<NP>(g, g.ekf.T - 1, ms, Ps, md);      // step T is its own smoothed estimate
  const long long nseg = g.seg_k1 - g.seg_k0;
  const long long nchunks = (nseg + g.chunk_len - 1) / g.chunk_len;
  for (long long ch = nchunks - 1; ch >= 0; --ch) {
    __syncthreads();
    ds_store<NP>(g.entP + (size_t)ch * nn, Ps, n);
    for (int i = tid; i < n; i += nth) g.entm[ch * n + i] = ms[i];
    ds_load<NP>(Em, g.aggE + (size_t)ch * nn, n);
    for (int i = tid; i < NP; i += nth) gk[i] = (i < n) ? g.aggg[ch * n + i] : 0.0;
    __syncthreads();
    ds_gemm<NP, false, false>(Em, Ps, 0, NP / 4, [&](int r, int c, double v0, double v1) {
      T1[r + c * LD] = v0; T1[r + (c + 1) * LD] = v1;
    });
    ds_matvec<NP, false>(Em, ms, gk, mt, n);
    __syncthreads();
    {
      const double* Lg = g.aggL + (size_t)ch * nn;
      ds_gemm<NP, false, true>(T1, Em, 0, NP / 4, [&](int r, int c, double v0, double v1) {
        const bool rr = r < n;
        Ps[r + c * LD] = (rr && c < n) ? v0 + Lg[r + (size_t)c * n] : 0.0;
        Ps[r + (c + 1) * LD] = (rr && c + 1 < n) ? v1 + Lg[r + (size_t)(c + 1) * n] : 0.0;
      });
    }
    for (int i = tid; i < NP; i += nth) ms[i] = (i < n) ? mt[i] : 0.0;
  }
  __syncthreads();
  ds_store<NP>(g.carryP, Ps, n);
  for (int i = tid; i < n; i += nth) g.carrym[i] = ms[i];
  if (tid < g.ekf.M) atomic_max_nonneg(g.maxdiff, md);
}

// ---------------------------------------------------------------------------------------------
// 4. apply.  shared: Gm | Ps | T1 | ms[NP] | mt[NP] | gk[NP]
template <int NP>
__global__ void __launch_bounds__(kDsThreads, 1)
ds_apply_kernel(const __grid_constant__ DsArgs g) {
  constexpr int LD = Ds<NP>::LD, MAT = Ds<NP>::MAT;
  const int tid = threadIdx.x, nth = blockDim.x;
  const int n = g.ekf.n;
  const size_t nn = (size_t)n * n;
  extern __shared__ double sm[];
  double* Gm = sm;
  double* Ps = Gm + MAT;
  double* T1 = Ps + MAT;
  double* ms = T1 + MAT;
  double* mt = ms + NP;
  double* gk = mt + NP;
  const long long nseg = g.seg_k1 - g.seg_k0;
  const long long nchunks = (nseg + g.chunk_len - 1) / g.chunk_len;
  double md = 0.0;
  for (long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const long long s0 = ch * g.chunk_len, s1 = min(s0 + (long long)g.chunk_len, nseg);
    __syncthreads();
    ds_load<NP>(Ps, g.entP + (size_t)ch * nn, n);
    for (int i = tid; i < NP; i += nth) ms[i] = (i < n) ? g.entm[ch * n + i] : 0.0;
    for (long long s = s1 - 1; s >= s0; --s) {
      const long long k = g.seg_k0 + s;
      __syncthreads();
      ds_load<NP>(Gm, g.Gt + (size_t)s * nn, n);
      for (int i = tid; i < NP; i += nth) gk[i] = (i < n) ? g.gv[s * n + i] : 0.0;
      if (s > s0) {
        ds_prefetch_l2(g.Gt + (size_t)(s - 1) * nn, n);
        ds_prefetch_l2(g.ekf.PS + (size_t)(k - 1) * g.ekf.ps_stride, n);
      }
      __syncthreads();
      ds_gemm<NP, true, false>(Gm, Ps, 0, NP / 4, [&](int r, int c, double v0, double v1) {
        T1[r + c * LD] = v0; T1[r + (c + 1) * LD] = v1;
      });
      ds_matvec<NP, true>(Gm, ms, gk, mt, n);
      __syncthreads();
      {
        double* Pg = g.ekf.PS + (size_t)k * g.ekf.ps_stride;            // holds L_k, receives the smoothed covariance (:249)
        ds_gemm<NP, false, false>(T1, Gm, 0, NP / 4, [&](int r, int c, double v0, double v1) {
          const bool rr = r < n;
          double w0 = 0.0, w1 = 0.0;
          if (rr && c < n) { w0 = v0 + Pg[r + (size_t)c * n]; Pg[r + (size_t)c * n] = w0; }
          if (rr && c + 1 < n) { w1 = v1 + Pg[r + (size_t)(c + 1) * n]; Pg[r + (size_t)(c + 1) * n] = w1; }
          Ps[r + c * LD] = w0; Ps[r + (c + 1) * LD] = w1;
        });
      }
      for (int i = tid; i < NP; i += nth) ms[i] = (i < n) ? mt[i] : 0.0;
      __syncthreads();
      for (int i = tid; i < n; i += nth) g.ekf.MS[k * n + i] = ms[i];
      ds_emit<NP>(g, k, ms, Ps, md);
    }
  }
  if (tid < g.ekf.M) atomic_max_nonneg(g.maxdiff, md);
}

}  // namespace nsagp
