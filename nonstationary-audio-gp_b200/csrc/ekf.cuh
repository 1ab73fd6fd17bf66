// Globally-iterated extended Kalman filter / RTS smoother: the comparison variant of the EP
// path (gf_giekf_modulator_nmf_constraints.m:162-258 predict mode, :376-468 energy;
// iekf_update1.m:110-117).  Here the covariance really is dense: the measurement
//   y = h(x) = (H_z x)' W softplus(H_g x)                       (:490-494)
// couples all latents through its Jacobian, so P (n-by-n, n <= 137) lives in shared memory and
// one CTA owns a signal.  A and Q stay block diagonal (one block per latent), which makes the
// prediction P <- A P A' + Q an in-place update of independent (block_i, block_j) tiles.
//
// The filter is a nonlinear recurrence (sequential in time); this first version also runs the
// dense RTS pass sequentially inside the CTA (Cholesky, two triangular solves and two products
// per step, all cooperative in shared memory).
#pragma once
#include "common.cuh"
#include "fastmath.cuh"

namespace nsagp {

constexpr int kEkfThreads = 256;

struct EkfArgs {
  int D, N, M, n, BM;
  long long T;
  const int* off;                // [M+1]
  const double* A;               // [M][BM*BM] padded blocks
  const double* Q;               // [M][BM*BM]
  const double* Pinf;            // [n*n] dense, column-major
  const double* h;               // [M][BM]
  const double* W;               // [D][N] row-major
  double sigma2;
  const double* y;               // [T]
  double* MS;                    // [T][n]
  double* PS;                    // [T][n*n]
  double* m_io;                  // [n] mean carried between global iterations (:165-168)
  double* edata;                 // [1] energy (energy mode)
  int* status;
};

// P <- A P A' (+ Q on the diagonal blocks), in place, P dense n-by-n in shared memory.
// Each (i, j) pair of latent blocks is an independent b_i-by-b_j tile.
template <int BM>
__device__ __forceinline__ void ekf_predict_cov(double* P, const EkfArgs& a, const double* sA, const double* sQ) {
  const int M = a.M, n = a.n;
  for (int pair = threadIdx.x; pair < M * M; pair += blockDim.x) {
    const int bi = pair % M, bj = pair / M;
    const int oi = a.off[bi], ni = a.off[bi + 1] - oi;
    const int oj = a.off[bj], nj = a.off[bj + 1] - oj;
    const double* Ai = sA + bi * BM * BM;
    const double* Aj = sA + bj * BM * BM;
    double X[BM * BM], Tm[BM * BM];
#pragma unroll
    for (int c = 0; c < BM; ++c)
#pragma unroll
      for (int r = 0; r < BM; ++r) X[r + c * BM] = (r < ni && c < nj) ? P[(oi + r) + (size_t)(oj + c) * n] : 0.0;
#pragma unroll
    for (int c = 0; c < BM; ++c)
#pragma unroll
      for (int r = 0; r < BM; ++r) {
        double s = 0.0;
#pragma unroll
        for (int l = 0; l < BM; ++l) s = fma(Ai[r + l * BM], X[l + c * BM], s);
        Tm[r + c * BM] = s;
      }
#pragma unroll
    for (int c = 0; c < BM; ++c)
#pragma unroll
      for (int r = 0; r < BM; ++r) {
        double s = (bi == bj) ? sQ[bi * BM * BM + r + c * BM] : 0.0;
#pragma unroll
        for (int l = 0; l < BM; ++l) s = fma(Tm[r + l * BM], Aj[c + l * BM], s);
        if (r < ni && c < nj) P[(oi + r) + (size_t)(oj + c) * n] = s;
      }
  }
}

__device__ __forceinline__ double block_sum(double v, double* s_red) {
  // all threads call; returns the CTA total to every thread
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_red[w];
  return t;
}

// Filter pass (predict mode :176-206) or energy pass (:376-468, energy != 0: prediction at
// every step including the first, one linearisation, no storage).
template <int BM>
__global__ void __launch_bounds__(kEkfThreads)
giekf_filter_kernel(const EkfArgs* __restrict__ argv, int l_iter, int energy) {
  const EkfArgs& a = argv[blockIdx.x];
  const int tid = threadIdx.x, nth = blockDim.x;
  const int D = a.D, N = a.N, M = a.M, n = a.n;
  extern __shared__ double sm[];
  double* P = sm;                          // [n*n]
  double* sA = P + (size_t)n * n;          // [M*BM*BM]
  double* sQ = sA + M * BM * BM;
  double* m = sQ + M * BM * BM;            // [n]
  double* JH = m + n;                      // [n]
  double* PJ = JH + n;                     // [n]
  double* f = PJ + n;                      // [M] H*m
  double* wl = f + M;                      // [D] W*link(g)
  double* zw = wl + D;                     // [N] (z'W) .* dlink(g)
  double* s_red = zw + N;                  // [32]
  __shared__ int s_blk[160];               // state index -> latent block
  for (int i = tid; i < n * n; i += nth) P[i] = a.Pinf[i];                     // :168
  for (int i = tid; i < M * BM * BM; i += nth) { sA[i] = a.A[i]; sQ[i] = a.Q[i]; }
  for (int i = tid; i < n; i += nth) m[i] = a.m_io[i];
  for (int b = tid; b < M; b += nth)
    for (int i = a.off[b]; i < a.off[b + 1]; ++i) s_blk[i] = b;
  __syncthreads();
  double e_acc = 0.0;
  bool bad = false;

  for (long long k = 0; k < a.T; ++k) {
    if (k > 0 || energy) {                                                     // :180-183 / :392-393
      // m <- A m (block diagonal)
      double mv = 0.0;
      int row = tid;
      if (row < n) {
        const int b = s_blk[row], o = a.off[b], nb = a.off[b + 1] - o;
        for (int c = 0; c < nb; ++c) mv = fma(sA[b * BM * BM + (row - o) + c * BM], m[o + c], mv);
      }
      __syncthreads();
      if (row < n) m[row] = mv;
      ekf_predict_cov<BM>(P, a, sA, sQ);
      __syncthreads();
    }
    const double y = a.y[k];
    if (!isnan(y) || energy) {                                                 // :186
      double S = 0.0, MU = 0.0;
      for (int it = 0; it < (energy ? 1 : l_iter); ++it) {                     // iekf_update1.m:110-116
        if (tid < M) {
          const int o = a.off[tid], nb = a.off[tid + 1] - o;
          double s = 0.0;
          for (int c = 0; c < nb; ++c) s = fma(a.h[tid * BM + c], m[o + c], s);
          f[tid] = s;                                                          // H*m
        }
        __syncthreads();
        if (tid < D) {                                                         // W*linkf(g)
          double s = 0.0;
          for (int j = 0; j < N; ++j) s += a.W[tid * N + j] * log(1.0 + exp(f[D + j]));
          wl[tid] = s;
        } else if (tid >= 32 && tid < 32 + N) {                                // (z'W) .* dlinkf(g)
          const int j = tid - 32;
          double s = 0.0;
          for (int d = 0; d < D; ++d) s += f[d] * a.W[d * N + j];
          const double eg = exp(f[D + j]);
          zw[j] = s * (eg / (eg + 1.0));
        }
        __syncthreads();
        if (tid < n) {                                                         // Jacobian row (:497-503)
          const int b = s_blk[tid];
          JH[tid] = (b < D ? wl[b] : zw[b - D]) * a.h[b * BM + (tid - a.off[b])];
        }
        __syncthreads();
        double part = 0.0;
        if (tid < n) {                                                         // P*JH'
          double s = 0.0;
          for (int c = 0; c < n; ++c) s = fma(P[tid + (size_t)c * n], JH[c], s);
          PJ[tid] = s;
          part = JH[tid] * s;
        }
        double mu_part = (tid < D) ? f[tid] * wl[tid] : 0.0;
        S = a.sigma2 + block_sum(part, s_red);                                 // S = R + H P H'
        MU = block_sum(mu_part, s_red);                                        // h(m)  (:490-494)
        if (energy) {
          if (!(S > 0.0)) bad = true;                                          // :417-427 -> NaN energy
          const double v = y - MU;
          e_acc += 0.5 * log(2.0 * 3.14159265358979323846) + log(sqrt(S)) + 0.5 * v * v / S;
        }
        __syncthreads();
        if (tid < n) m[tid] = fma(PJ[tid] / S, y - MU, m[tid]);                // M = M + K (y - MU)
        __syncthreads();
      }
      // P = P - K S K'  with K, S of the last linearisation (iekf_update1.m:117)
      for (int i = tid; i < n * n; i += nth) {
        const int r = i % n, c = i / n;
        P[i] = fma(-(PJ[r] / S) * S, PJ[c] / S, P[i]);
      }
      __syncthreads();
    }
    if (!energy) {                                                             // :201-203
      for (int i = tid; i < n; i += nth) a.MS[k * n + i] = m[i];
      double* dst = a.PS + (size_t)k * n * n;
      for (int i = tid; i < n * n; i += nth) dst[i] = P[i];
    }
  }
  if (energy && tid == 0) { a.edata[0] = bad ? NAN : e_acc; if (bad || isnan(e_acc)) atomicCAS(a.status, 0, 3); }
}

// ---------------------------------------------------------------------------------------------
// Filter / energy pass, latency-oriented form (the default).  Same arithmetic as giekf_filter_kernel
// above, reorganised so that one time step costs five CTA barriers:
//   A  finish the previous step's covariance update P -= (K S) K' tile by tile, store that tile to
//      PS(:,:,k-1), and predict it in registers (A_i tile A_j' + Q) -- one pass over P instead of three.
//      With the reference's block sizes (BZ, BG > 0: exact compile-time tile shapes) only the tiles on and
//      below the block diagonal are computed and mirrored, which keeps P exactly symmetric;
//   B1 H m per latent; link and its derivative per modulator;   B2 Jacobian row (:497-503), h(m) terms;
//   C  P*JH' with four threads per row, S and h(m) reduced together;
//   D  gain, mean update, MS store.
// exp / log / reciprocal are the straight-line versions of fastmath.cuh (<= 2 ulp).
constexpr int ekf_round_bm(int b) { return b <= 2 ? 2 : b <= 3 ? 3 : b <= 4 ? 4 : b <= 6 ? 6 : 8; }
template <int BM> struct EkfF2 { static constexpr int TH = BM <= 4 ? 512 : 256; };   // register budget of the tile pass

// One (block_i, block_j) tile of exact shape NI x NJ (blocks padded to BM in sA / sQ).
template <int NI, int NJ, int BM>
__device__ __forceinline__ void ekf_tile(double* P, int n, int oi, int oj, const double* Ai, const double* Aj, const double* Qi,
                                         const double* Ks, const double* Kv, bool predict, double* store, bool mirror) {
  double X[NI * NJ];
#pragma unroll
  for (int c = 0; c < NJ; ++c)
#pragma unroll
    for (int r = 0; r < NI; ++r)
      X[r + c * NI] = fma(-Ks[oi + r], Kv[oj + c], P[(oi + r) + (oj + c) * n]);      // P - K S K' (iekf_update1.m:117)
  if (store) {
#pragma unroll
    for (int c = 0; c < NJ; ++c)
#pragma unroll
      for (int r = 0; r < NI; ++r) {
        store[(oi + r) + (size_t)(oj + c) * n] = X[r + c * NI];
        if (mirror) store[(oj + c) + (size_t)(oi + r) * n] = X[r + c * NI];
      }
  }
  if (predict) {
    double Tm[NI * NJ];
#pragma unroll
    for (int c = 0; c < NJ; ++c)
#pragma unroll
      for (int r = 0; r < NI; ++r) {
        double s = 0.0;
#pragma unroll
        for (int l = 0; l < NI; ++l) s = fma(Ai[r + l * BM], X[l + c * NI], s);
        Tm[r + c * NI] = s;
      }
#pragma unroll
    for (int c = 0; c < NJ; ++c)
#pragma unroll
      for (int r = 0; r < NI; ++r) {
        double s = Qi ? Qi[r + c * BM] : 0.0;
#pragma unroll
        for (int l = 0; l < NJ; ++l) s = fma(Tm[r + l * NI], Aj[c + l * BM], s);
        X[r + c * NI] = s;
      }
  }
#pragma unroll
  for (int c = 0; c < NJ; ++c)
#pragma unroll
    for (int r = 0; r < NI; ++r) {
      P[(oi + r) + (oj + c) * n] = X[r + c * NI];
      if (mirror) P[(oj + c) + (oi + r) * n] = X[r + c * NI];
    }
}

// BZ, BG > 0: the model's exact block sizes (reference kernels: BZ in {2,4,6,8}, BG in {1,2,3,4});
// BZ = BG = 0: any block sizes <= BM, tiles padded to BM with run-time guards.
template <int BZ, int BG, int BM>
__global__ void __launch_bounds__(EkfF2<BM>::TH, 1)
giekf_filter2_kernel(const EkfArgs* __restrict__ argv, int l_iter, int energy) {
  const EkfArgs& a = argv[blockIdx.x];
  const int tid = threadIdx.x, nth = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nth >> 5;
  const int D = a.D, N = a.N, M = a.M, n = a.n;
  const long long T = a.T;
  const double sigma2 = a.sigma2;
  const double* __restrict__ yv = a.y;
  double* __restrict__ MSg = a.MS;
  double* __restrict__ PSg = a.PS;
  extern __shared__ double sm[];
  double* P = sm;                          // [n*n]
  double* sA = P + (size_t)n * n;          // [M*BM*BM]
  double* sQ = sA + M * BM * BM;
  double* sh = sQ + M * BM * BM;           // [M*BM]
  double* sW = sh + M * BM;                // [D*N]
  double* mbuf = sW + D * N;               // [2][n]
  double* JH = mbuf + 2 * n;               // [n]
  double* PJ = JH + n;                     // [n]
  double* Kv = PJ + n;                     // [n]  K
  double* Ks = Kv + n;                     // [n]  K*S
  double* red = Ks + n;                    // [2][32]
  double* fv = red + 64;                   // [M]  H m
  double* spv = fv + M;                    // [N]  linkf(g)
  double* dlv = spv + N;                   // [N]  dlinkf(g)
  __shared__ int s_blk[160];
  __shared__ int s_off[kMaxSites + 1];
  __shared__ unsigned short s_pair[(kMaxSites * (kMaxSites + 1)) / 2];
  for (int i = tid; i < n * n; i += nth) P[i] = a.Pinf[i];                     // :168
  for (int i = tid; i < M * BM * BM; i += nth) { sA[i] = a.A[i]; sQ[i] = a.Q[i]; }
  for (int i = tid; i < M * BM; i += nth) sh[i] = a.h[i];
  for (int i = tid; i < D * N; i += nth) sW[i] = a.W[i];
  for (int i = tid; i < n; i += nth) { mbuf[i] = a.m_io[i]; Kv[i] = 0.0; Ks[i] = 0.0; }
  for (int b = tid; b <= M; b += nth) s_off[b] = a.off[b];
  for (int b = tid; b < M; b += nth)
    for (int i = a.off[b]; i < a.off[b + 1]; ++i) s_blk[i] = b;
  // pair list of the exact-shape path: [z,z lower incl. diagonal | g,z | g,g lower incl. diagonal]
  const int nzz = D * (D + 1) / 2, ngz = N * D, ngg = N * (N + 1) / 2;
  if (BZ > 0) {
    for (int p = tid; p < nzz + ngz + ngg; p += nth) {
      int bi, bj;
      if (p < nzz) { bi = 0; while ((bi + 1) * (bi + 2) / 2 <= p) ++bi; bj = p - bi * (bi + 1) / 2; }
      else if (p < nzz + ngz) { const int u = p - nzz; bi = D + u / D; bj = u % D; }
      else { const int u = p - nzz - ngz; int ii = 0; while ((ii + 1) * (ii + 2) / 2 <= u) ++ii; bi = D + ii; bj = D + u - ii * (ii + 1) / 2; }
      s_pair[p] = (unsigned short)(bi | (bj << 8));
    }
  }
  __syncthreads();
  double* m = mbuf;
  double* m2 = mbuf + n;
  double e_acc = 0.0;
  bool bad = false;
  int my_b = 0, my_o = 0, my_nb = 0;
  double my_h = 0.0;
  if (tid < n) { my_b = s_blk[tid]; my_o = s_off[my_b]; my_nb = s_off[my_b + 1] - my_o; my_h = sh[my_b * BM + (tid - my_o)]; }
  double y_next = yv[0];

  // tile pass: finish the pending update (if any), optionally store to PS(:,:,k-1), optionally predict
  auto tile_pass = [&](bool predict, double* store) {
    if (BZ > 0) {
      for (int p = tid; p < nzz + ngz + ngg; p += nth) {
        const int pr = s_pair[p], bi = pr & 255, bj = pr >> 8;
        const int oi = s_off[bi], oj = s_off[bj];
        const double* Ai = sA + bi * BM * BM;
        const double* Aj = sA + bj * BM * BM;
        const double* Qi = (bi == bj) ? sQ + bi * BM * BM : nullptr;
        if (p < nzz) ekf_tile<(BZ > 0 ? BZ : 1), (BZ > 0 ? BZ : 1), BM>(P, n, oi, oj, Ai, Aj, Qi, Ks, Kv, predict, store, bi != bj);
        else if (p < nzz + ngz) ekf_tile<(BG > 0 ? BG : 1), (BZ > 0 ? BZ : 1), BM>(P, n, oi, oj, Ai, Aj, Qi, Ks, Kv, predict, store, true);
        else ekf_tile<(BG > 0 ? BG : 1), (BG > 0 ? BG : 1), BM>(P, n, oi, oj, Ai, Aj, Qi, Ks, Kv, predict, store, bi != bj);
      }
    } else {
      for (int pair = tid; pair < M * M; pair += nth) {
        const int bi = pair % M, bj = pair / M;
        const int oi = s_off[bi], ni = s_off[bi + 1] - oi;
        const int oj = s_off[bj], nj = s_off[bj + 1] - oj;
        double X[BM * BM];
#pragma unroll
        for (int c = 0; c < BM; ++c)
#pragma unroll
          for (int r = 0; r < BM; ++r) {
            double v = 0.0;
            if (r < ni && c < nj) {
              v = fma(-Ks[oi + r], Kv[oj + c], P[(oi + r) + (size_t)(oj + c) * n]);
              if (store) store[(oi + r) + (size_t)(oj + c) * n] = v;
            }
            X[r + c * BM] = v;
          }
        if (predict) {
          const double* Ai = sA + bi * BM * BM;
          const double* Aj = sA + bj * BM * BM;
          double Tm[BM * BM];
#pragma unroll
          for (int c = 0; c < BM; ++c)
#pragma unroll
            for (int r = 0; r < BM; ++r) {
              double s2 = 0.0;
#pragma unroll
              for (int l = 0; l < BM; ++l) s2 = fma(Ai[r + l * BM], X[l + c * BM], s2);
              Tm[r + c * BM] = s2;
            }
#pragma unroll
          for (int c = 0; c < BM; ++c)
#pragma unroll
            for (int r = 0; r < BM; ++r) {
              double s2 = (bi == bj) ? sQ[bi * BM * BM + r + c * BM] : 0.0;
#pragma unroll
              for (int l = 0; l < BM; ++l) s2 = fma(Tm[r + l * BM], Aj[c + l * BM], s2);
              X[r + c * BM] = s2;
            }
        }
#pragma unroll
        for (int c = 0; c < BM; ++c)
#pragma unroll
          for (int r = 0; r < BM; ++r)
            if (r < ni && c < nj) P[(oi + r) + (size_t)(oj + c) * n] = X[r + c * BM];
      }
    }
  };

  for (long long k = 0; k < T; ++k) {
    const double y = y_next;
    if (k + 1 < T) y_next = yv[k + 1];
    // ---- A
    if (k > 0 || energy) {                                                     // :180-183 / :392-393
      if (tid < n) {
        double mv = 0.0;
        for (int c = 0; c < my_nb; ++c) mv = fma(sA[my_b * BM * BM + (tid - my_o) + c * BM], m[my_o + c], mv);
        m2[tid] = mv;
      }
      tile_pass(true, (!energy && k > 0) ? PSg + (size_t)(k - 1) * n * n : nullptr);
      double* t = m; m = m2; m2 = t;
      __syncthreads();
    }
    const bool upd = !isnan(y) || energy;                                      // :186
    double Kr = 0.0, S = 0.0;
    if (upd) {
      const int iters = energy ? 1 : l_iter;
      for (int it = 0; it < iters; ++it) {                                     // iekf_update1.m:110-116
        // ---- B1: f = H m; link values of the modulators
        if (tid < M) {
          const int o = s_off[tid], nb = s_off[tid + 1] - o;
          double f = 0.0;
          for (int c = 0; c < nb; ++c) f = fma(sh[tid * BM + c], m[o + c], f);
          fv[tid] = f;
          if (tid >= D) {
            const double eg = exp_fast(f);
            spv[tid - D] = log_ge1_fast(1.0 + eg);                             // linkf(g) = log(1+exp(g))
            dlv[tid - D] = eg * rcp_fast(eg + 1.0);                            // dlinkf(g) = exp(g)/(exp(g)+1)
          }
        }
        __syncthreads();
        // ---- B2: Jacobian row entry of this thread's state (:497-503), h(m) contributions (:490-494)
        double mu_part = 0.0;
        if (tid < n) {
          double jh;
          if (my_b < D) {
            double wl = 0.0;
            for (int j = 0; j < N; ++j) wl = fma(sW[my_b * N + j], spv[j], wl);
            jh = wl * my_h;
            if (tid == my_o) mu_part = fv[my_b] * wl;
          } else {
            const int j = my_b - D;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            int d = 0;
            for (; d + 3 < D; d += 4) {
              s0 = fma(fv[d], sW[d * N + j], s0);
              s1 = fma(fv[d + 1], sW[(d + 1) * N + j], s1);
              s2 = fma(fv[d + 2], sW[(d + 2) * N + j], s2);
              s3 = fma(fv[d + 3], sW[(d + 3) * N + j], s3);
            }
            for (; d < D; ++d) s0 = fma(fv[d], sW[d * N + j], s0);
            jh = ((s0 + s1) + (s2 + s3)) * dlv[j] * my_h;
          }
          JH[tid] = jh;
        }
        __syncthreads();
        // ---- C: P*JH' (four threads per row), S = R + JH P JH', h(m)
        double s_part = 0.0;
        for (int r0 = 0; r0 < n; r0 += nth >> 2) {                               // uniform trip count: shuffles inside
          const int r = r0 + (tid >> 2);
          double sa = 0.0, sb = 0.0;
          if (r < n) {
            int c = tid & 3;
            for (; c + 4 < n; c += 8) {
              sa = fma(P[r + (size_t)c * n], JH[c], sa);
              sb = fma(P[r + (size_t)(c + 4) * n], JH[c + 4], sb);
            }
            if (c < n) sa = fma(P[r + (size_t)c * n], JH[c], sa);
          }
          double s = sa + sb;
          s += __shfl_xor_sync(0xffffffffu, s, 1);
          s += __shfl_xor_sync(0xffffffffu, s, 2);
          if ((tid & 3) == 0 && r < n) { PJ[r] = s; s_part = fma(JH[r], s, s_part); }
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
          s_part += __shfl_xor_sync(0xffffffffu, s_part, o);
          mu_part += __shfl_xor_sync(0xffffffffu, mu_part, o);
        }
        if (lane == 0) { red[warp] = s_part; red[32 + warp] = mu_part; }
        __syncthreads();
        S = sigma2;
        double MU = 0.0;
        for (int w = 0; w < nwarps; ++w) { S += red[w]; MU += red[32 + w]; }
        if (energy) {
          if (!(S > 0.0)) bad = true;                                          // :417-427 -> NaN energy
          const double v = y - MU;
          if (tid == 0) e_acc += 0.5 * log(2.0 * 3.14159265358979323846) + log(sqrt(S)) + 0.5 * v * v / S;
        }
        // ---- D
        if (tid < n) {
          Kr = PJ[tid] / S;
          m[tid] = fma(Kr, y - MU, m[tid]);                                    // M = M + K (y - MU)
        }
        if (it + 1 < iters) __syncthreads();
      }
    }
    if (tid < n) {
      Kv[tid] = Kr; Ks[tid] = Kr * S;                                          // consumed by the next tile pass
      if (!energy) MSg[k * n + tid] = m[tid];                                  // :201
    }
    __syncthreads();
  }
  if (!energy) tile_pass(false, PSg + (size_t)(T - 1) * n * n);                // last covariance (:202)
  if (energy && tid == 0) { a.edata[0] = bad ? NAN : e_acc; if (bad || isnan(e_acc)) atomicCAS(a.status, 0, 3); }
}

// C (n-by-n) = op(X) * op(Y) helpers on shared-memory matrices, all threads cooperate.
__device__ __forceinline__ void mm_nn(const double* X, const double* Y, double* C, int n) {
  for (int i = threadIdx.x; i < n * n; i += blockDim.x) {
    const int r = i % n, c = i / n;
    double s = 0.0;
    for (int l = 0; l < n; ++l) s = fma(X[r + (size_t)l * n], Y[l + (size_t)c * n], s);
    C[i] = s;
  }
}
__device__ __forceinline__ void mm_nt(const double* X, const double* Y, double* C, int n) {   // X * Y'
  for (int i = threadIdx.x; i < n * n; i += blockDim.x) {
    const int r = i % n, c = i / n;
    double s = 0.0;
    for (int l = 0; l < n; ++l) s = fma(X[r + (size_t)l * n], Y[c + (size_t)l * n], s);
    C[i] = s;
  }
}

// RTS smoother (:221-253), sequential in time, dense.  Shared memory: five n-by-n matrices.
template <int BM>
__global__ void __launch_bounds__(kEkfThreads)
giekf_smoother_kernel(const EkfArgs* __restrict__ argv, const double* __restrict__ hv_dense, double* __restrict__ EV,
                      unsigned long long* __restrict__ maxdiff) {
  // hv_dense: [M][n] dense rows of H; EV: [T][2][M] (H m, diag(H P H')) in/out for outputs and maxDiffP
  const EkfArgs& a = argv[blockIdx.x];
  const int tid = threadIdx.x, nth = blockDim.x;
  const int M = a.M, n = a.n;
  const size_t nn = (size_t)n * n;
  extern __shared__ double sm[];
  double* Pk = sm;               // filtered P_k
  double* Pp = Pk + nn;          // PSkp, then its Cholesky factor L (lower)
  double* G = Pp + nn;           // smoother gain
  double* Ps = G + nn;           // smoothed P_{k+1} -> P_k
  double* T1 = Ps + nn;          // scratch
  double* sA = T1 + nn;          // [M*BM*BM]
  double* sQ = sA + M * BM * BM;
  double* ms = sQ + M * BM * BM; // [n] smoothed mean
  double* mk = ms + n;           // [n] filtered mean of step k
  double* dm = mk + n;           // [n]
  __shared__ int s_blk[160];
  __shared__ int s_fail;
  for (int i = tid; i < M * BM * BM; i += nth) { sA[i] = a.A[i]; sQ[i] = a.Q[i]; }
  for (int b = tid; b < M; b += nth)
    for (int i = a.off[b]; i < a.off[b + 1]; ++i) s_blk[i] = b;
  const long long T = a.T;
  for (int i = tid; i < n; i += nth) ms[i] = a.MS[(T - 1) * n + i];
  for (size_t i = tid; i < nn; i += nth) Ps[i] = a.PS[(size_t)(T - 1) * nn + i];
  if (tid == 0) s_fail = 0;
  __syncthreads();
  double md = 0.0;
  auto emit = [&](long long k) {
    // H m and diag(H P H') of the smoothed estimate; maxDiffP against the previous global iteration (:250)
    if (tid < M) {
      const double* hr = hv_dense + (size_t)tid * n;
      double e = 0.0, v = 0.0;
      for (int c = 0; c < n; ++c) {
        e = fma(hr[c], ms[c], e);
        double hp = 0.0;
        for (int l = 0; l < n; ++l) hp = fma(hr[l], Ps[l + (size_t)c * n], hp);
        v = fma(hp, hr[c], v);
      }
      double* ev = EV + (size_t)k * 2 * M;
      md = fmax(md, fabs(ev[M + tid] - v));
      ev[tid] = e; ev[M + tid] = v;
    }
  };
  emit(T - 1);
  for (long long k = T - 2; k >= 0; --k) {
    for (int i = tid; i < n; i += nth) mk[i] = a.MS[k * n + i];
    for (size_t i = tid; i < nn; i += nth) { const double v = a.PS[(size_t)k * nn + i]; Pk[i] = v; Pp[i] = v; }
    __syncthreads();
    ekf_predict_cov<BM>(Pp, a, sA, sQ);                                         // PSkp = A PSk A' + Q (:229)
    // T1 = PSk * A'  (block-diagonal A: column block j of T1 = PSk(:, block j) * A_j')
    for (int i = tid; i < n * n; i += nth) {
      const int r = i % n, c = i / n, b = s_blk[c], o = a.off[b], nb = a.off[b + 1] - o;
      double s = 0.0;
      for (int l = 0; l < nb; ++l) s = fma(Pk[r + (size_t)(o + l) * n], sA[b * BM * BM + (c - o) + l * BM], s);
      T1[i] = s;
    }
    // dm = m - A*MS(:,k)
    if (tid < n) {
      const int b = s_blk[tid], o = a.off[b], nb = a.off[b + 1] - o;
      double s = 0.0;
      for (int c = 0; c < nb; ++c) s = fma(sA[b * BM * BM + (tid - o) + c * BM], mk[o + c], s);
      dm[tid] = ms[tid] - s;
    }
    __syncthreads();
    // Cholesky of PSkp in place (lower), right-looking, one column at a time (:232)
    for (int j = 0; j < n; ++j) {
      if (tid == 0) {
        const double d = Pp[j + (size_t)j * n];
        if (!(d > 0.0)) { s_fail = 1; Pp[j + (size_t)j * n] = 1.0; } else Pp[j + (size_t)j * n] = sqrt(d);
      }
      __syncthreads();
      const double ljj = Pp[j + (size_t)j * n];
      for (int r = j + 1 + tid; r < n; r += nth) Pp[r + (size_t)j * n] /= ljj;
      __syncthreads();
      for (int i = tid; i < (n - j - 1) * (n - j - 1); i += nth) {
        const int r = j + 1 + i % (n - j - 1), c = j + 1 + i / (n - j - 1);
        if (r >= c) Pp[r + (size_t)c * n] = fma(-Pp[r + (size_t)j * n], Pp[c + (size_t)j * n], Pp[r + (size_t)c * n]);
      }
      __syncthreads();
    }
    // G = (PSk*A')/L'/L : row i of G solves  g L L' = T1(i,:)  (:242)
    if (tid < n) {
      const int i = tid;
      for (int j = 0; j < n; ++j) {                  // x L' = b  (forward)
        double v = T1[i + (size_t)j * n];
        for (int l = 0; l < j; ++l) v = fma(-G[i + (size_t)l * n], Pp[j + (size_t)l * n], v);
        G[i + (size_t)j * n] = v / Pp[j + (size_t)j * n];
      }
      for (int j = n - 1; j >= 0; --j) {             // g L = x  (backward), in place
        double v = G[i + (size_t)j * n];
        for (int l = j + 1; l < n; ++l) v = fma(-G[i + (size_t)l * n], Pp[l + (size_t)j * n], v);
        G[i + (size_t)j * n] = v / Pp[j + (size_t)j * n];
      }
    }
    __syncthreads();
    // recompute PSkp (the factorisation overwrote it): Pp = A PSk A' + Q
    for (size_t i = tid; i < nn; i += nth) Pp[i] = Pk[i];
    __syncthreads();
    ekf_predict_cov<BM>(Pp, a, sA, sQ);
    __syncthreads();
    // m = MS(:,k) + G*dm ;  P = PSk + G*(P - PSkp)*G'  (:245-246)
    if (tid < n) {
      double s = mk[tid];
      for (int c = 0; c < n; ++c) s = fma(G[tid + (size_t)c * n], dm[c], s);
      ms[tid] = s;
    }
    for (size_t i = tid; i < nn; i += nth) Ps[i] -= Pp[i];
    __syncthreads();
    mm_nn(G, Ps, T1, n);
    __syncthreads();
    mm_nt(T1, G, Ps, n);
    __syncthreads();
    for (size_t i = tid; i < nn; i += nth) Ps[i] += Pk[i];
    __syncthreads();
    for (int i = tid; i < n; i += nth) a.MS[k * n + i] = ms[i];                 // :249
    for (size_t i = tid; i < nn; i += nth) a.PS[(size_t)k * nn + i] = Ps[i];
    emit(k);
    __syncthreads();
  }
  if (tid < n) a.m_io[tid] = ms[tid];          // m carries into the next global iteration (:165-168)
  if (tid < M) atomic_max_nonneg(maxdiff, md);
  if (tid == 0 && s_fail) atomicCAS(a.status, 0, 1);
}

}  // namespace nsagp
