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
  double* PS;                    // [T][ps_stride], n*n used per step (dense, column-major)
  long long ps_stride;           // doubles per step: n*n rounded up to even, so that every step is 16-byte aligned (bulk stores)
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
      double* dst = a.PS + (size_t)k * a.ps_stride;
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
#ifdef NSAGP_EKF_PROFILE
// cycles between the barriers of one step, summed in registers and written once at the end
// (profiles/microbench/ekf_step.cu)
__device__ long long g_ekf_prof[8];
#define EKF_CLK(i) do { const long long t_ = clock64(); prof_acc[i] += t_ - t_prev; t_prev = t_; } while (0)
#define EKF_CLK_AT(i, who) do { if (tid == (who)) prof_acc[i] += clock64() - t_step; } while (0)
#define EKF_CLK_STEP() const long long t_step = clock64()
#define EKF_CLK_DECL() long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long t_prev = clock64()
#define EKF_CLK_FLUSH(nm, nmean, nthr) do { \
    if (tid == 0) { g_ekf_prof[0] = prof_acc[0]; g_ekf_prof[1] = prof_acc[1]; g_ekf_prof[2] = prof_acc[2]; } \
    if (tid == (nm)) g_ekf_prof[3] = prof_acc[3]; \
    if (tid == (nmean)) g_ekf_prof[4] = prof_acc[4]; \
    if (tid == (nthr)) g_ekf_prof[5] = prof_acc[5]; } while (0)
#else
#define EKF_CLK(i) do { } while (0)
#define EKF_CLK_AT(i, who) do { } while (0)
#define EKF_CLK_STEP() do { } while (0)
#define EKF_CLK_DECL() do { } while (0)
#define EKF_CLK_FLUSH(nm, nmean, nthr) do { } while (0)
#endif

constexpr int ekf_round_bm(int b) { return b <= 2 ? 2 : b <= 3 ? 3 : b <= 4 ? 4 : b <= 6 ? 6 : 8; }
template <int BM> struct EkfF2 { static constexpr int TH = BM <= 3 ? 768 : BM <= 4 ? 512 : 256; };   // register budget of the tile pass

// One (block_i, block_j) tile of exact shape NI x NJ (blocks padded to BM in sA / sQ): finish the pending update,
// stage the updated tile (and its mirror image) for the bulk store, predict in registers, write back.
template <int NI, int NJ, int BM>
__device__ __forceinline__ void ekf_tile(double* P, int n, int oi, int oj, const double* Ai, const double* Aj, const double* Qi,
                                         const double* Ks, const double* Kv, double* Pout, bool mirror) {
  double X[NI * NJ], Tm[NI * NJ];
#pragma unroll
  for (int c = 0; c < NJ; ++c)
#pragma unroll
    for (int r = 0; r < NI; ++r) {
      const double v = fma(-Ks[oi + r], Kv[oj + c], P[(oi + r) + (oj + c) * n]);      // P - K S K' (iekf_update1.m:117)
      X[r + c * NI] = v;
      Pout[(oi + r) + (oj + c) * n] = v;
      if (mirror) Pout[(oj + c) + (oi + r) * n] = v;
    }
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
      P[(oi + r) + (oj + c) * n] = s;
      if (mirror) P[(oj + c) + (oi + r) * n] = s;
    }
}

// BZ, BG > 0: the model's exact block sizes (reference kernels: BZ in {2,4,6,8}, BG in {1,2,3,4}); the updated
// covariance of each step is staged in shared memory and leaves through one bulk copy (cp.async.bulk, the TMA
// engine) per step -- 5 000 scattered 8-byte stores per step from one SM were the slowest part of the step.
// BZ = BG = 0: any block sizes <= BM, tiles padded to BM with run-time guards, direct stores (also the route for
// states too large for the staging buffer).
// Warp roles: warps [0, NMW) ("mean group", NMW = ceil(max(n, M) / 32)) predict the mean and evaluate the
// Jacobian row from it (phases B1, B2; they synchronise among themselves on named barrier 1) WHILE the
// other warps run the tile pass -- the Jacobian needs the predicted mean only, not the predicted covariance.
// The loop body is kept small (single call sites, < 32 KB of SASS): one CTA running alone streams its
// instructions through a 32 KB L1.5 instruction cache, and a larger body made every phase 3-5x slower.
template <int BZ, int BG, int BM>
__global__ void __launch_bounds__(EkfF2<BM>::TH, 1)
giekf_filter2_kernel(const EkfArgs* __restrict__ argv, int l_iter, int energy) {
  constexpr bool kStaged = BZ > 0;
  const EkfArgs& a = argv[blockIdx.x];
  const int tid = threadIdx.x, nth = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int D = a.D, N = a.N, M = a.M, n = a.n;
  const long long T = a.T, ps_stride = a.ps_stride;
  const double sigma2 = a.sigma2;
  const double* __restrict__ yv = a.y;
  double* __restrict__ MSg = a.MS;
  double* __restrict__ PSg = a.PS;
  extern __shared__ __align__(16) double sm[];
  double* Pstage = sm;                       // [ps_stride] (kStaged only) 16-byte aligned source of the bulk store
  double* P = sm + (kStaged ? ps_stride : 0); // [n*n]
  double* sA = P + (size_t)n * n;            // [M*BM*BM]
  double* sQ = sA + M * BM * BM;
  double* sh = sQ + M * BM * BM;             // [M*BM]
  double* shA = sh + M * BM;                 // [M*BM]  h*A
  double* sW = shA + M * BM;                 // [D*N]
  double* mbuf = sW + D * N;                 // [2][n]
  double* JH = mbuf + 2 * n;                 // [n]
  double* PJ = JH + n;                       // [n]
  double* Kv = PJ + n;                       // [n]  K
  double* Ks = Kv + n;                       // [n]  K*S
  double* red = Ks + n;                      // [2][32]
  double* fv = red + 64;                     // [M]  H m
  double* spv = fv + M;                      // [N]  linkf(g)
  double* dlv = spv + N;                     // [N]  dlinkf(g)
  double* sS = dlv + N;                      // [2]  S, y - h(m)
  double* zWv = sS + 2;                      // [N]  z'W
  double* mup = zWv + N;                     // [D]  terms of h(m)
  __shared__ int s_blk[160];
  __shared__ int s_off[kMaxSites + 1];
  __shared__ unsigned short s_pair[(kMaxSites * (kMaxSites + 1)) / 2 + 96];
  if (kStaged) for (int i = tid; i < ps_stride; i += nth) Pstage[i] = 0.0;
  for (int i = tid; i < n * n; i += nth) P[i] = a.Pinf[i];                     // :168
  for (int i = tid; i < M * BM * BM; i += nth) { sA[i] = a.A[i]; sQ[i] = a.Q[i]; }
  for (int i = tid; i < M * BM; i += nth) {
    sh[i] = a.h[i];
    const int bq = i / BM, cq = i % BM;
    double acc = 0.0;
    for (int r = 0; r < BM; ++r) acc = fma(a.h[bq * BM + r], a.A[bq * BM * BM + r + cq * BM], acc);
    shA[i] = acc;
  }
  for (int i = tid; i < D * N; i += nth) sW[i] = a.W[i];
  for (int i = tid; i < n; i += nth) { mbuf[i] = a.m_io[i]; Kv[i] = 0.0; Ks[i] = 0.0; }
  for (int b = tid; b <= M; b += nth) s_off[b] = a.off[b];
  for (int b = tid; b < M; b += nth)
    for (int i = a.off[b]; i < a.off[b + 1]; ++i) s_blk[i] = b;
  // pair list of the exact-shape path: [g,g lower incl. diagonal | g,z | z,z lower incl. diagonal], every class
  // starting at a multiple of 32 so that no warp mixes tile shapes (0xffff = padding entry)
  const int nzz = D * (D + 1) / 2, ngz = N * D, ngg = N * (N + 1) / 2;
  const int o_gz = (ngg + 31) & ~31, o_zz = o_gz + ((ngz + 31) & ~31);
  const int npairs = BZ > 0 ? o_zz + nzz : M * M;
  if (BZ > 0) {
    for (int p = tid; p < npairs; p += nth) {
      int bi = 255, bj = 255;
      if (p < ngg) { int ii = 0; while ((ii + 1) * (ii + 2) / 2 <= p) ++ii; bi = D + ii; bj = D + p - ii * (ii + 1) / 2; }
      else if (p >= o_gz && p < o_gz + ngz) { const int u = p - o_gz; bi = D + u / D; bj = u % D; }
      else if (p >= o_zz) { const int u = p - o_zz; bi = 0; while ((bi + 1) * (bi + 2) / 2 <= u) ++bi; bj = u - bi * (bi + 1) / 2; }
      s_pair[p] = (unsigned short)(bi | (bj << 8));
    }
  }
  __syncthreads();
  const int NMW = max(2, (max(n, M) + 31) >> 5);       // warps of the mean group
  const int n_mean = NMW * 32;
  const bool mean_grp = warp < NMW;
  const int t_idx = tid - n_mean, t_cnt = nth - n_mean; // tile-pass thread index / count
  const int issuer = n_mean;                            // first thread of the tile warps issues the bulk stores
  double* m = mbuf;
  double* m2 = mbuf + n;
  double e_acc = 0.0;
  bool bad = false;
  int my_b = 0, my_o = 0, my_nb = 0;
  double my_h = 0.0;
  if (tid < n) { my_b = s_blk[tid]; my_o = s_off[my_b]; my_nb = s_off[my_b + 1] - my_o; my_h = sh[my_b * BM + (tid - my_o)]; }
  double y_next = yv[0];
  EKF_CLK_DECL();
  auto bar_mean = [&]() { asm volatile("bar.sync 1, %0;" ::"r"(n_mean) : "memory"); };

  // Step k = T is the flush of the last covariance (:202): tile pass + store only (predict mode).
  for (long long k = 0; k <= T; ++k) {
    const bool last = k == T;
    if (last && energy) break;
    const double y = last ? NAN : y_next;
    if (k + 1 < T) y_next = yv[k + 1];
    const bool do_pred = k > 0 || energy;                                      // :180-183 / :392-393
    const bool upd = !last && (!isnan(y) || energy);                           // :186
    const bool st = !energy && k > 0;                                          // PS(:,:,k-1) leaves in this step
    const int iters = upd ? (energy ? 1 : l_iter) : 0;
    double mu_part = 0.0, Kr = 0.0, S = 0.0;
    EKF_CLK_STEP();
    for (int it = 0; it < max(iters, 1); ++it) {                               // iekf_update1.m:110-116
      // ---- A (tile warps, first iteration)  ||  mean prediction, B1, B2 (mean group)
      if (mean_grp) {
        const bool pred_now = it == 0 && do_pred && !last;
        const double* mc = pred_now ? m2 : m;
        if (it > 0) bar_mean();                                                // phase D's mean is complete
        // B1: f = H (A m) = (H A) m straight from the previous mean (hA is precomputed: no wait for the predicted
        // mean), the predicted mean row by row on the state threads, z'W reduced over the subbands by warp 0, the
        // modulators' threads evaluate the link and its derivative
        const double* hsel = pred_now ? shA : sh;
        const double* msrc = pred_now ? m : mc;
        double f = 0.0;
        if (tid < M && (upd || pred_now)) {
          const int o = s_off[tid], nb = s_off[tid + 1] - o;
#pragma unroll 1
          for (int c = 0; c < nb; ++c) f = fma(hsel[tid * BM + c], msrc[o + c], f);
          fv[tid] = f;
        }
        if (pred_now && tid < n) {
          const int nb = s_off[my_b + 1] - my_o;
          double mv = 0.0;
#pragma unroll 1
          for (int c = 0; c < nb; ++c) mv = fma(sA[my_b * BM * BM + (tid - my_o) + c * BM], m[my_o + c], mv);
          m2[tid] = mv;
        }
        if (upd) {
          if (warp == 0) {
            double f2 = 0.0;                                                   // subband lane + 32 (D <= 64)
            const int d2 = lane + 32;
            if (d2 < D) {
              const int o = s_off[d2], nb = s_off[d2 + 1] - o;
#pragma unroll 1
              for (int c = 0; c < nb; ++c) f2 = fma(hsel[d2 * BM + c], msrc[o + c], f2);
            }
            const double f1 = lane < D ? f : 0.0;
#pragma unroll 1
            for (int j = 0; j < N; j += 2) {                                   // two modulators per round: independent chains
              const bool two = j + 1 < N;
              double za = lane < D ? f1 * sW[lane * N + j] : 0.0;
              double zb = (two && lane < D) ? f1 * sW[lane * N + j + 1] : 0.0;
              if (d2 < D) { za = fma(f2, sW[d2 * N + j], za); if (two) zb = fma(f2, sW[d2 * N + j + 1], zb); }
#pragma unroll
              for (int o = 16; o >= 1; o >>= 1) {
                za += __shfl_xor_sync(0xffffffffu, za, o);
                zb += __shfl_xor_sync(0xffffffffu, zb, o);
              }
              if (lane == 0) { zWv[j] = za; if (two) zWv[j + 1] = zb; }
            }
          }
          if (tid >= D && tid < M) {
            const double eg = exp_fast(f);
            spv[tid - D] = log_ge1_fast(1.0 + eg);                             // linkf(g) = log(1+exp(g))
            dlv[tid - D] = eg * rcp_fast(eg + 1.0);                            // dlinkf(g) = exp(g)/(exp(g)+1)
          }
          bar_mean();
          // B2: Jacobian row entry of this thread's state (:497-503), terms of h(m) (:490-494)
          if (tid < n) {
            double jh;
            if (my_b < D) {
              double wl = 0.0;
              for (int j = 0; j < N; ++j) wl = fma(sW[my_b * N + j], spv[j], wl);
              jh = wl * my_h;
              if (tid == my_o) mup[my_b] = fv[my_b] * wl;
            } else {
              jh = zWv[my_b - D] * dlv[my_b - D] * my_h;
            }
            JH[tid] = jh;
          }
        } else if (pred_now) {
          bar_mean();
        }
        EKF_CLK_AT(3, n - 1);
      } else if (it == 0 && do_pred) {
        double* store = PSg + (size_t)(st ? k - 1 : 0) * ps_stride;
        if (BZ > 0) {
#pragma unroll 1
          for (int p = t_idx; p < npairs; p += t_cnt) {
            const int pr = s_pair[p], bi = pr & 255, bj = pr >> 8;
            if (bi == 255) continue;
            const int oi = s_off[bi], oj = s_off[bj];
            const double* Ai = sA + bi * BM * BM;
            const double* Aj = sA + bj * BM * BM;
            const double* Qi = (bi == bj) ? sQ + bi * BM * BM : nullptr;
            if (p < o_gz) ekf_tile<(BG > 0 ? BG : 1), (BG > 0 ? BG : 1), BM>(P, n, oi, oj, Ai, Aj, Qi, Ks, Kv, Pstage, bi != bj);
            else if (p < o_zz) ekf_tile<(BG > 0 ? BG : 1), (BZ > 0 ? BZ : 1), BM>(P, n, oi, oj, Ai, Aj, Qi, Ks, Kv, Pstage, true);
            else ekf_tile<(BZ > 0 ? BZ : 1), (BZ > 0 ? BZ : 1), BM>(P, n, oi, oj, Ai, Aj, Qi, Ks, Kv, Pstage, bi != bj);
          }
          if (st) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // staged tiles -> visible to the copy engine
        } else {
#pragma unroll 1
          for (int pair = t_idx; pair < npairs; pair += t_cnt) {
            const int bi = pair % M, bj = pair / M;
            const int oi = s_off[bi], ni = s_off[bi + 1] - oi;
            const int oj = s_off[bj], nj = s_off[bj + 1] - oj;
            const double* Ai = sA + bi * BM * BM;
            const double* Aj = sA + bj * BM * BM;
            double X[BM * BM], Tm[BM * BM];
#pragma unroll
            for (int c = 0; c < BM; ++c)
#pragma unroll
              for (int r = 0; r < BM; ++r) {
                double v = 0.0;
                if (r < ni && c < nj) {
                  v = fma(-Ks[oi + r], Kv[oj + c], P[(oi + r) + (size_t)(oj + c) * n]);
                  if (st) store[(oi + r) + (size_t)(oj + c) * n] = v;
                }
                X[r + c * BM] = v;
              }
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
                if (r < ni && c < nj) P[(oi + r) + (size_t)(oj + c) * n] = s2;
              }
          }
        }
        EKF_CLK_AT(4, n_mean);
        EKF_CLK_AT(5, nth - 1);
      }
      if (it == 0 && do_pred && !last) { double* t = m; m = m2; m2 = t; }
      __syncthreads();
      if (kStaged && it == 0 && st && tid == issuer) {                         // one bulk copy of the staged covariance
        const unsigned src = (unsigned)__cvta_generic_to_shared(Pstage);
        double* dst = PSg + (size_t)(k - 1) * ps_stride;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"((unsigned)(ps_stride * 8)) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      EKF_CLK(0);
      if (!upd) break;
      // ---- C: P*JH' (eight lanes per row)
      for (int r0 = 0; r0 < n; r0 += nth >> 3) {                               // uniform trip count: shuffles inside
        const int r = r0 + (tid >> 3);
        double sa = 0.0, sb = 0.0;
        if (r < n) {
          int c = tid & 7;
#pragma unroll 1
          for (; c + 8 < n; c += 16) {
            sa = fma(P[r + (size_t)c * n], JH[c], sa);
            sb = fma(P[r + (size_t)(c + 8) * n], JH[c + 8], sb);
          }
          if (c < n) sa = fma(P[r + (size_t)c * n], JH[c], sa);
        }
        double sv = sa + sb;
        sv += __shfl_xor_sync(0xffffffffu, sv, 1);
        sv += __shfl_xor_sync(0xffffffffu, sv, 2);
        sv += __shfl_xor_sync(0xffffffffu, sv, 4);
        if ((tid & 7) == 0 && r < n) PJ[r] = sv;
      }
      __syncthreads();
      EKF_CLK(1);
      // ---- D (mean group only; the tile warps go straight to the barrier that ends the step)
      if (mean_grp) {
        if (warp == 0) {                                                       // S = R + JH P JH' (:S), MU = h(m)
          double sv = 0.0, mv = 0.0;
          for (int r = lane; r < n; r += 32) sv = fma(JH[r], PJ[r], sv);
          for (int d = lane; d < D; d += 32) mv += mup[d];
#pragma unroll
          for (int o = 16; o >= 1; o >>= 1) {
            sv += __shfl_xor_sync(0xffffffffu, sv, o);
            mv += __shfl_xor_sync(0xffffffffu, mv, o);
          }
          if (lane == 0) { sS[0] = sigma2 + sv; sS[1] = y - mv; }
        }
        bar_mean();
        S = sS[0];
        const double v = sS[1];
        if (energy) {
          if (!(S > 0.0)) bad = true;                                          // :417-427 -> NaN energy
          if (tid == 0) e_acc += 0.5 * log(2.0 * 3.14159265358979323846) + log(sqrt(S)) + 0.5 * v * v / S;
        }
        if (tid < n) {
          Kr = PJ[tid] * rcp_fast(S);                                          // K = P JH' / S
          m[tid] = fma(Kr, v, m[tid]);                                         // M = M + K (y - MU)
        }
      }
    }
    if (last) break;
    if (tid < n) {
      Kv[tid] = Kr; Ks[tid] = Kr * S;                                          // consumed by the next tile pass
      if (!energy) MSg[k * n + tid] = m[tid];                                  // :201
    }
    if (kStaged && st && tid == issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging buffer reusable
    __syncthreads();
    EKF_CLK(2);
  }
  if (kStaged && !energy && tid == issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  EKF_CLK_FLUSH(n - 1, n_mean, nth - 1);
  if (energy && tid == 0) {                                                    // thread 0 is in the mean group, which tracks `bad`
    a.edata[0] = bad ? NAN : e_acc;
    if (bad || isnan(e_acc)) atomicCAS(a.status, 0, 3);
  }
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
  for (size_t i = tid; i < nn; i += nth) Ps[i] = a.PS[(size_t)(T - 1) * a.ps_stride + i];
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
    for (size_t i = tid; i < nn; i += nth) { const double v = a.PS[(size_t)k * a.ps_stride + i]; Pk[i] = v; Pp[i] = v; }
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
    for (size_t i = tid; i < nn; i += nth) a.PS[(size_t)k * a.ps_stride + i] = Ps[i];
    emit(k);
    __syncthreads();
  }
  if (tid < n) a.m_io[tid] = ms[tid];          // m carries into the next global iteration (:165-168)
  if (tid < M) atomic_max_nonneg(maxdiff, md);
  if (tid == 0 && s_fail) atomicCAS(a.status, 0, 1);
}

}  // namespace nsagp
