// Batched moment matching: T independent (y, mu, s2) triples -> (lZ, dlZ, d2lZ).
// C ABI: nsagp_mom_batch / nsagp_mom_batch_warp (include/nsagp.h).  These are the
// two device forms of mom.cuh exposed on their own so that parity tests can hit
// the likelihood arithmetic directly (reference: likModulatorNMFPower.m:28-87).
#pragma once
#include "common.cuh"
#include "mom.cuh"

namespace nsagp {

struct MomBatchArgs {
  MomParams p;            // W / wn / xn are device pointers
  double alpha;
  long long T;
  int M;
  const double* y;        // [T]
  const double* mu;       // [T][M]
  const double* s2;       // [T][M]
  double* lZ;             // [T]
  double* d1;             // [T][M]
  double* d2;             // [T][M]
};

template <int DP, int TPB>
__global__ void __launch_bounds__(TPB) mom_batch_thread_kernel(MomBatchArgs a) {
  extern __shared__ double sm[];
  const int tid = threadIdx.x;
  const int M = a.M;
  double* s_mu = sm;
  double* s_s2 = s_mu + M * TPB;
  double* s_d1 = s_s2 + M * TPB;
  double* s_d2 = s_d1 + M * TPB;
  double* s_W = s_d2 + M * TPB;
  double* s_wn = s_W + DP * kNP;
  double* s_xn = s_wn + a.p.S;
  for (int i = tid; i < DP * kNP; i += TPB) s_W[i] = a.p.W[i];
  for (int i = tid; i < a.p.S; i += TPB) s_wn[i] = a.p.wn[i];
  for (int i = tid; i < kNP * a.p.S; i += TPB) s_xn[i] = a.p.xn[i];
  __syncthreads();
  const long long k = (long long)blockIdx.x * TPB + tid;
  if (k >= a.T) return;
  MomParams p = a.p;
  p.W = s_W; p.wn = s_wn; p.xn = s_xn;
  for (int n = 0; n < M; ++n) {
    s_mu[n * TPB + tid] = a.mu[k * M + n];
    s_s2[n * TPB + tid] = a.s2[k * M + n];
  }
  a.lZ[k] = mom_thread<DP>(p, a.alpha, a.y[k], s_mu + tid, s_s2 + tid, TPB, s_d1 + tid, s_d2 + tid);
  for (int n = 0; n < M; ++n) {
    a.d1[k * M + n] = s_d1[n * TPB + tid];
    a.d2[k * M + n] = s_d2[n * TPB + tid];
  }
}

// One warp per step, WPB warps per CTA.
template <int DP, int WPB>
__global__ void __launch_bounds__(32 * WPB) mom_batch_warp_kernel(MomBatchArgs a) {
  extern __shared__ double sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int M = a.M;
  double* s_W = sm;
  double* s_wn = s_W + DP * kNP;
  double* s_xn = s_wn + a.p.S;
  double* s_ms = s_xn + kNP * a.p.S;     // [WPB][64]
  for (int i = tid; i < DP * kNP; i += 32 * WPB) s_W[i] = a.p.W[i];
  for (int i = tid; i < a.p.S; i += 32 * WPB) s_wn[i] = a.p.wn[i];
  for (int i = tid; i < kNP * a.p.S; i += 32 * WPB) s_xn[i] = a.p.xn[i];
  __syncthreads();
  const long long k = (long long)blockIdx.x * WPB + warp;
  if (k >= a.T) return;
  MomParams p = a.p;
  p.W = s_W; p.wn = s_wn; p.xn = s_xn;
  double* mu = s_ms + warp * 64;
  double* s2 = mu + 32;
  if (lane < M) { mu[lane] = a.mu[k * M + lane]; s2[lane] = a.s2[k * M + lane]; }
  __syncwarp();
  double d1, d2;
  const double lz = mom_warp<DP>(p, a.alpha, a.y[k], mu, s2, lane, d1, d2);
  if (lane == 0) a.lZ[k] = lz;
  if (lane < M) { a.d1[k * M + lane] = d1; a.d2[k * M + lane] = d2; }
}

}  // namespace nsagp
