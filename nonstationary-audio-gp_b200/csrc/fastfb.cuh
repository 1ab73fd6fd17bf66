// Stationary (infinite-horizon) Kalman filter / RTS smoother of the probabilistic filter bank -- the step BEFORE the
// hot path, which initialises the subbands and W.
//
// Reference: matlab/unifying_prob_tf/kernel_ss_kalmanFastFB.m:46-151.  ONE scalar observation couples all subbands, so
// the constant-gain recursion is dense in the state (n = 2 tau D, 16..96):
//     filter    m_k = (A - K H A) m_{k-1} + K y_k          (y_k observed; :90-101)
//               m_k = A m_{k-1}                             (y_k missing;  :103-106)
//     smoother  m_k = MS_k + G (m_{k+1} - A MS_k)           (:133-143)
// Both are affine in m with (almost) constant matrices, so a pass is a scan over chunks of kFbChunk steps:
//   reduce : one CTA per chunk runs the chunk from a ZERO state (the chunk's offset z_c); a chunk without missing samples
//            maps its entering state by F^len, computed once by repeated squaring; a chunk with missing samples also
//            propagates the identity to get its own transfer matrix;
//   carry  : one CTA walks the chunks, m_in(c+1) = Phi_c m_in(c) + z_c;
//   apply  : one CTA per chunk re-runs the reference's literal steps from its entering state, stores MS and the
//            chunk's share of sum v^2 / (2 S).
// Only the chunk-entry states are re-associated (tolerance class 1e-6).
#pragma once
#include "common.cuh"

namespace nsagp {

constexpr int kFbChunk = 256;      // steps per chunk (a power of two: F^256 by eight squarings)
constexpr int kFbMaxN = 96;

struct FbArgs {
  int n, smooth;                   // smooth: 0 forward filter pass, 1 backward smoother pass
  long long T;
  const double* F;                 // [n*n] filter: A - K H A; smoother: G
  const double* A;                 // [n*n]
  const double* Kg;                // [n] gain (filter)
  const double* HA;                // [n] H A (filter: innovation)
  double S;
  const double* y;                 // [T]
  double* MS;                      // [T][n] filter: out; smoother: in (filtered) / out (smoothed)
  double* Fpow;                    // [n*n] F^kFbChunk
  double* z;                       // [nchunks][n]
  double* Phi;                     // [nchunks][n*n] (only chunks with missing samples)
  int* has_nan;                    // [nchunks]
  double* m_in;                    // [nchunks][n] state entering the chunk (in processing order)
  double* lik_part;                // [nchunks]
};

// y = Mx v for an n x n column-major matrix in shared memory; thread i < n owns row i.
__device__ __forceinline__ double fb_row_dot(const double* Mx, const double* v, int n, int i) {
  double s0 = 0.0, s1 = 0.0;
  int j = 0;
  for (; j + 1 < n; j += 2) {
    s0 = fma(Mx[i + j * n], v[j], s0);
    s1 = fma(Mx[i + (j + 1) * n], v[j + 1], s1);
  }
  if (j < n) s0 = fma(Mx[i + j * n], v[j], s0);
  return s0 + s1;
}

// F^kFbChunk by repeated squaring (one CTA, n*n threads-strided).
__global__ void fb_power_kernel(FbArgs a) {
  extern __shared__ double sm[];
  const int n = a.n, nn = n * n;
  double* X = sm;
  double* Y = sm + nn;
  for (int i = threadIdx.x; i < nn; i += blockDim.x) X[i] = a.F[i];
  __syncthreads();
  for (int it = 1; it < kFbChunk; it <<= 1) {
    for (int i = threadIdx.x; i < nn; i += blockDim.x) {
      const int r = i % n, c = i / n;
      double s = 0.0;
      for (int k = 0; k < n; ++k) s = fma(X[r + k * n], X[k + c * n], s);
      Y[i] = s;
    }
    __syncthreads();
    double* t = X; X = Y; Y = t;
  }
  for (int i = threadIdx.x; i < nn; i += blockDim.x) a.Fpow[i] = X[i];
}

// One step of the recursion on the state in shared memory (v_in -> v_out), all threads of the CTA.
// FILTER: returns (through vinn, thread 0) the innovation y - HA m of an observed step.
__device__ __forceinline__ void fb_step(const FbArgs& a, const double* sF, const double* sA, const double* sK, const double* sHA,
                                        long long k, const double* v_in, double* v_out, double* tmp, bool zero_input,
                                        double& quad) {
  const int n = a.n, i = threadIdx.x;
  if (!a.smooth) {
    const double yk = a.y[k];
    const bool obs = !isnan(yk);
    if (i < n) {
      v_out[i] = obs ? fb_row_dot(sF, v_in, n, i) + (zero_input ? 0.0 : sK[i] * yk) : fb_row_dot(sA, v_in, n, i);
    }
    if (i == 0 && obs && !zero_input) {
      double hm = 0.0;
      for (int j = 0; j < n; ++j) hm = fma(sHA[j], v_in[j], hm);
      const double v = yk - hm;                                 // :93
      quad += 0.5 * v * v / a.S;                                // :99
    }
    __syncthreads();
  } else {
    // m = MS_k + G (m - A MS_k); zero_input: the homogeneous part only, m = G m
    const double* msk = a.MS + k * n;
    if (i < n) tmp[i] = zero_input ? v_in[i] : v_in[i] - fb_row_dot(sA, msk, n, i);
    __syncthreads();
    if (i < n) v_out[i] = (zero_input ? 0.0 : msk[i]) + fb_row_dot(sF, tmp, n, i);
    __syncthreads();
  }
}

// mode 0: reduce (zero-state response + transfer matrix of chunks with missing samples); mode 1: apply.
__global__ void fb_chunk_kernel(FbArgs a, int mode) {
  extern __shared__ double sm[];
  const int n = a.n, nn = n * n, tid = threadIdx.x;
  double* sF = sm;
  double* sA = sF + nn;
  double* sK = sA + nn;
  double* sHA = sK + n;
  double* va = sHA + n;
  double* vb = va + n;
  double* tmp = vb + n;
  for (int i = tid; i < nn; i += blockDim.x) { sF[i] = a.F[i]; sA[i] = a.A[i]; }
  for (int i = tid; i < n; i += blockDim.x) { sK[i] = a.Kg ? a.Kg[i] : 0.0; sHA[i] = a.HA ? a.HA[i] : 0.0; }
  const long long nsteps = a.smooth ? a.T - 1 : a.T;            // the smoother leaves step T-1 as it is
  const long long nchunks = (nsteps + kFbChunk - 1) / kFbChunk;
  for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const long long s0 = c * kFbChunk, s1 = min(s0 + (long long)kFbChunk, nsteps);
    auto K = [&](long long s) { return a.smooth ? a.T - 2 - s : s; };      // processing order -> time
    __syncthreads();
    double quad = 0.0;
    if (mode == 0) {
      // the chunk from a zero state WITH its inputs: z_c
      if (tid < n) va[tid] = 0.0;
      __syncthreads();
      double* vi = va; double* vo = vb;
      bool any_nan = false;
      for (long long s = s0; s < s1; ++s) {
        if (!a.smooth && isnan(a.y[K(s)])) any_nan = true;
        fb_step(a, sF, sA, sK, sHA, K(s), vi, vo, tmp, false, quad);
        double* t = vi; vi = vo; vo = t;
      }
      if (tid < n) a.z[c * n + tid] = vi[tid];
      if (tid == 0) a.has_nan[c] = any_nan ? 1 : 0;
      if (any_nan) {
        // transfer matrix of this chunk: the homogeneous recursion applied to the columns of the identity
        for (int col = 0; col < n; ++col) {
          __syncthreads();
          if (tid < n) va[tid] = (tid == col) ? 1.0 : 0.0;
          __syncthreads();
          vi = va; vo = vb;
          for (long long s = s0; s < s1; ++s) {
            fb_step(a, sF, sA, sK, sHA, K(s), vi, vo, tmp, true, quad);
            double* t = vi; vi = vo; vo = t;
          }
          if (tid < n) a.Phi[(size_t)c * nn + tid + (size_t)col * n] = vi[tid];
        }
      }
    } else {
      if (tid < n) va[tid] = a.m_in[c * n + tid];
      __syncthreads();
      double* vi = va; double* vo = vb;
      for (long long s = s0; s < s1; ++s) {
        const long long k = K(s);
        fb_step(a, sF, sA, sK, sHA, k, vi, vo, tmp, false, quad);
        if (tid < n) a.MS[k * n + tid] = vo[tid];              // (:109, :146)
        double* t = vi; vi = vo; vo = t;
      }
      if (tid == 0 && a.lik_part) a.lik_part[c] = quad;
    }
  }
}

// One CTA: the state entering every chunk.  Filter: m starts at zero (:38); smoother: at the filtered mean of step T-1.
__global__ void fb_carry_kernel(FbArgs a) {
  extern __shared__ double sm[];
  const int n = a.n, nn = n * n, tid = threadIdx.x;
  double* sP = sm;                 // F^chunk
  double* sX = sP + nn;            // a chunk's own transfer matrix
  double* va = sX + nn;
  double* vb = va + n;
  for (int i = tid; i < nn; i += blockDim.x) sP[i] = a.Fpow[i];
  if (tid < n) va[tid] = a.smooth ? a.MS[(a.T - 1) * n + tid] : 0.0;
  __syncthreads();
  const long long nsteps = a.smooth ? a.T - 1 : a.T;
  const long long nchunks = (nsteps + kFbChunk - 1) / kFbChunk;
  double* vi = va; double* vo = vb;
  for (long long c = 0; c < nchunks; ++c) {
    if (tid < n) a.m_in[c * n + tid] = vi[tid];
    if (c + 1 == nchunks) break;
    const double* Mx = sP;
    if (a.has_nan[c]) {
      for (int i = tid; i < nn; i += blockDim.x) sX[i] = a.Phi[(size_t)c * nn + i];
      __syncthreads();
      Mx = sX;
    }
    if (tid < n) vo[tid] = fb_row_dot(Mx, vi, n, tid) + a.z[c * n + tid];
    __syncthreads();
    double* t = vi; vi = vo; vo = t;
  }
}

// deterministic sum of the chunks' shares
__global__ void fb_sum_kernel(const double* __restrict__ part, long long nchunks, double* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (long long c = 0; c < nchunks; ++c) s += part[c];
    *out = s;
  }
}

}  // namespace nsagp
