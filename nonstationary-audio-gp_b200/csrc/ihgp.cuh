// Infinite-horizon (steady-state gain) Power-EP kernels.
//
// Reference recursion: matlab/ihgp_ep_modulator_nmf.m:233-310 (filter),
// :373-442 (smoother + site update).  Covariances are never propagated: each
// block looks its predictive covariance / smoother gain up in a table indexed
// by the site's equivalent noise R = 1/ttau, so only the n-vector of means is a
// recurrence.  That recurrence is
//   * nonlinear in the ADF pass (sites at k come from the prediction at k), so
//     it runs sequentially, one warp per signal, lanes over the D+N blocks for
//     the Kalman part and over sigma points for the moment matching;
//   * affine once the sites are frozen (filter passes >= 2, every smoother
//     pass): m_k = F_k m_{k-1} + c_k per block, evaluated by a chunked
//     three-phase scan (compose per chunk, carry across chunks, re-apply);
//   * absent in the site update of the smoother pass, which is independent per
//     time step: one thread per step.
#pragma once
#include "common.cuh"
#include "lookup.cuh"
#include "mom.cuh"

namespace nsagp {

constexpr int kScanChunk = 128;    // time steps composed by one thread in the scans

__device__ __forceinline__ MomParams make_mom_params(const DevProblem& P, const double* W,
                                                     const double* wn, const double* xn) {
  MomParams mp;
  mp.D = P.D; mp.N = P.N; mp.S = P.S; mp.kind = P.lik_kind;
  mp.sn2 = P.sn2; mp.shift = P.link_shift;
  mp.W = W; mp.wn = wn; mp.xn = xn;
  return mp;
}

// ------------------------------------------------------------------ ADF pass
// One warp per signal.  Steps k0..k1-1.  mom_all: moment matching at every step
// (first pass) or only at k == T-1 (later passes call this for the last step).
// running: the _constraints nlZ variant's running site vectors.
template <int DP, int BM>
__global__ void __launch_bounds__(32)
ihgp_adf_kernel(const DevProblem* __restrict__ probs, const DevState* __restrict__ states,
                long long T, long long k0, long long k1, int mom_all, double ep_damp, int running) {
  const DevProblem& P = probs[blockIdx.x];
  const DevState& St = states[blockIdx.x];
  const int lane = threadIdx.x;
  const int M = P.M, nr = P.nr;
  const bool active = lane < M;
  const int n = active ? lane : M - 1;

  extern __shared__ double sm[];
  double* s_mu = sm;                      // [32]
  double* s_s2 = sm + 32;                 // [32]
  double* s_W = sm + 64;                  // [DP*kNP]
  double* s_wn = s_W + DP * kNP;          // [S]
  double* s_xn = s_wn + P.S;              // [kNP*S]
  for (int i = lane; i < DP * kNP; i += 32) s_W[i] = P.W[i];
  for (int i = lane; i < P.S; i += 32) s_wn[i] = P.wn[i];
  for (int i = lane; i < kNP * P.S; i += 32) s_xn[i] = P.xn[i];
  __syncwarp();
  const MomParams mp = make_mom_params(P, s_W, s_wn, s_xn);

  double A[BM * BM], hA[BM], hv[BM], m[BM];
#pragma unroll
  for (int i = 0; i < BM * BM; ++i) A[i] = P.A[n * BM * BM + i];
#pragma unroll
  for (int i = 0; i < BM; ++i) { hA[i] = P.hA[n * BM + i]; hv[i] = P.h[n * BM + i]; }
  const int off = P.off[n];
  const int b = P.off[n + 1] - off;

  int idx;
  if (k0 == 0) {
    idx = nr;                             // PP = Pinf at the first step (:246)
#pragma unroll
    for (int i = 0; i < BM; ++i) m[i] = St.mcarry[n * BM + i];
  } else {
#pragma unroll
    for (int i = 0; i < BM; ++i) m[i] = (i < b) ? St.MS[(k0 - 1) * P.n + off + i] : 0.0;
    const double ttp = fmax(St.ttau[(k0 - 1) * M + n], 0.0);
    const double Rp = (ttp == 0.0) ? INFINITY : St.R[(k0 - 1) * M + n];
    idx = lookup_filter(P.r, P.thr, nr, Rp);
  }
  double tt_run = 0.0, tn_run = 0.0;

  for (long long k = k0; k < k1; ++k) {
    const double y = St.y[k];
    double Am[BM];
    double fmu = 0.0;
#pragma unroll
    for (int i = 0; i < BM; ++i) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < BM; ++j) acc = fma(A[i + j * BM], m[j], acc);
      Am[i] = acc;
      fmu = fma(hA[i], m[i], fmu);        // fmu = (H*A)*m (:250)
    }
    const double* wrow = P.Wtab + ((size_t)n * (nr + 1) + idx) * BM;
    const double HPH = P.HPHtab[(size_t)n * (nr + 1) + idx];
    double Wv[BM];
#pragma unroll
    for (int i = 0; i < BM; ++i) Wv[i] = wrow[i];

    double tt, tn, Rk;
    if (mom_all || k == T - 1) {
      if (active) { s_mu[lane] = fmu; s_s2[lane] = HPH; }
      __syncwarp();
      double d1, d2;
      const double lz = mom_warp<DP>(mp, 1.0, y, s_mu, s_s2, lane, d1, d2);   // alpha = 1 in the filter (:256)
      __syncwarp();
      const double tt_old = running ? tt_run : St.ttau[k * M + n];
      const double tn_old = running ? tn_run : St.tnu[k * M + n];
      const double den = 1.0 + d2 * HPH;
      tt = (1.0 - ep_damp) * tt_old + ep_damp * (-d2 / den);                  // :265
      tn = (1.0 - ep_damp) * tn_old + ep_damp * ((d1 - fmu * d2) / den);      // :266
      Rk = 1.0 / tt;                                                          // :269 (before the clamp)
      if (lane == 0) St.lZ[k] = lz;
    } else {
      tt = St.ttau[k * M + n]; tn = St.tnu[k * M + n]; Rk = St.R[k * M + n];
    }
    tt = fmax(tt, 0.0);                   // NaN -> 0, as MATLAB max (:274)
    if (tt == 0.0) {
      Rk = INFINITY;                      // :287
#pragma unroll
      for (int i = 0; i < BM; ++i) m[i] = Am[i];
    } else {
      const double ys = tn / tt;          // :277
      const double g = 1.0 / (HPH + Rk);
      const double innov = ys - fmu;
#pragma unroll
      for (int i = 0; i < BM; ++i) m[i] = fma(Wv[i] * g, innov, Am[i]);       // (A-K h A) m + K ys
    }
    if (active) {
      St.ttau[k * M + n] = tt; St.tnu[k * M + n] = tn; St.R[k * M + n] = Rk;
#pragma unroll
      for (int i = 0; i < BM; ++i) if (i < b) St.MS[k * P.n + off + i] = m[i];
      if (k == T - 1) {
        double e = 0.0;
#pragma unroll
        for (int i = 0; i < BM; ++i) e = fma(hv[i], m[i], e);
        St.E[k * M + n] = e;
      }
    }
    idx = lookup_filter(P.r, P.thr, nr, Rk);
    tt_run = tt; tn_run = tn;
  }
}

// ------------------------------------------------------ affine scan elements
template <int BM>
struct Affine {
  double F[BM * BM];     // column-major
  double c[BM];
};

// Frozen-site filter step k as an affine map of the block mean (:280-304).
template <int BM>
struct FilterElem {
  const DevProblem& P; const DevState& St; int n;
  double A[BM * BM], hA[BM];
  __device__ FilterElem(const DevProblem& P_, const DevState& S_, int n_) : P(P_), St(S_), n(n_) {
#pragma unroll
    for (int i = 0; i < BM * BM; ++i) A[i] = P.A[n * BM * BM + i];
#pragma unroll
    for (int i = 0; i < BM; ++i) hA[i] = P.hA[n * BM + i];
  }
  // writes back the clamp / Inf marking of the reference's filter loop when `commit`
  __device__ __forceinline__ void get(long long k, Affine<BM>& e, bool commit) const {
    const int M = P.M, nr = P.nr;
    const double tt = fmax(St.ttau[k * M + n], 0.0);                          // :274
    if (tt == 0.0) {
#pragma unroll
      for (int i = 0; i < BM * BM; ++i) e.F[i] = A[i];
#pragma unroll
      for (int i = 0; i < BM; ++i) e.c[i] = 0.0;
      if (commit) { St.ttau[k * M + n] = 0.0; St.R[k * M + n] = INFINITY; }
      return;
    }
    int idx = nr;
    if (k > 0) {
      const double ttp = fmax(St.ttau[(k - 1) * M + n], 0.0);
      const double Rp = (ttp == 0.0) ? INFINITY : St.R[(k - 1) * M + n];
      idx = lookup_filter(P.r, P.thr, nr, Rp);
    }
    const double* wrow = P.Wtab + ((size_t)n * (nr + 1) + idx) * BM;
    const double HPH = P.HPHtab[(size_t)n * (nr + 1) + idx];
    const double g = 1.0 / (HPH + St.R[k * M + n]);
    const double ys = St.tnu[k * M + n] / tt;
#pragma unroll
    for (int i = 0; i < BM; ++i) {
      const double Ki = wrow[i] * g;
      e.c[i] = Ki * ys;
#pragma unroll
      for (int j = 0; j < BM; ++j) e.F[i + j * BM] = fma(-Ki, hA[j], A[i + j * BM]);   // A - K (h A)
    }
  }
  __device__ __forceinline__ void emit(long long k, const double (&m)[BM]) const {
    const int off = P.off[n], b = P.off[n + 1] - off;
#pragma unroll
    for (int i = 0; i < BM; ++i) if (i < b) St.MS[k * P.n + off + i] = m[i];
  }
};

// RTS mean step k (:379-394): m_k = MS_k + G_k (m_{k+1} - A MS_k).
template <int BM>
struct SmootherElem {
  const DevProblem& P; const DevState& St; int n;
  double A[BM * BM], hv[BM];
  __device__ SmootherElem(const DevProblem& P_, const DevState& S_, int n_) : P(P_), St(S_), n(n_) {
#pragma unroll
    for (int i = 0; i < BM * BM; ++i) A[i] = P.A[n * BM * BM + i];
#pragma unroll
    for (int i = 0; i < BM; ++i) hv[i] = P.h[n * BM + i];
  }
  __device__ __forceinline__ void get(long long k, Affine<BM>& e, bool commit) const {
    const int M = P.M, nr = P.nr;
    const int off = P.off[n], b = P.off[n + 1] - off;
    const int idx = lookup_smoother(P.r, P.thr, nr, St.R[k * M + n]);
    const double* G = P.Gtab + ((size_t)n * nr + idx) * BM * BM;
    double ms[BM], t[BM];
#pragma unroll
    for (int i = 0; i < BM; ++i) ms[i] = (i < b) ? St.MS[k * P.n + off + i] : 0.0;
#pragma unroll
    for (int i = 0; i < BM; ++i) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < BM; ++j) acc = fma(A[i + j * BM], ms[j], acc);
      t[i] = acc;
    }
#pragma unroll
    for (int i = 0; i < BM * BM; ++i) e.F[i] = G[i];
#pragma unroll
    for (int i = 0; i < BM; ++i) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < BM; ++j) acc = fma(e.F[i + j * BM], t[j], acc);
      e.c[i] = ms[i] - acc;
    }
    if (commit && k == 0) {
      // marginal variance of the last look-up: the reference's Varft (:492) and maxDiffP (:444)
      const double vm = P.vmtab[(size_t)n * nr + idx];
      atomic_max_nonneg(St.maxdiff + 1, fabs(St.vm0[n] - vm));
      St.vm0[n] = vm;
    }
  }
  __device__ __forceinline__ void emit(long long k, const double (&m)[BM]) const {
    const int off = P.off[n], b = P.off[n + 1] - off;
    double e = 0.0;
#pragma unroll
    for (int i = 0; i < BM; ++i) {
      if (i < b) St.MS[k * P.n + off + i] = m[i];
      e = fma(hv[i], m[i], e);
    }
    // maxDiffM (:440): against H*MS of the previous iteration's smoother
    atomic_max_nonneg(St.maxdiff, fabs(St.E[k * P.M + n] - e));
    St.E[k * P.M + n] = e;
  }
};

template <int BM>
__device__ __forceinline__ void affine_compose(Affine<BM>& acc, const Affine<BM>& e) {
  // acc <- e o acc :  F = Fe Facc,  c = Fe cacc + ce
  double F[BM * BM], c[BM];
#pragma unroll
  for (int i = 0; i < BM; ++i) {
    double ci = e.c[i];
#pragma unroll
    for (int l = 0; l < BM; ++l) ci = fma(e.F[i + l * BM], acc.c[l], ci);
    c[i] = ci;
#pragma unroll
    for (int j = 0; j < BM; ++j) {
      double s = 0.0;
#pragma unroll
      for (int l = 0; l < BM; ++l) s = fma(e.F[i + l * BM], acc.F[l + j * BM], s);
      F[i + j * BM] = s;
    }
  }
#pragma unroll
  for (int i = 0; i < BM * BM; ++i) acc.F[i] = F[i];
#pragma unroll
  for (int i = 0; i < BM; ++i) acc.c[i] = c[i];
}

template <int BM>
__device__ __forceinline__ void affine_apply(const Affine<BM>& e, double (&m)[BM]) {
  double r[BM];
#pragma unroll
  for (int i = 0; i < BM; ++i) {
    double s = e.c[i];
#pragma unroll
    for (int j = 0; j < BM; ++j) s = fma(e.F[i + j * BM], m[j], s);
    r[i] = s;
  }
#pragma unroll
  for (int i = 0; i < BM; ++i) m[i] = r[i];
}

// Steps are numbered 0..nsteps-1 in PROCESSING order; step s maps to time
// k = kfirst + s (forward) or kfirst - s (backward).
// Phase 1: one thread per (chunk, block) composes its chunk's affine maps.
// grid = (ceil(nchunks / blockDim.y), B), block = (32, CH).
template <int BM, class Elem, int DIR>
__global__ void affine_reduce_kernel(const DevProblem* __restrict__ probs, const DevState* __restrict__ states,
                                     long long kfirst, long long nsteps, double* __restrict__ chunk_buf) {
  const DevProblem& P = probs[blockIdx.y];
  const DevState& St = states[blockIdx.y];
  const int n = threadIdx.x;
  const long long nchunks = (nsteps + kScanChunk - 1) / kScanChunk;
  const long long c = (long long)blockIdx.x * blockDim.y + threadIdx.y;
  if (n >= P.M || c >= nchunks) return;
  Elem el(P, St, n);
  Affine<BM> acc, e;
  const long long s0 = c * kScanChunk;
  const long long s1 = (s0 + kScanChunk < nsteps) ? s0 + kScanChunk : nsteps;
  el.get(kfirst + DIR * s0, acc, false);
  for (long long s = s0 + 1; s < s1; ++s) {
    el.get(kfirst + DIR * s, e, false);
    affine_compose<BM>(acc, e);
  }
  constexpr int W = BM * BM + BM;
  double* dst = chunk_buf + (((size_t)blockIdx.y * nchunks + c) * P.M + n) * W;
#pragma unroll
  for (int i = 0; i < BM * BM; ++i) dst[i] = acc.F[i];
#pragma unroll
  for (int i = 0; i < BM; ++i) dst[BM * BM + i] = acc.c[i];
}

// Phase 2: one warp per signal walks the chunk aggregates, recording the state
// entering each chunk.  init: 0 = St.mcarry (filter), 1 = MS[kinit] (smoother).
template <int BM>
__global__ void __launch_bounds__(32)
affine_carry_kernel(const DevProblem* __restrict__ probs, const DevState* __restrict__ states,
                    long long nsteps, int init, long long kinit, const double* __restrict__ chunk_buf,
                    double* __restrict__ start_buf) {
  const DevProblem& P = probs[blockIdx.x];
  const DevState& St = states[blockIdx.x];
  const int n = threadIdx.x;
  if (n >= P.M) return;
  const long long nchunks = (nsteps + kScanChunk - 1) / kScanChunk;
  const int off = P.off[n], b = P.off[n + 1] - off;
  double m[BM];
#pragma unroll
  for (int i = 0; i < BM; ++i)
    m[i] = init ? ((i < b) ? St.MS[kinit * P.n + off + i] : 0.0) : St.mcarry[n * BM + i];
  constexpr int W = BM * BM + BM;
  for (long long c = 0; c < nchunks; ++c) {
    const size_t base = (((size_t)blockIdx.x * nchunks + c) * P.M + n);
    double* st = start_buf + base * BM;
#pragma unroll
    for (int i = 0; i < BM; ++i) st[i] = m[i];
    const double* src = chunk_buf + base * W;
    Affine<BM> e;
#pragma unroll
    for (int i = 0; i < BM * BM; ++i) e.F[i] = src[i];
#pragma unroll
    for (int i = 0; i < BM; ++i) e.c[i] = src[BM * BM + i];
    affine_apply<BM>(e, m);
  }
}

// Phase 3: re-apply the maps inside each chunk from its entering state and emit.
template <int BM, class Elem, int DIR>
__global__ void affine_apply_kernel(const DevProblem* __restrict__ probs, const DevState* __restrict__ states,
                                    long long kfirst, long long nsteps, const double* __restrict__ start_buf) {
  const DevProblem& P = probs[blockIdx.y];
  const DevState& St = states[blockIdx.y];
  const int n = threadIdx.x;
  const long long nchunks = (nsteps + kScanChunk - 1) / kScanChunk;
  const long long c = (long long)blockIdx.x * blockDim.y + threadIdx.y;
  if (n >= P.M || c >= nchunks) return;
  Elem el(P, St, n);
  double m[BM];
  const double* st = start_buf + (((size_t)blockIdx.y * nchunks + c) * P.M + n) * BM;
#pragma unroll
  for (int i = 0; i < BM; ++i) m[i] = st[i];
  Affine<BM> e;
  const long long s0 = c * kScanChunk;
  const long long s1 = (s0 + kScanChunk < nsteps) ? s0 + kScanChunk : nsteps;
  for (long long s = s0; s < s1; ++s) {
    const long long k = kfirst + DIR * s;
    el.get(k, e, true);
    affine_apply<BM>(e, m);
    el.emit(k, m);
  }
}

// -------------------------------------------------------- site update (EP)
// One thread per time step k in [0, T-1): cavity from the smoothed marginal,
// moments, damped Power-EP update on sites with positive cavity variance
// (ihgp_ep_modulator_nmf.m:397-436; gf_ep_modulator_nmf.m:236-267, :486-510).
// FULL = false: marginal variance from the steady-state table (IHGP);
// FULL = true : marginal variance from the smoothed covariance (St.V).
// clamp_R: the full-state predict mode clamps ttau at 0 and refreshes R for every
// site after the update (:262-265); its nlZ mode and IHGP do not.
// grid = (ceil((T-1)/TPB), B).
template <int DP, int TPB, bool FULL>
__global__ void __launch_bounds__(TPB)
site_update_kernel(const DevProblem* __restrict__ probs, const DevState* __restrict__ states,
                   long long T, double alpha, double ep_damp, int write_lZ, int clamp_R) {
  const DevProblem& P = probs[blockIdx.y];
  const DevState& St = states[blockIdx.y];
  const int tid = threadIdx.x;
  const long long k = (long long)blockIdx.x * TPB + tid;
  const int M = P.M, nr = P.nr;

  extern __shared__ double sm[];
  double* s_mu = sm;                       // [M][TPB]
  double* s_s2 = s_mu + M * TPB;           // [M][TPB]
  double* s_d1 = s_s2 + M * TPB;
  double* s_d2 = s_d1 + M * TPB;
  double* s_W = s_d2 + M * TPB;            // [DP*kNP]
  double* s_wn = s_W + DP * kNP;
  double* s_xn = s_wn + P.S;
  for (int i = tid; i < DP * kNP; i += TPB) s_W[i] = P.W[i];
  for (int i = tid; i < P.S; i += TPB) s_wn[i] = P.wn[i];
  for (int i = tid; i < kNP * P.S; i += TPB) s_xn[i] = P.xn[i];
  __syncthreads();
  if (k >= T - 1) return;
  const double y = St.y[k];
  if (isnan(y)) {                          // ihgp :398, gf_ep :237
    if (!FULL && write_lZ) St.lZ[k] = 0.0; // IHGP accumulates a scalar: no term for this step
    return;
  }
  const MomParams mp = make_mom_params(P, s_W, s_wn, s_xn);
  for (int n = 0; n < M; ++n) {
    double vm;
    if (FULL) {
      vm = St.V[k * M + n];
    } else {
      const int idx = lookup_smoother(P.r, P.thr, nr, St.R[k * M + n]);
      vm = P.vmtab[(size_t)n * nr + idx];
    }
    const double mm = St.E[k * M + n];
    const double vcav = 1.0 / (1.0 / vm - alpha * St.ttau[k * M + n]);      // ihgp :407
    const double mcav = vcav * (mm / vm - alpha * St.tnu[k * M + n]);       // ihgp :408
    s_mu[n * TPB + tid] = mcav;
    s_s2[n * TPB + tid] = vcav;
  }
  const double lz = mom_thread<DP>(mp, alpha, y, s_mu + tid, s_s2 + tid, TPB, s_d1 + tid, s_d2 + tid);
  if (write_lZ) St.lZ[k] = lz;             // ihgp :420 only from the second iteration on
  int neg = 0;
  const double keep = 1.0 - ep_damp * alpha;
  for (int n = 0; n < M; ++n) {
    const double vcav = s_s2[n * TPB + tid];
    double tt = St.ttau[k * M + n];
    const bool upd = vcav > 0.0;           // ihgp :411
    if (upd) {
      const double mcav = s_mu[n * TPB + tid];
      const double d1 = s_d1[n * TPB + tid], d2 = s_d2[n * TPB + tid];
      const double den = 1.0 + d2 * vcav;
      tt = keep * tt + ep_damp * (-d2 / den);                                            // :428
      const double tn = keep * St.tnu[k * M + n] + ep_damp * ((d1 - mcav * d2) / den);   // :430
      St.tnu[k * M + n] = tn;
    } else {
      ++neg;
    }
    if (clamp_R) {
      tt = fmax(tt, 0.0);                  // gf_ep :262
      St.ttau[k * M + n] = tt;
      St.R[k * M + n] = 1.0 / tt;          // gf_ep :265
    } else if (upd) {
      St.ttau[k * M + n] = tt;
      if (!FULL) St.R[k * M + n] = 1.0 / tt;   // ihgp :434 (no clamp in the smoother pass)
    }
  }
  if (neg) atomicAdd(St.negcav, (unsigned long long)neg);
}

// Deterministic sum of x[0..n) into out[0] (negated if neg): one CTA per signal.
__global__ void __launch_bounds__(1024)
sum_kernel(const DevState* __restrict__ states, long long n, double* __restrict__ out, int stride, int slot, int neg) {
  const double* x = states[blockIdx.x].lZ;
  __shared__ double sh[1024];
  double acc = 0.0;
  for (long long i = threadIdx.x; i < n; i += 1024) acc += x[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 512; s >= 1; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[(size_t)blockIdx.x * stride + slot] = neg ? -sh[0] : sh[0];
}

// mcarry <- MS[0] (the reference's `m` carries from the smoother into the next
// filter pass: ihgp_ep_modulator_nmf.m:198 sits outside the EP loop).
template <int BM>
__global__ void carry_mean_kernel(const DevProblem* __restrict__ probs, const DevState* __restrict__ states) {
  const DevProblem& P = probs[blockIdx.x];
  const DevState& St = states[blockIdx.x];
  const int n = threadIdx.x;
  if (n >= P.M) return;
  const int off = P.off[n], b = P.off[n + 1] - off;
  for (int i = 0; i < BM; ++i) St.mcarry[n * BM + i] = (i < b) ? St.MS[off + i] : 0.0;
}

}  // namespace nsagp
