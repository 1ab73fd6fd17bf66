// CTA-cooperative moment matching for the sequential (ADF) passes.
//
// The ADF pass is a nonlinear recurrence: step k cannot start before step k-1
// has produced its posterior mean, so what matters is the LATENCY of one
// likModulatorNMFPower evaluation (likModulatorNMFPower.m:28-87), not
// throughput.  A single warp needs three rounds of 32 sigma points with ~2500
// dependent-heavy FP64 instructions each.  Here one CTA owns the signal and the
// S x D work of a step is spread over 4*S threads:
//
//   thread (s, dg): sigma point s, subband group dg in 0..3 (d = dg, dg+4, ...)
//   - link: lanes dg < N evaluate softplus for modulator j = dg, the 4 lanes of a
//     point exchange them by shuffle (no redundancy, no shared memory);
//   - a(s,d), v_s and m_s partials over the thread's DPT subbands, completed over
//     the 4 lanes with two xor-shuffles;
//   - pdf / weights once per point (computed by all 4 lanes: SIMT makes it free);
//   - sums over s: xor-shuffles over the 8 points of a warp, then shared memory
//     across warps (two barriers).
//
// When 4*S <= blockDim.x (SINGLE) every thread owns the same sigma point in every
// step, so its abscissa, weight and rows of W live in registers for the whole pass.
#pragma once
#include "common.cuh"
#include "mom.cuh"

namespace nsagp {

template <int DPT>
struct MomCta {
  static constexpr int NV = 2 * DPT + 3;          // per-thread sums: a1[DPT] a2[DPT] g1 g2 z
  // shared-memory doubles needed besides mu/s2: partials [nwarps][4][NV] + finals [4][NV]
  __host__ __device__ static constexpr int smem_doubles(int nwarps) { return (nwarps + 1) * 4 * NV; }
};

// Per-thread constants (fixed for the whole pass).
template <int DPT>
struct MomCtaThread {
  double W[DPT][kNP];     // rows dg, dg+4, ... of W
  double xn0, wn0;        // SINGLE: this thread's abscissa (modulator jj) and weight (0 if inactive)
  int dg, jj;
  __device__ __forceinline__ void init(const MomParams& p, int tid) {
    dg = tid & 3;
    jj = dg < p.N ? dg : 0;
#pragma unroll
    for (int i = 0; i < DPT; ++i) {
      const int d = dg + 4 * i;
#pragma unroll
      for (int j = 0; j < kNP; ++j) W[i][j] = (d < p.D) ? p.W[d * kNP + j] : 0.0;
    }
    const int s = tid >> 2;
    const bool act = s < p.S;
    xn0 = act ? p.xn[jj * p.S + s] : 0.0;
    wn0 = act ? p.wn[s] : 0.0;
  }
};

// All threads of the CTA must call (contains two __syncthreads).  blockDim.x is a
// multiple of 32.  mu/s2: D+N cavity values in shared memory, written by the caller
// and made visible by a barrier before the call.  On return s_fin (4*NV doubles)
// holds the raw sums; use mom_cta_result() to read one site's derivatives.
template <int DPT, bool SINGLE>
__device__ __forceinline__ void mom_cta(const MomParams& p, const MomCtaThread<DPT>& th, double alpha, double y,
                                        const double* mu, const double* s2, double* s_part, double* s_fin) {
  constexpr int NV = MomCta<DPT>::NV;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x;
  const int dg = th.dg, jj = th.jj;
  const double noise = p.sn2 / alpha;
  double acc[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) acc[v] = 0.0;

  const double mug = mu[p.D + jj];
  const double s2g = s2[p.D + jj];
  const double sdg = sqrt(s2g);                     // NaN for a negative cavity variance (reference goes complex)
  const double rs2g = 1.0 / s2g;
  // this thread's subband cavity values
  double muz[DPT], s2z[DPT];
#pragma unroll
  for (int i = 0; i < DPT; ++i) {
    const int d = dg + 4 * i;
    const bool in = d < p.D;
    muz[i] = in ? mu[d] : 0.0;
    s2z[i] = in ? s2[d] : 0.0;
  }
  const int items = 4 * p.S;
  const int rounds = SINGLE ? 1 : (items + nthreads - 1) / nthreads;
  for (int r = 0; r < rounds; ++r) {
    const int it = r * nthreads + tid;
    const bool act = it < items;
    const int s = act ? (it >> 2) : 0;
    // link (one modulator per lane), then exchange within the 4 lanes of the point
    const double xi = SINGLE ? th.xn0 : p.xn[jj * p.S + s];
    const double xj = mug + sdg * xi;
    const double lj = (dg < p.N) ? log(1.0 + exp(xj - p.shift)) : 0.0;     // literal link (parity)
    double l[kNP];
#pragma unroll
    for (int j = 0; j < kNP; ++j) l[j] = __shfl_sync(0xffffffffu, lj, (lane & ~3) | j);
    double a[DPT];
    double vs = 0.0, ms = 0.0;
#pragma unroll
    for (int i = 0; i < DPT; ++i) {
      double ad = 0.0;
#pragma unroll
      for (int j = 0; j < kNP; ++j) ad = fma(l[j], th.W[i][j], ad);
      if (p.kind == 1) ad = sqrt(ad);
      a[i] = ad;                                     // padded rows of W are zero -> a = 0
      vs = fma(ad * ad, s2z[i], vs);
      ms = fma(ad, muz[i], ms);
    }
    vs += __shfl_xor_sync(0xffffffffu, vs, 1);
    ms += __shfl_xor_sync(0xffffffffu, ms, 1);
    vs += __shfl_xor_sync(0xffffffffu, vs, 2);
    ms += __shfl_xor_sync(0xffffffffu, ms, 2);
    const double v = noise + vs;
    const double rv = 1.0 / v;
    const double rsd = rsqrt(v);
    const double res = y - ms;
    const double t = res * rsd;
    const double pdf = exp(-0.5 * (t * t)) * (rsd * kInvSqrt2Pi);
    const double wgt = SINGLE ? th.wn0 : (act ? p.wn[s] : 0.0);
    const double wp = act ? wgt * pdf : 0.0;         // inactive threads contribute exact zeros
    const double q = res * rv;
    const double c1 = wp * q;
    const double c2 = wp * (q * q - rv);
    if (act) {
#pragma unroll
      for (int i = 0; i < DPT; ++i) {
        acc[i] = fma(a[i], c1, acc[i]);
        acc[DPT + i] = fma(a[i] * a[i], c2, acc[DPT + i]);
      }
      if (dg < p.N) {
        const double e = (xj - mug) * rs2g;
        acc[2 * DPT] = fma(wp, e, acc[2 * DPT]);
        acc[2 * DPT + 1] = fma(wp, e * e - rs2g, acc[2 * DPT + 1]);
      }
      if (dg == 0) acc[2 * DPT + 2] += wp;
    }
  }
  // sums over the 8 points of this warp (lanes with equal dg)
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    double x = acc[v];
    x += __shfl_xor_sync(0xffffffffu, x, 4);
    x += __shfl_xor_sync(0xffffffffu, x, 8);
    x += __shfl_xor_sync(0xffffffffu, x, 16);
    acc[v] = x;
  }
  if (lane < 4) {
#pragma unroll
    for (int v = 0; v < NV; ++v) s_part[(warp * 4 + lane) * NV + v] = acc[v];
  }
  __syncthreads();
  if (tid < 4 * NV) {
    const int nw = nthreads >> 5;
    double x = 0.0;
    for (int w = 0; w < nw; ++w) x += s_part[w * 4 * NV + tid];
    s_fin[tid] = x;
  }
  __syncthreads();
}

// Derivatives for site n from the raw sums (likModulatorNMFPower.m:55-80).  pep is
// pep_const(kind, sn2, alpha), a per-pass constant the caller computes once.
template <int DPT>
__device__ __forceinline__ void mom_cta_result(const MomParams& p, double pep, const double* s_fin, int n,
                                               double& Z, double& d1, double& d2) {
  constexpr int NV = MomCta<DPT>::NV;
  Z = pep * fmax(s_fin[2 * DPT + 2], kJitter);      // fmax(NaN, jitter) = jitter, as MATLAB max
  const double zp = (1.0 / Z) * pep;
  double r1, r2;
  if (n < p.D) {
    const int dg = n & 3, i = n >> 2;
    r1 = s_fin[dg * NV + i];
    r2 = s_fin[dg * NV + DPT + i];
  } else {
    const int j = n - p.D;
    r1 = s_fin[j * NV + 2 * DPT];
    r2 = s_fin[j * NV + 2 * DPT + 1];
  }
  d1 = zp * r1;
  d2 = -d1 * d1 + zp * r2;
}

}  // namespace nsagp
