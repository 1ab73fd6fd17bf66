// Tilted-distribution moments of the modulator-NMF likelihood by sigma-point
// integration -- the compute-dominant piece of every EP step.
//
// Follows matlab/likModulatorNMFPower.m:28-87 (kind 0) and
// matlab/experiments/likModulatorPreCalcwn.m:28-86 (kind 1): subbands z are
// integrated analytically, the N modulators g by S sigma points
//   xn = mu_g + sqrt(s2_g) .* xi_s                     (:34)
//   a_s = link(xn) W'   [kind 1: sqrt of that]         (:44)
//   v_s = sn2/alpha + a_s.^2 s2_z,  m_s = a_s mu_z     (:45-46)
//   Z   = pEP max(sum_s w_s N(y; m_s, v_s), 1e-10)     (:51-55)
// and first/second derivatives of log Z w.r.t. all D+N cavity means (:59-80).
//
// Two device forms share the per-point arithmetic:
//   mom_thread : one thread owns a whole time step (S points in sequence, all
//                accumulators in registers) -- used by the passes that are
//                parallel over time.
//   mom_warp   : one warp owns a time step, lanes over sigma points, sums by a
//                register-transposing shuffle reduction -- used by the
//                sequential (ADF) pass where latency per step is what matters.
#pragma once
#include "common.cuh"
#include "fastmath.cuh"

namespace nsagp {

constexpr int kNP = 4;           // padded number of modulators

// Uniform (per-problem) likelihood parameters as seen by the device code.
struct MomParams {
  int D, N, S, kind;
  double sn2, shift;
  const double* W;               // [DP][kNP] row-major, zero padded
  const double* wn;              // [S]
  const double* xn;              // [kNP][S]
  int nd = 0;                    // link table (thread form): distinct coordinates per modulator, 0 = none
  const double* xd = nullptr;    // [kNP][nd]
  const unsigned char* xi = nullptr;   // [kNP][S]
};

template <int DP>
struct MomAcc {
  double Z;
  double a1[DP], a2[DP];
  double g1[kNP], g2[kNP];
  __device__ __forceinline__ void clear() {
    Z = 0.0;
#pragma unroll
    for (int d = 0; d < DP; ++d) { a1[d] = 0.0; a2[d] = 0.0; }
#pragma unroll
    for (int j = 0; j < kNP; ++j) { g1[j] = 0.0; g2[j] = 0.0; }
  }
};

// Per-call modulator-side constants, uniform over sigma points.
struct MomG {
  double mu[kNP], sd[kNP], rs2[kNP];
};

__device__ __forceinline__ double pep_const(int kind, double sn2, double alpha) {
  // likModulatorNMFPower.m:49 -> 1 ; likModulatorPreCalcwn.m:48
  if (kind == 0) return 1.0;
  return pow(2.0 * 3.14159265358979323846 * sn2, 0.5 * (1.0 - alpha)) * (1.0 / sqrt(alpha));
}

// One sigma point.  muz/s2z are read with a stride so the same code serves the
// warp form (stride 1, broadcast) and the thread form (stride = threads per CTA).
// FAST: the straight-line routines of fastmath.cuh (<= 2 ulp, NaN-propagating variants) instead of
// the library's exp / log / sqrt / division; the warp form keeps the library versions so that the
// two sequential-pass implementations stay numerically independent of each other.
// The per-point arithmetic after the link: x[j] = the modulators' sigma-point coordinates, l[j] = link(x[j]).
template <int DP, bool FAST>
__device__ __forceinline__ void mom_point_xl(MomAcc<DP>& acc, const MomParams& p, const MomG& g, int s, double y, double noise,
                                             const double* muz, const double* s2z, int stride, const double (&x)[kNP],
                                             const double (&l)[kNP]);

template <int DP, bool FAST = false>
__device__ __forceinline__ void mom_point(MomAcc<DP>& acc, const MomParams& p, const MomG& g, int s,
                                          double y, double noise, const double* muz,
                                          const double* s2z, int stride) {
  double x[kNP], l[kNP];
#pragma unroll
  for (int j = 0; j < kNP; ++j) {
    if (j < p.N) {
      x[j] = g.mu[j] + g.sd[j] * p.xn[j * p.S + s];
      l[j] = FAST ? softplus_fast(x[j] - p.shift) : log(1.0 + exp(x[j] - p.shift));   // literal link, not log1p (parity)
    } else {
      x[j] = 0.0; l[j] = 0.0;
    }
  }
  mom_point_xl<DP, FAST>(acc, p, g, s, y, noise, muz, s2z, stride, x, l);
}

template <int DP, bool FAST>
__device__ __forceinline__ void mom_point_xl(MomAcc<DP>& acc, const MomParams& p, const MomG& g, int s, double y, double noise,
                                             const double* muz, const double* s2z, int stride, const double (&x)[kNP],
                                             const double (&l)[kNP]) {
  double a[DP];
  double vs = 0.0, ms = 0.0;
#pragma unroll
  for (int d = 0; d < DP; ++d) {
    double ad = 0.0;
    if (d < p.D) {
#pragma unroll
      for (int j = 0; j < kNP; ++j) ad = fma(l[j], p.W[d * kNP + j], ad);
    }
    a[d] = ad;
  }
  if (p.kind == 1) {               // one branch around all square roots: they interleave instead of queueing up
#pragma unroll
    for (int d = 0; d < DP; ++d)
      if (d < p.D) a[d] = FAST ? sqrt_fast2(a[d]) : sqrt(a[d]);
  }
#pragma unroll
  for (int d = 0; d < DP; ++d) {
    if (d < p.D) {
      vs = fma(a[d] * a[d], s2z[d * stride], vs);
      ms = fma(a[d], muz[d * stride], ms);
    }
  }
  const double v = noise + vs;
  const double rv = FAST ? rcp_fast2(v) : 1.0 / v;
  const double rsd = FAST ? rsqrt_fast2(v) : rsqrt(v);
  const double res = y - ms;
  const double t = res * rsd;
  const double pdf = (FAST ? exp_fast(-0.5 * (t * t)) : exp(-0.5 * (t * t))) * (rsd * kInvSqrt2Pi);
  const double wp = p.wn[s] * pdf;
  const double q = res * rv;
  const double c1 = wp * q;
  const double c2 = wp * (q * q - rv);
  acc.Z += wp;
#pragma unroll
  for (int d = 0; d < DP; ++d) {
    if (d < p.D) {
      acc.a1[d] = fma(a[d], c1, acc.a1[d]);
      acc.a2[d] = fma(a[d] * a[d], c2, acc.a2[d]);
    }
  }
#pragma unroll
  for (int j = 0; j < kNP; ++j) {
    if (j < p.N) {
      const double e = (x[j] - g.mu[j]) * g.rs2[j];
      acc.g1[j] = fma(wp, e, acc.g1[j]);
      acc.g2[j] = fma(wp, e * e - g.rs2[j], acc.g2[j]);
    }
  }
}

__device__ __forceinline__ void mom_setup_g(MomG& g, const MomParams& p, const double* mug,
                                            const double* s2g, int stride) {
#pragma unroll
  for (int j = 0; j < kNP; ++j) {
    if (j < p.N) {
      const double s2 = s2g[j * stride];
      g.mu[j] = mug[j * stride];
      g.sd[j] = sqrt(s2);            // NaN for a negative cavity variance (reference goes complex)
      g.rs2[j] = 1.0 / s2;
    } else {
      g.mu[j] = 0.0; g.sd[j] = 0.0; g.rs2[j] = 0.0;
    }
  }
}

// ---------------------------------------------------------------- thread form
// mu/s2 hold the D+N cavity means/variances of this thread's step with the given
// stride.  Results are written to d1/d2 with the same stride; returns lZ.
// tab (optional, with p.nd > 0): this thread's scratch for the link table, 2 * kNP * p.nd doubles at the given stride.
template <int DP>
__device__ __forceinline__ double mom_thread(const MomParams& p, double alpha, double y,
                                             const double* mu, const double* s2, int stride,
                                             double* d1, double* d2, double* tab = nullptr) {
  MomAcc<DP> acc;
  acc.clear();
  MomG g;
  mom_setup_g(g, p, mu + p.D * stride, s2 + p.D * stride, stride);
  const double noise = p.sn2 / alpha;
  if (p.nd <= 0) tab = nullptr;
  if (tab) {
    // Link table of this step: the coordinates of a symmetric / tensor rule take p.nd distinct values per modulator,
    // so link() is evaluated N * nd times instead of N * S times; entry (j, q) = (x, link(x)).  The entries of point
    // s + 1 are fetched while point s is integrated (the fetch is two dependent shared-memory loads).
    for (int j = 0; j < p.N; ++j)
      for (int q = 0; q < p.nd; ++q) {
        const double x = g.mu[j] + g.sd[j] * p.xd[j * p.nd + q];
        tab[(j * p.nd + q) * 2 * stride] = x;
        tab[(j * p.nd + q) * 2 * stride + stride] = softplus_fast(x - p.shift);
      }
    double xc[kNP], lc[kNP], xq[kNP], lq[kNP];
    auto fetch = [&](int s, double (&x)[kNP], double (&l)[kNP]) {
#pragma unroll
      for (int j = 0; j < kNP; ++j) {
        if (j < p.N) {
          const int q = (j * p.nd + p.xi[j * p.S + s]) * 2 * stride;
          x[j] = tab[q];
          l[j] = tab[q + stride];
        } else {
          x[j] = 0.0; l[j] = 0.0;
        }
      }
    };
    fetch(0, xq, lq);
    for (int s = 0; s < p.S; ++s) {
#pragma unroll
      for (int j = 0; j < kNP; ++j) { xc[j] = xq[j]; lc[j] = lq[j]; }
      if (s + 1 < p.S) fetch(s + 1, xq, lq);
      mom_point_xl<DP, true>(acc, p, g, s, y, noise, mu, s2, stride, xc, lc);
    }
  } else {
    for (int s = 0; s < p.S; ++s) mom_point<DP, true>(acc, p, g, s, y, noise, mu, s2, stride);
  }
  const double pep = pep_const(p.kind, p.sn2, alpha);
  const double Z = pep * fmax(acc.Z, kJitter);      // fmax(NaN, jitter) = jitter, as MATLAB max
  const double zp = (1.0 / Z) * pep;
#pragma unroll
  for (int d = 0; d < DP; ++d) {
    if (d < p.D) {
      const double dl = zp * acc.a1[d];
      d1[d * stride] = dl;
      d2[d * stride] = -dl * dl + zp * acc.a2[d];
    }
  }
#pragma unroll
  for (int j = 0; j < kNP; ++j) {
    if (j < p.N) {
      const double dl = zp * acc.g1[j];
      d1[(p.D + j) * stride] = dl;
      d2[(p.D + j) * stride] = -dl * dl + zp * acc.g2[j];
    }
  }
  return log(Z);
}

// ------------------------------------------------------------------ warp form
// Register-transposing warp reduction: K values per lane in, afterwards lane L
// holds in v[0] the warp total of the original v[L >> (5 - log2 K)].
// K + (5 - log2 K) - 1 double shuffles instead of 5 K.
template <int K>
__device__ __forceinline__ void warp_reduce_array(double (&v)[K], int lane) {
  static_assert(K == 1 || K == 2 || K == 4 || K == 8 || K == 16 || K == 32, "K must be a power of two");
  int off = 16;
#pragma unroll
  for (int half = K / 2; half >= 1; half >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const double send = upper ? v[i] : v[i + half];
      const double keep = upper ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
    off >>= 1;
  }
  for (; off >= 1; off >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
}

// All 32 lanes must call.  mu/s2: D+N values, contiguous (shared memory).
// On return lane n < D+N holds (d1, d2) of site n; every lane gets lZ.
// Requires D <= DP, D+N <= 32, N <= kNP.
template <int DP, bool FAST = false>
__device__ __forceinline__ double mom_warp(const MomParams& p, double alpha, double y,
                                           const double* mu, const double* s2, int lane,
                                           double& d1, double& d2) {
  static_assert(DP == 16 || DP == 32, "DP is 16 or 32");
  MomAcc<DP> acc;
  acc.clear();
  MomG g;
  mom_setup_g(g, p, mu + p.D, s2 + p.D, 1);
  const double noise = p.sn2 / alpha;
  for (int s = lane; s < p.S; s += kWarp) mom_point<DP, FAST>(acc, p, g, s, y, noise, mu, s2, 1);

  // modulator sums: [g1 | g2] -> lane L holds entry L>>2
  double gv[2 * kNP];
#pragma unroll
  for (int j = 0; j < kNP; ++j) { gv[j] = acc.g1[j]; gv[kNP + j] = acc.g2[j]; }
  warp_reduce_array<2 * kNP>(gv, lane);
  double zs = acc.Z;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) zs += __shfl_xor_sync(0xffffffffu, zs, off);

  double r1, r2;   // raw sums for this lane's site
  if (DP == 16) {
    double zv[32];
#pragma unroll
    for (int d = 0; d < 16; ++d) { zv[d] = acc.a1[d]; zv[16 + d] = acc.a2[d]; }
    warp_reduce_array<32>(zv, lane);
    r1 = __shfl_sync(0xffffffffu, zv[0], lane & 15);
    r2 = __shfl_sync(0xffffffffu, zv[0], 16 + (lane & 15));
  } else {
    warp_reduce_array<DP>(acc.a1, lane);
    warp_reduce_array<DP>(acc.a2, lane);
    r1 = acc.a1[0];
    r2 = acc.a2[0];
  }
  int j = lane - p.D;
  j = j < 0 ? 0 : (j >= kNP ? kNP - 1 : j);
  const double q1 = __shfl_sync(0xffffffffu, gv[0], 4 * j);
  const double q2 = __shfl_sync(0xffffffffu, gv[0], 4 * (kNP + j));
  if (lane >= p.D) { r1 = q1; r2 = q2; }

  const double pep = pep_const(p.kind, p.sn2, alpha);
  const double Z = pep * fmax(zs, kJitter);
  const double zp = (1.0 / Z) * pep;
  d1 = zp * r1;
  d2 = -d1 * d1 + zp * r2;
  return log(Z);
}

}  // namespace nsagp
