// CTA-cooperative moment matching for the sequential (ADF) passes.
//
// The ADF pass is a nonlinear recurrence: step k cannot start before step k-1
// has produced its posterior mean, so what matters is the LATENCY of one
// likModulatorNMFPower evaluation (likModulatorNMFPower.m:28-87), not
// throughput.  One CTA owns the signal.  Warp 0 is the "Kalman warp" (lane n owns
// latent block n); the remaining warps are "moment warps" that spread the S x D
// work of a step over 4*S threads:
//
//   thread (s, dg): sigma point s, subband group dg in 0..3 (d = dg, dg+4, ...)
//   - link: lanes dg < N evaluate softplus for modulator j = dg, the 4 lanes of a
//     point exchange them by shuffle (no redundancy, no shared memory);
//   - a(s,d), v_s and m_s partials over the thread's DPT subbands, completed over
//     the 4 lanes with two xor-shuffles;
//   - pdf / weights once per point (computed by all 4 lanes: SIMT makes it free);
//   - y = NaN (missing sample) never reaches these threads: the Kalman warp applies the
//     reference's NaN semantics itself, so everything here is finite and branch-free;
//   - sums over the 8 points of a warp by a register-transposing butterfly
//     (NVP/2 + NVP/4 + NVP/4 shuffles instead of 3 NV), then one shared-memory
//     row per (warp, dg); the Kalman lanes add the rows they need themselves.
//
// The two sides meet at two named barriers per step (producer: bar.arrive,
// consumer: bar.sync), so the Kalman warp's stores and logarithm overlap the
// moment warps' work.  All arithmetic on the critical path is straight-line
// (fastmath.cuh).
//
// When 4*S <= (moment threads) (SINGLE) every thread owns the same sigma point in
// every step, so its abscissa, weight and rows of W live in registers for the pass.
#pragma once
#include "common.cuh"
#include "fastmath.cuh"
#include "mom.cuh"

namespace nsagp {

constexpr int kBarCavity = 1;     // Kalman warp -> moment warps: cavity (mu, s2, sd, 1/s2) is in shared memory
constexpr int kBarSums = 2;       // moment warps -> Kalman warp: partial sums are in shared memory

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int DPT>
struct MomCta {
  static constexpr int NV = 2 * DPT + 3;          // per-thread sums: a1[DPT] a2[DPT] g1 g2 z
  static constexpr int NVP = (NV + 3) & ~3;       // padded to a multiple of 4 for the transposing butterfly
  static constexpr int kCav = 72;                 // mu[32] s2[32] sd[4] rs2[4]
  // shared-memory doubles: cavity + partial rows [moment warps][4][NVP]
  __host__ __device__ static constexpr int smem_doubles(int nmw) { return kCav + nmw * 4 * NVP; }
};

// Per-thread constants of a moment thread (fixed for the whole pass).
template <int DPT>
struct MomCtaThread {
  double W[DPT][kNP];     // rows dg, dg+4, ... of W
  double xn0, wn0;        // SINGLE: this thread's abscissa (modulator jj) and weight (0 if inactive)
  int dg, jj;
  unsigned ltab = 0;      // shared-memory byte address of the logarithm table (fastmath.cuh: log_ge1_tab)
  __device__ __forceinline__ void init(const MomParams& p, int mtid) {
    dg = mtid & 3;
    jj = dg < p.N ? dg : 0;
    // Rows beyond D repeat a real row (not zeros): their cavity entries are zero, so they add
    // nothing to v_s / m_s, and their sums land in slots no Kalman lane reads -- but a = sqrt(.)
    // stays an ordinary positive number, so the straight-line sqrt needs no zero case.
#pragma unroll
    for (int i = 0; i < DPT; ++i) {
      const int d = (dg + 4 * i) % p.D;
#pragma unroll
      for (int j = 0; j < kNP; ++j) W[i][j] = p.W[d * kNP + j];
    }
    const int s = mtid >> 2;
    const bool act = mtid >= 0 && s < p.S;
    xn0 = act ? p.xn[jj * p.S + s] : 0.0;
    wn0 = act ? p.wn[s] : 0.0;
  }
};

// Work of one moment thread for one step.  mtid: index among the moment threads,
// nmt: their number (multiple of 32).  s_cav: cavity written by the Kalman warp
// (visible after kBarCavity).  Writes this warp's partial rows to s_part; the
// caller then arrives at kBarSums.
template <int DPT, bool SINGLE>
__device__ __forceinline__ void mom_cta_points(const MomParams& p, const MomCtaThread<DPT>& th, double noise, double y,
                                               const double* s_cav, double* s_part, int mtid, int nmt) {
  constexpr int NVP = MomCta<DPT>::NVP;
  const int lane = mtid & 31, mw = mtid >> 5;
  const int dg = th.dg, jj = th.jj;
  double acc[NVP];
#pragma unroll
  for (int v = 0; v < NVP; ++v) acc[v] = 0.0;

  const double mug = s_cav[p.D + jj];
  const double sdg = s_cav[64 + jj];                // sqrt(s2_g)  (likModulatorNMFPower.m:34)
  const double rs2g = s_cav[68 + jj];               // 1 / s2_g
  double muz[DPT], s2z[DPT];
#pragma unroll
  for (int i = 0; i < DPT; ++i) {
    const int d = dg + 4 * i;
    const bool in = d < p.D;
    muz[i] = in ? s_cav[d] : 0.0;
    s2z[i] = in ? s_cav[32 + d] : 0.0;
  }
  const int items = 4 * p.S;
  const int rounds = SINGLE ? 1 : (items + nmt - 1) / nmt;
  for (int r = 0; r < rounds; ++r) {
    const int it = r * nmt + mtid;
    const bool act = it < items;
    const int s = act ? (it >> 2) : 0;
    // link (one modulator per lane), then exchange within the 4 lanes of the point
    const double xi = SINGLE ? th.xn0 : p.xn[jj * p.S + s];
    const double xj = fma(sdg, xi, mug);
    // lanes dg >= N repeat modulator 0; their value meets the zero-padded column of W
    const double lj = softplus_tab_finite(xj - p.shift, th.ltab);
    double l[kNP];
#pragma unroll
    for (int j = 0; j < kNP; ++j) l[j] = __shfl_sync(0xffffffffu, lj, (lane & ~3) | j);
    double a[DPT], a2[DPT];
    double vs = 0.0, ms = 0.0;
#pragma unroll
    for (int i = 0; i < DPT; ++i) {
      double ad = 0.0;
#pragma unroll
      for (int j = 0; j < kNP; ++j) ad = fma(l[j], th.W[i][j], ad);
      a[i] = ad;
    }
    // one branch on the likelihood kind around all DPT evaluations (a branch per subband kept the DPT square
    // roots in separate basic blocks, i.e. one after the other on the critical path)
    if (p.kind == 1) {
#pragma unroll
      for (int i = 0; i < DPT; ++i) { a2[i] = a[i]; a[i] = sqrt_fast2_pos(a[i]); }   // a = sqrt(W link): a.^2 is the argument itself
    } else {
#pragma unroll
      for (int i = 0; i < DPT; ++i) a2[i] = a[i] * a[i];
    }
#pragma unroll
    for (int i = 0; i < DPT; ++i) {
      vs = fma(a2[i], s2z[i], vs);
      ms = fma(a[i], muz[i], ms);
    }
    vs += __shfl_xor_sync(0xffffffffu, vs, 1);
    ms += __shfl_xor_sync(0xffffffffu, ms, 1);
    vs += __shfl_xor_sync(0xffffffffu, vs, 2);
    ms += __shfl_xor_sync(0xffffffffu, ms, 2);
    const double v = noise + vs;
    const double rv = rcp_fast2(v);
    const double rsd = rsqrt_fast2(v);
    const double res = y - ms;
    // exp(-(y-m)^2 / (2v)) with the argument formed from 1/v: that reciprocal's chain is two instructions shorter than
    // 1/sqrt(v)'s, which is only needed for the prefactor and finishes while the exponential runs
    const double q = res * rv;
    const double ex = exp_fast_sel((-0.5 * res) * q);
    const double wgt = SINGLE ? th.wn0 : (act ? p.wn[s] : 0.0);
    const double wp = ex * (wgt * (rsd * kInvSqrt2Pi));     // inactive threads: weight 0, finite density
    const double c1 = wp * q;
    const double c2 = wp * fma(q, q, -rv);
#pragma unroll
    for (int i = 0; i < DPT; ++i) {
      acc[i] = fma(a[i], c1, acc[i]);
      acc[DPT + i] = fma(a2[i], c2, acc[DPT + i]);
    }
    // (rows dg >= N of the g-sums and rows dg > 0 of Z are never read)
    const double e = (xj - mug) * rs2g;
    acc[2 * DPT] = fma(wp, e, acc[2 * DPT]);
    acc[2 * DPT + 1] = fma(wp, fma(e, e, -rs2g), acc[2 * DPT + 1]);
    acc[2 * DPT + 2] += wp;
  }
  // Sums over the 8 points of this warp (lanes with equal dg: xor 16, 8, 4).
  // Level 1 and 2 halve the number of live values (each lane keeps the half selected
  // by its own bit and receives the partner's copy of it); level 3 is a plain butterfly.
  constexpr int H = NVP / 2, Qn = NVP / 4;
  {
    const bool up = (lane & 16) != 0;
#pragma unroll
    for (int i = 0; i < H; ++i) {
      const double send = up ? acc[i] : acc[i + H];
      const double keep = up ? acc[i + H] : acc[i];
      acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool up = (lane & 8) != 0;
#pragma unroll
    for (int i = 0; i < Qn; ++i) {
      const double send = up ? acc[i] : acc[i + Qn];
      const double keep = up ? acc[i + Qn] : acc[i];
      acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
#pragma unroll
  for (int i = 0; i < Qn; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 4);
  if ((lane & 4) == 0) {
    // this lane holds entries j0 .. j0+Qn-1 of row (mw, dg)
    const int j0 = ((lane >> 4) & 1) * H + ((lane >> 3) & 1) * Qn;
    double* row = s_part + (mw * 4 + dg) * NVP + j0;
#pragma unroll
    for (int i = 0; i < Qn; ++i) row[i] = acc[i];
  }
}

__device__ __forceinline__ double lds_f64(unsigned addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}

// Kalman lane n: shared-memory byte addresses (first moment warp's rows) of the three sums
// of its site:  Zs = sum_s w_s N_s,  r1 = sum_s w_s N_s [.]',  r2 = sum_s w_s N_s [.]''
// (likModulatorNMFPower.m:55-80).
template <int DPT>
struct MomCtaSumAddr {
  unsigned z, a, b;
  __device__ __forceinline__ void init(const double* s_part, int n, int D) {
    constexpr int NVP = MomCta<DPT>::NVP;
    int dg, j1, j2;
    if (n < D) { dg = n & 3; j1 = n >> 2; j2 = DPT + (n >> 2); }
    else { dg = n - D; j1 = 2 * DPT; j2 = 2 * DPT + 1; }
    const unsigned base = (unsigned)__cvta_generic_to_shared(s_part);
    z = base + 8u * (2 * DPT + 2);                  // row dg = 0
    a = base + 8u * (dg * NVP + j1);
    b = base + 8u * (dg * NVP + j2);
  }
};

// Add the rows of the nmw moment warps (two accumulators per sum: shorter chain).
template <int DPT>
__device__ __forceinline__ void mom_cta_sums(const MomCtaSumAddr<DPT>& ad, int nmw, double& Zs, double& r1, double& r2) {
  constexpr unsigned kRow = 4u * MomCta<DPT>::NVP * 8u;    // bytes between consecutive warps' rows
  double z0 = 0.0, z1 = 0.0, a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
  unsigned o = 0;
  int w = 0;
#pragma unroll 5
  for (; w + 1 < nmw; w += 2, o += 2 * kRow) {
    z0 += lds_f64(ad.z + o); z1 += lds_f64(ad.z + o + kRow);
    a0 += lds_f64(ad.a + o); a1 += lds_f64(ad.a + o + kRow);
    b0 += lds_f64(ad.b + o); b1 += lds_f64(ad.b + o + kRow);
  }
  if (w < nmw) { z0 += lds_f64(ad.z + o); a0 += lds_f64(ad.a + o); b0 += lds_f64(ad.b + o); }
  Zs = z0 + z1; r1 = a0 + a1; r2 = b0 + b1;
}

// Damped ADF site update from the raw sums, with a single reciprocal:
//   dlZ = r1/Zm, d2lZ = -dlZ^2 + r2/Zm, Zm = max(Zs, jitter)           (likModulatorNMFPower.m:55-80)
//   ttau_new = -d2lZ / (1 + d2lZ s2) = -Nn / (Zm^2 + Nn s2),  Nn = r2 Zm - r1^2
//   tnu_new  = (dlZ - mu d2lZ) / (1 + d2lZ s2) = (r1 Zm - mu Nn) / (Zm^2 + Nn s2)
// (ihgp_ep_modulator_nmf.m:265-266, gf_ep_modulator_nmf.m:147-148).
__device__ __forceinline__ void adf_site_from_sums(double Zs, double r1, double r2, double mu, double s2,
                                                   double& tt_new, double& tn_new, double& Zm) {
  Zm = fmax(Zs, kJitter);                           // fmax(NaN, jitter) = jitter, as MATLAB max
  const double Nn = fma(r2, Zm, -r1 * r1);
  const double den = fma(Nn, s2, Zm * Zm);
  const double num_t = fma(r1, Zm, -mu * Nn);
  // straight-line reciprocal first (den is an ordinary number in all but degenerate steps); the range test depends on
  // den alone, so it resolves while the reciprocal's chain runs and costs nothing on it
  const double rd = rcp_fast2(den);
  tt_new = -Nn * rd;
  tn_new = num_t * rd;
  if (!(fabs(den) > 1e-280 && fabs(den) < 1e280)) { // 0, Inf, NaN, subnormal: IEEE division semantics
    tt_new = -Nn / den;
    tn_new = num_t / den;
  }
}

}  // namespace nsagp
