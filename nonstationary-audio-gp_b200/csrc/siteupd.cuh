// Smoother-side Power-EP site update, four lanes per time step.
//
// Reference: matlab/ihgp_ep_modulator_nmf.m:397-436, matlab/gf_ep_modulator_nmf.m:236-267 (predict) / :486-510 (nlZ),
// with the moments of matlab/likModulatorNMFPower.m:28-87 / experiments/likModulatorPreCalcwn.m:28-86.
//
// The steps are independent, so this pass is bound by FP64 throughput -- if the SM is kept busy.  The first form
// (ihgp.cuh: site_update_kernel, one thread per step, 1 + 2 (D+N) accumulators and D coefficients in registers) needs
// 238 registers: 8 warps per SM, FP64 pipe 36 % busy, strided site accesses staged through 39 KB of shared memory.
// Here a CTA of 128 threads owns 32 consecutive steps:
//   phase A  all threads: the steps' rows of E, V / R, ttau, tnu (contiguous in HBM: M doubles per step) are read
//            coalesced, the cavity of every (step, site) is formed in parallel, rows stay in shared memory;
//   phase B  thread (step g, lane dg): subbands dg, dg+4, ... of step g over all S sigma points; the partial sums of
//            v_s and m_s over the four lanes by two xor-shuffles; the link function from a per-step table of the rule's
//            distinct coordinates (common.cuh: DevProblem::ndist) -- N * ndist evaluations instead of N * S;
//   phase C  every lane updates the sites it holds the sums of, in shared memory;
//   phase D  all threads: rows back to HBM, coalesced.
// ~110 registers, 16 warps per SM, no strided global access.  The per-point arithmetic is mom_point<FAST>'s, only the
// sums over subbands are associated differently (four partial sums).
//
// PAIR (used whenever the link table exists): the four lanes of a step take the sigma points two at a time.  In the
// single-point form all four lanes evaluate 1/v, 1/sqrt(v) and the Gaussian density of EVERY point -- 40 % of a lane's
// FP64 work, done four times over.  Here lanes 0,1 finish point A and lanes 2,3 point B (the second shuffle round hands
// each half the partial sums of its own point), the two halves swap the resulting coefficients (c1, c2), and the sums
// that need only the density (Z, the modulators' g-sums) are kept per half and added once per step: the same number of
// shuffles, 20 % fewer FP64 instructions.  The rows of W come from shared memory (broadcast reads) instead of 32
// registers, which pays for the second point's coefficients.
#pragma once
#include "common.cuh"
#include "fastmath.cuh"
#include "mom.cuh"
#include "ihgp.cuh"

namespace nsagp {

constexpr int kSiteSteps = 32;       // steps per CTA
constexpr int kSiteThreads = 4 * kSiteSteps;
constexpr int kSiteRow = 33;         // doubles per step of the cavity rows: odd, so the 8 steps of a warp hit 8 different banks
__host__ __device__ inline int site_tab_stride(int ndist) { return (2 * kNP * ndist) | 1; }   // the same for the link table

// shared-memory doubles
__host__ __device__ inline int site4_smem_doubles(int M, int S, int ndist, bool full) {
  int o = 0;
  o += (full ? 5 : 4) * kSiteSteps * M;                 // E, tt, tn, R (ihgp) | V, Rout (full)
  o += 2 * kSiteSteps * kSiteRow;                       // cavity mean / variance, one padded row per step
  o += 2 * kSiteSteps * kNP;                            // sd, 1/s2 of the modulators
  o += 2 * kSiteSteps;                                  // y, lZ
  o += S + kNP * S;                                     // wn, xn
  o += ndist > 0 ? kSiteSteps * site_tab_stride(ndist) : 0;   // link table
  o += (kNP * S + 7) / 8;                               // index map (bytes)
  o += 32 * kNP;                                        // W rows (PAIR form)
  return o;
}

// Launch bound: three CTAs per SM (154 registers).  Four (128 registers) measured the same on the infinite-horizon path
// and 15 % slower on the full-state path, whose fifth staged array leaves room for three CTAs per SM anyway; five (96
// registers, spills) 30 % slower (profiles/r3g_site.jsonl).
template <int DPT, bool FULL, bool PAIR>
__global__ void __launch_bounds__(kSiteThreads, DPT == 4 ? 3 : 2)
site_update4_kernel(const DevProblem* __restrict__ probs, const DevState* __restrict__ states, long long k0, long long k1,
                    double alpha, double ep_damp, int write_lZ, int clamp_R) {
  const DevProblem& P = probs[blockIdx.y];
  const DevState& St = states[blockIdx.y];
  const int tid = threadIdx.x, lane = tid & 31;
  const int g = tid >> 2, dg = tid & 3;
  const int M = P.M, D = P.D, N = P.N, S = P.S, nd = P.ndist, nr = P.nr;
  const long long kb = k0 + (long long)blockIdx.x * kSiteSteps;
  const int ns = (int)((k1 - kb < kSiteSteps) ? (k1 - kb) : kSiteSteps);
  if (ns <= 0) return;

  extern __shared__ double sm[];
  double* s_E = sm;
  double* s_tt = s_E + kSiteSteps * M;
  double* s_tn = s_tt + kSiteSteps * M;
  double* s_R = s_tn + kSiteSteps * M;                  // ihgp: R (in/out); full: V (in)
  double* s_Ro = s_R + kSiteSteps * M;                  // full: R (out)
  double* s_mu = s_Ro + (FULL ? kSiteSteps * M : 0);
  double* s_s2 = s_mu + kSiteSteps * kSiteRow;
  double* s_sd = s_s2 + kSiteSteps * kSiteRow;
  double* s_rs2 = s_sd + kSiteSteps * kNP;
  double* s_y = s_rs2 + kSiteSteps * kNP;
  double* s_lz = s_y + kSiteSteps;
  double* s_wn = s_lz + kSiteSteps;
  double* s_xn = s_wn + S;
  double* s_tab = s_xn + kNP * S;                       // [step][kNP][nd][2] = (x, link(x))
  unsigned char* s_xi = reinterpret_cast<unsigned char*>(s_tab + (nd > 0 ? kSiteSteps * site_tab_stride(nd) : 0));
  double* s_W = reinterpret_cast<double*>(s_xi) + (kNP * S + 7) / 8;     // [4 * DPT][kNP] (PAIR)

  // ---- phase A: rows in, cavities ------------------------------------------------------------------------------
  for (int i = tid; i < S; i += kSiteThreads) s_wn[i] = P.wn[i];
  for (int i = tid; i < kNP * S; i += kSiteThreads) s_xn[i] = P.xn[i];
  if (nd > 0)
    for (int i = tid; i < kNP * S; i += kSiteThreads) s_xi[i] = P.xidx[i];
  if (PAIR)
    for (int i = tid; i < 4 * DPT * kNP; i += kSiteThreads) s_W[i] = (i / kNP < D) ? P.W[i] : 0.0;
  if (tid < kSiteSteps) s_y[tid] = (tid < ns) ? St.y[kb + tid] : NAN;
  for (int i = tid; i < ns * M; i += kSiteThreads) {
    const int gi = i / M, n = i - gi * M;
    const size_t o = (size_t)kb * M + i;
    const double mm = St.E[o], tt = St.ttau[o], tn = St.tnu[o];
    double vm;
    if (FULL) {
      vm = St.V[o];
      s_R[i] = vm;
    } else {
      const double R = St.R[o];
      s_R[i] = R;
      vm = P.vmtab[(size_t)n * nr + lookup_smoother_hint(P, R)];
    }
    s_E[i] = mm; s_tt[i] = tt; s_tn[i] = tn;
    const double vcav = 1.0 / (1.0 / vm - alpha * tt);                      // ihgp :407, gf_ep :248
    const double mcav = vcav * (mm / vm - alpha * tn);                      // ihgp :408, gf_ep :249
    s_mu[gi * kSiteRow + n] = mcav;
    s_s2[gi * kSiteRow + n] = vcav;
    if (n >= D) {
      s_sd[gi * kNP + n - D] = sqrt(vcav);          // NaN for a negative cavity variance (the reference goes complex)
      s_rs2[gi * kNP + n - D] = 1.0 / vcav;
    }
  }
  __syncthreads();

  // ---- phase B: the sigma-point sums of step g, subbands dg, dg+4, ... --------------------------------------------
  const double y = s_y[g];
  const bool valid = g < ns && !isnan(y);           // ihgp :398, gf_ep :237 (a missing sample is skipped)
  const int jj = dg < N ? dg : 0;
  double muz[DPT], s2z[DPT];
#pragma unroll
  for (int i = 0; i < DPT; ++i) {
    const int d = dg + 4 * i;
    const bool in = d < D;
    muz[i] = (in && valid) ? s_mu[g * kSiteRow + d] : 0.0;
    s2z[i] = (in && valid) ? s_s2[g * kSiteRow + d] : 0.0;
  }
  const double mug = valid ? s_mu[g * kSiteRow + D + jj] : 0.0;
  const double sdg = valid ? s_sd[g * kNP + jj] : 1.0;
  const double rs2g = valid ? s_rs2[g * kNP + jj] : 1.0;
  const double yv = valid ? y : 0.0;
  const double noise = P.sn2 / alpha;
  const double shift = P.link_shift;
  double* tabg = s_tab + (size_t)g * site_tab_stride(nd);
  if (nd > 0) {
    if (dg < N) {
      for (int q = 0; q < nd; ++q) {
        const double x = mug + sdg * P.xdist[jj * nd + q];
        tabg[(jj * nd + q) * 2] = x;
        tabg[(jj * nd + q) * 2 + 1] = softplus_fast(x - shift);
      }
    }
    __syncwarp();
  }
  double a1[DPT], a2[DPT], g1 = 0.0, g2 = 0.0, Zs = 0.0;
#pragma unroll
  for (int i = 0; i < DPT; ++i) { a1[i] = 0.0; a2[i] = 0.0; }

  if constexpr (PAIR) {
    // two points per round: lanes 0,1 of the step finish point A = s, lanes 2,3 point B = s + 1
    const bool hi = (dg & 2) != 0;
    const int j0 = dg & 1, j1 = j0 + 2;             // the modulators whose g-sums this lane keeps (for its half's points)
    const bool has0 = j0 < N, has1 = j1 < N;
    const double mu0 = (valid && has0) ? s_mu[g * kSiteRow + D + j0] : 0.0, rq0 = (valid && has0) ? s_rs2[g * kNP + j0] : 1.0;
    const double mu1 = (valid && has1) ? s_mu[g * kSiteRow + D + j1] : 0.0, rq1 = (valid && has1) ? s_rs2[g * kNP + j1] : 1.0;
    double ga1 = 0.0, ga2 = 0.0, gb1 = 0.0, gb2 = 0.0;
    const double* Wr = s_W + dg * kNP;              // row d = dg + 4 i at Wr + 4 i kNP
    auto links = [&](int s, double (&l)[kNP]) {
#pragma unroll
      for (int j = 0; j < kNP; ++j) l[j] = (j < N) ? tabg[(j * nd + s_xi[j * S + s]) * 2 + 1] : 0.0;
    };
    auto coeffs = [&](const double (&l)[kNP], double (&a)[DPT]) {
#pragma unroll
      for (int i = 0; i < DPT; ++i) {
        double ad = 0.0;
#pragma unroll
        for (int j = 0; j < kNP; ++j) ad = fma(l[j], Wr[(4 * i) * kNP + j], ad);
        a[i] = ad;
      }
      if (P.lik_kind == 1) {
#pragma unroll
        for (int i = 0; i < DPT; ++i) a[i] = sqrt_fast2(a[i]);
      }
    };
    for (int s = 0; s < S; s += 2) {
      const bool hasB = s + 1 < S;
      double aA[DPT], aB[DPT];
      {
        double l[kNP];
        links(s, l);
        coeffs(l, aA);
        links(hasB ? s + 1 : s, l);
        coeffs(l, aB);
      }
      double vsA = 0.0, msA = 0.0, vsB = 0.0, msB = 0.0;
#pragma unroll
      for (int i = 0; i < DPT; ++i) {
        vsA = fma(aA[i] * aA[i], s2z[i], vsA);
        msA = fma(aA[i], muz[i], msA);
        vsB = fma(aB[i] * aB[i], s2z[i], vsB);
        msB = fma(aB[i], muz[i], msB);
      }
      vsA += __shfl_xor_sync(0xffffffffu, vsA, 1);
      msA += __shfl_xor_sync(0xffffffffu, msA, 1);
      vsB += __shfl_xor_sync(0xffffffffu, vsB, 1);
      msB += __shfl_xor_sync(0xffffffffu, msB, 1);
      // second round: each half keeps its own point and receives the other half's partial of it
      const double vs = (hi ? vsB : vsA) + __shfl_xor_sync(0xffffffffu, hi ? vsA : vsB, 2);
      const double ms = (hi ? msB : msA) + __shfl_xor_sync(0xffffffffu, hi ? msA : msB, 2);
      const int so = (hi && hasB) ? s + 1 : s;
      const double v = noise + vs;
      const double rv = rcp_fast2(v);
      const double rsd = rsqrt_fast2(v);
      const double res = yv - ms;
      const double t = res * rsd;
      const double pdf = exp_fast(-0.5 * (t * t)) * (rsd * kInvSqrt2Pi);
      const double wp = ((hi && !hasB) ? 0.0 : s_wn[so]) * pdf;
      const double q = res * rv;
      const double c1 = wp * q;
      const double c2 = wp * (q * q - rv);
      const double c1o = __shfl_xor_sync(0xffffffffu, c1, 2);
      const double c2o = __shfl_xor_sync(0xffffffffu, c2, 2);
      const double c1A = hi ? c1o : c1, c1B = hi ? c1 : c1o;
      const double c2A = hi ? c2o : c2, c2B = hi ? c2 : c2o;
#pragma unroll
      for (int i = 0; i < DPT; ++i) {
        a1[i] = fma(aA[i], c1A, a1[i]);
        a1[i] = fma(aB[i], c1B, a1[i]);
        a2[i] = fma(aA[i] * aA[i], c2A, a2[i]);
        a2[i] = fma(aB[i] * aB[i], c2B, a2[i]);
      }
      Zs += wp;
      if (has0) {
        const double e = (tabg[(j0 * nd + s_xi[j0 * S + so]) * 2] - mu0) * rq0;
        ga1 = fma(wp, e, ga1);
        ga2 = fma(wp, e * e - rq0, ga2);
      }
      if (has1) {
        const double e = (tabg[(j1 * nd + s_xi[j1 * S + so]) * 2] - mu1) * rq1;
        gb1 = fma(wp, e, gb1);
        gb2 = fma(wp, e * e - rq1, gb2);
      }
    }
    // the two halves hold the sums over their own points
    Zs += __shfl_xor_sync(0xffffffffu, Zs, 2);
    ga1 += __shfl_xor_sync(0xffffffffu, ga1, 2);
    ga2 += __shfl_xor_sync(0xffffffffu, ga2, 2);
    gb1 += __shfl_xor_sync(0xffffffffu, gb1, 2);
    gb2 += __shfl_xor_sync(0xffffffffu, gb2, 2);
    g1 = hi ? gb1 : ga1;                            // lane dg updates modulator dg = j0 + 2 (dg >> 1)
    g2 = hi ? gb2 : ga2;
  } else {
  double W[DPT][kNP];
#pragma unroll
  for (int i = 0; i < DPT; ++i) {
    const int d = dg + 4 * i;
    const bool in = d < D;
#pragma unroll
    for (int j = 0; j < kNP; ++j) W[i][j] = in ? P.W[d * kNP + j] : 0.0;
  }
  double xq = 0.0, lq[kNP];
  auto fetch = [&](int s, double& x, double (&l)[kNP]) {
    if (nd > 0) {
#pragma unroll
      for (int j = 0; j < kNP; ++j) l[j] = (j < N) ? tabg[(j * nd + s_xi[j * S + s]) * 2 + 1] : 0.0;
      x = tabg[(jj * nd + s_xi[jj * S + s]) * 2];
    } else {
      x = mug + sdg * s_xn[jj * S + s];
      const double lj = softplus_fast(x - shift);
#pragma unroll
      for (int j = 0; j < kNP; ++j) {
        const double v = __shfl_sync(0xffffffffu, lj, (lane & ~3) | j);
        l[j] = (j < N) ? v : 0.0;
      }
    }
  };
  fetch(0, xq, lq);
  for (int s = 0; s < S; ++s) {
    const double xj = xq;
    double l[kNP];
#pragma unroll
    for (int j = 0; j < kNP; ++j) l[j] = lq[j];
    if (s + 1 < S) fetch(s + 1, xq, lq);            // the next point's entries arrive while this one is integrated
    double a[DPT];
#pragma unroll
    for (int i = 0; i < DPT; ++i) {
      double ad = 0.0;
#pragma unroll
      for (int j = 0; j < kNP; ++j) ad = fma(l[j], W[i][j], ad);
      a[i] = ad;
    }
    if (P.lik_kind == 1) {
#pragma unroll
      for (int i = 0; i < DPT; ++i) a[i] = sqrt_fast2(a[i]);
    }
    double vs = 0.0, ms = 0.0;
#pragma unroll
    for (int i = 0; i < DPT; ++i) {
      vs = fma(a[i] * a[i], s2z[i], vs);
      ms = fma(a[i], muz[i], ms);
    }
    vs += __shfl_xor_sync(0xffffffffu, vs, 1);
    ms += __shfl_xor_sync(0xffffffffu, ms, 1);
    vs += __shfl_xor_sync(0xffffffffu, vs, 2);
    ms += __shfl_xor_sync(0xffffffffu, ms, 2);
    const double v = noise + vs;
    const double rv = rcp_fast2(v);
    const double rsd = rsqrt_fast2(v);
    const double res = yv - ms;
    const double t = res * rsd;
    const double pdf = exp_fast(-0.5 * (t * t)) * (rsd * kInvSqrt2Pi);
    const double wp = s_wn[s] * pdf;
    const double q = res * rv;
    const double c1 = wp * q;
    const double c2 = wp * (q * q - rv);
    Zs += wp;
#pragma unroll
    for (int i = 0; i < DPT; ++i) {
      a1[i] = fma(a[i], c1, a1[i]);
      a2[i] = fma(a[i] * a[i], c2, a2[i]);
    }
    const double e = (xj - mug) * rs2g;
    g1 = fma(wp, e, g1);
    g2 = fma(wp, e * e - rs2g, g2);
  }
  }

  // ---- phase C: moments -> damped Power-EP update of the sites this lane holds -----------------------------------
  const double pep = pep_const(P.lik_kind, P.sn2, alpha);
  const double Zc = pep * fmax(Zs, kJitter);        // fmax(NaN, jitter) = jitter, as MATLAB max
  const double zp = (1.0 / Zc) * pep;
  const double keep = 1.0 - ep_damp * alpha;
  int neg = 0;
  auto update = [&](int n, double r1, double r2) {
    const double d1 = zp * r1;
    const double d2 = -d1 * d1 + zp * r2;
    const int o = g * M + n;
    const double vcav = s_s2[g * kSiteRow + n];
    double tt = s_tt[o];
    const bool upd = vcav > 0.0;                    // ihgp :411
    if (upd) {
      const double mcav = s_mu[g * kSiteRow + n];
      const double den = 1.0 + d2 * vcav;
      tt = keep * tt + ep_damp * (-d2 / den);                                  // :428
      s_tn[o] = keep * s_tn[o] + ep_damp * ((d1 - mcav * d2) / den);           // :430
    } else {
      ++neg;
    }
    if (clamp_R) {
      tt = fmax(tt, 0.0);                           // gf_ep :262
      s_tt[o] = tt;
      if (FULL) s_Ro[o] = 1.0 / tt;                 // gf_ep :265
    } else if (upd) {
      s_tt[o] = tt;
      if (!FULL) s_R[o] = 1.0 / tt;                 // ihgp :434 (no clamp in the smoother pass)
    }
  };
  if (valid) {
#pragma unroll
    for (int i = 0; i < DPT; ++i) {
      const int d = dg + 4 * i;
      if (d < D) update(d, a1[i], a2[i]);
    }
    if (dg < N) update(D + dg, g1, g2);
    if (dg == 0) s_lz[g] = log(Zc);
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) neg += __shfl_xor_sync(0xffffffffu, neg, off);
  if (lane == 0 && neg) atomicAdd(St.negcav, (unsigned long long)neg);
  __syncthreads();

  // ---- phase D: rows out ---------------------------------------------------------------------------------------------
  for (int i = tid; i < ns * M; i += kSiteThreads) {
    const int gi = i / M;
    if (isnan(s_y[gi])) continue;                   // a skipped step changes nothing
    const size_t o = (size_t)kb * M + i;
    St.ttau[o] = s_tt[i];
    St.tnu[o] = s_tn[i];
    if (FULL) { if (clamp_R) St.R[o] = s_Ro[i]; }
    else St.R[o] = s_R[i];
  }
  if (write_lZ && tid < ns) {
    if (!isnan(s_y[tid])) St.lZ[kb + tid] = s_lz[tid];
    else if (!FULL) St.lZ[kb + tid] = 0.0;          // IHGP accumulates a scalar: no term for this step
  }
}

}  // namespace nsagp
