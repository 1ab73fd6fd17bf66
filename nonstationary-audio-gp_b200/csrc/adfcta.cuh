// Sequential (ADF) filter passes, one CTA per signal.
//
// References: matlab/ihgp_ep_modulator_nmf.m:233-310 (infinite-horizon filter),
// matlab/gf_ep_modulator_nmf.m:126-184 / :400-447 (full-state filter).
//
// The pass is a nonlinear recurrence in time, so one step's latency is the whole
// cost.  The CTA splits a step as
//   warp 0, lane n < M : Kalman predict / update of block n (registers);
//   all 4*S threads    : the sigma-point moment matching (momcta.cuh);
// with three CTA barriers per step.  Everything a step touches that does not
// depend on the recurrence (y, old site values) is loaded one step ahead; the
// steady-state tables and the look-up thresholds sit in shared memory, and the
// nearest-neighbour look-up first tries the previous step's row.
#pragma once
#include "common.cuh"
#include "lookup.cuh"
#include "mom.cuh"
#include "momcta.cuh"
#include "ihgp.cuh"
#include "gfep.cuh"

namespace nsagp {

constexpr int kAdfMaxThreads = 384;

// ind = #{i : thr[i] <= R}, trying `hint` first (R moves slowly from step to step).
__device__ __forceinline__ int nearest_by_threshold_hint(const double* thr, int nr, double R, int hint) {
  if (hint >= 0 && hint < nr) {
    const double lo = (hint > 0) ? thr[hint - 1] : -INFINITY;
    const double hi = (hint < nr - 1) ? thr[hint] : INFINITY;
    if (lo <= R && !(hi <= R)) return hint;
  }
  return nearest_by_threshold(thr, nr, R);
}

__device__ __forceinline__ int lookup_filter_hint(const double* r, const double* thr, int nr, double R, int hint) {
  if (!(R > 0.0)) return 0;
  if (isinf(R)) return 0;
  if (R >= kLookupBig) return nearest_bruteforce(r, nr, R);
  return nearest_by_threshold_hint(thr, nr, R, hint);
}

// Shared-memory layout helper (doubles).
struct AdfSmem {
  int mu, s2, part, fin, wn, xn, thr, hph, wtab, total;
  __host__ __device__ AdfSmem(int nwarps, int NV, int S, int M, int nr, int BM, bool tables) {
    int o = 0;
    mu = o; o += 32;
    s2 = o; o += 32;
    part = o; o += nwarps * 4 * NV;
    fin = o; o += 4 * NV;
    wn = o; o += S;
    xn = o; o += kNP * S;
    thr = o; o += tables ? (nr > 0 ? nr - 1 : 0) : 0;
    hph = o; o += tables ? M * (nr + 1) : 0;
    wtab = o; o += tables ? M * (nr + 1) * BM : 0;
    total = o;
  }
};

// ----------------------------------------------------------------- IHGP
// Steps k0..k1-1.  mom_all: moment matching at every step (first pass) or only at
// k == T-1.  running: the _constraints nlZ variant's running site vectors.
// tab_smem: tables copied to shared memory (host checked that they fit).
template <int DPT, int BM, bool SINGLE>
__global__ void __launch_bounds__(kAdfMaxThreads)
ihgp_adf_cta_kernel(const DevProblem* __restrict__ probs, const DevState* __restrict__ states,
                    long long T, long long k0, long long k1, int mom_all, double ep_damp, int running,
                    int tab_smem) {
  constexpr int NV = MomCta<DPT>::NV;
  const DevProblem& P = probs[blockIdx.x];
  const DevState& St = states[blockIdx.x];
  const int tid = threadIdx.x, lane = tid & 31;
  const int nthreads = blockDim.x;
  const int M = P.M, nr = P.nr;
  const bool kal = tid < 32;                 // warp 0 owns the Kalman blocks
  const bool active = kal && lane < M;
  const int n = (lane < M) ? lane : M - 1;

  extern __shared__ double sm[];
  const AdfSmem L(nthreads >> 5, NV, P.S, M, nr, BM, tab_smem != 0);
  double* s_mu = sm + L.mu;
  double* s_s2 = sm + L.s2;
  double* s_part = sm + L.part;
  double* s_fin = sm + L.fin;
  double* s_wn = sm + L.wn;
  double* s_xn = sm + L.xn;
  for (int i = tid; i < P.S; i += nthreads) s_wn[i] = P.wn[i];
  for (int i = tid; i < kNP * P.S; i += nthreads) s_xn[i] = P.xn[i];
  const double* thr = P.thr;
  const double* hphtab = P.HPHtab;
  const double* wtab = P.Wtab;
  if (tab_smem) {
    double* d0 = sm + L.thr; double* d1_ = sm + L.hph; double* d2_ = sm + L.wtab;
    for (int i = tid; i < nr - 1; i += nthreads) d0[i] = P.thr[i];
    for (int i = tid; i < M * (nr + 1); i += nthreads) d1_[i] = P.HPHtab[i];
    for (int i = tid; i < M * (nr + 1) * BM; i += nthreads) d2_[i] = P.Wtab[i];
    thr = d0; hphtab = d1_; wtab = d2_;
  }
  __syncthreads();
  MomParams mp = make_mom_params(P, P.W, s_wn, s_xn);
  MomCtaThread<DPT> th;
  th.init(mp, tid);
  const double pep = pep_const(mp.kind, mp.sn2, 1.0);     // alpha = 1 in the filter (:256)

  double A[BM * BM], hA[BM], hv[BM], m[BM];
#pragma unroll
  for (int i = 0; i < BM * BM; ++i) A[i] = P.A[n * BM * BM + i];
#pragma unroll
  for (int i = 0; i < BM; ++i) { hA[i] = P.hA[n * BM + i]; hv[i] = P.h[n * BM + i]; }
  const int off = P.off[n];
  const int b = P.off[n + 1] - off;

  int idx;
  if (k0 == 0) {
    idx = nr;                                 // PP = Pinf at the first step (:246)
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
  // one-step-ahead loads of everything that does not depend on the recurrence
  double y_nx = St.y[k0];
  double tt_nx = kal ? St.ttau[k0 * M + n] : 0.0;
  double tn_nx = kal ? St.tnu[k0 * M + n] : 0.0;
  double R_nx = (kal && !mom_all) ? St.R[k0 * M + n] : 0.0;

  for (long long k = k0; k < k1; ++k) {
    const double y = y_nx;
    const double tt_ld = tt_nx, tn_ld = tn_nx, R_ld = R_nx;
    if (k + 1 < k1) {
      y_nx = St.y[k + 1];
      if (kal) {
        tt_nx = St.ttau[(k + 1) * M + n];
        tn_nx = St.tnu[(k + 1) * M + n];
        if (!mom_all) R_nx = St.R[(k + 1) * M + n];
      }
    }
    const bool do_mom = mom_all || k == T - 1;
    double Am[BM], Wv[BM];
    double fmu = 0.0, HPH = 0.0;
    if (kal) {
#pragma unroll
      for (int i = 0; i < BM; ++i) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < BM; ++j) acc = fma(A[i + j * BM], m[j], acc);
        Am[i] = acc;
        fmu = fma(hA[i], m[i], fmu);          // fmu = (H*A)*m (:250)
      }
      const double* wrow = wtab + ((size_t)n * (nr + 1) + idx) * BM;
      HPH = hphtab[(size_t)n * (nr + 1) + idx];
#pragma unroll
      for (int i = 0; i < BM; ++i) Wv[i] = wrow[i];
      if (do_mom && active) { s_mu[lane] = fmu; s_s2[lane] = HPH; }
    }
    if (do_mom) {
      __syncthreads();
      mom_cta<DPT, SINGLE>(mp, th, 1.0, y, s_mu, s_s2, s_part, s_fin);
    }
    if (kal) {
      double tt, tn, Rk;
      if (do_mom) {
        double Z, d1, d2;
        mom_cta_result<DPT>(mp, pep, s_fin, n, Z, d1, d2);
        const double tt_old = running ? tt_run : tt_ld;
        const double tn_old = running ? tn_run : tn_ld;
        const double den = 1.0 + d2 * HPH;
        tt = (1.0 - ep_damp) * tt_old + ep_damp * (-d2 / den);                  // :265
        tn = (1.0 - ep_damp) * tn_old + ep_damp * ((d1 - fmu * d2) / den);      // :266
        Rk = 1.0 / tt;                                                          // :269 (before the clamp)
        if (lane == 0) St.lZ[k] = log(Z);
      } else {
        tt = tt_ld; tn = tn_ld; Rk = R_ld;
      }
      tt = fmax(tt, 0.0);                     // NaN -> 0, as MATLAB max (:274)
      if (tt == 0.0) {
        Rk = INFINITY;                        // :287
#pragma unroll
        for (int i = 0; i < BM; ++i) m[i] = Am[i];
      } else {
        const double ys = tn / tt;            // :277
        const double g = 1.0 / (HPH + Rk);
        const double innov = ys - fmu;
#pragma unroll
        for (int i = 0; i < BM; ++i) m[i] = fma(Wv[i] * g, innov, Am[i]);       // (A-K h A) m + K ys
      }
      idx = lookup_filter_hint(P.r, thr, nr, Rk, idx);
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
      tt_run = tt; tn_run = tn;
    }
  }
}

// ------------------------------------------------------------ full-state EP
// One CTA per signal, steps 0..T-1 (gf_ep_modulator_nmf.m:126-184; nlZ mode :400-447).
// mom_all: moment matching at every observed step (first EP iteration) or only at
// k == T-1.  nlz: nlZ-mode update rules (clamp at every step, all-sites z-form when
// any site is at the bound, :424-439).
template <int DPT, int BM, bool SINGLE>
__global__ void __launch_bounds__(kAdfMaxThreads)
gfep_filter_cta_kernel(const DevProblem* __restrict__ probs, const DevState* __restrict__ states, long long T,
                       int mom_all, double ep_damp, int nlz, int store) {
  constexpr int NV = MomCta<DPT>::NV;
  const DevProblem& P_ = probs[blockIdx.x];
  const DevState& St = states[blockIdx.x];
  const int tid = threadIdx.x, lane = tid & 31;
  const int nthreads = blockDim.x;
  const int M = P_.M;
  const bool kal = tid < 32;
  const bool active = kal && lane < M;
  const int n = (lane < M) ? lane : M - 1;

  extern __shared__ double sm[];
  const AdfSmem L(nthreads >> 5, NV, P_.S, M, 0, BM, false);
  double* s_mu = sm + L.mu;
  double* s_s2 = sm + L.s2;
  double* s_part = sm + L.part;
  double* s_fin = sm + L.fin;
  double* s_wn = sm + L.wn;
  double* s_xn = sm + L.xn;
  for (int i = tid; i < P_.S; i += nthreads) s_wn[i] = P_.wn[i];
  for (int i = tid; i < kNP * P_.S; i += nthreads) s_xn[i] = P_.xn[i];
  __syncthreads();
  MomParams mp = make_mom_params(P_, P_.W, s_wn, s_xn);
  MomCtaThread<DPT> th;
  th.init(mp, tid);
  const double pep = pep_const(mp.kind, mp.sn2, 1.0);     // alpha = 1 (:144)

  double A[BM * BM], Q[BM * BM], hv[BM], m[BM], P[BM * BM];
#pragma unroll
  for (int i = 0; i < BM * BM; ++i) {
    A[i] = P_.A[n * BM * BM + i];
    Q[i] = P_.Q[n * BM * BM + i];
    P[i] = P_.Pinf[n * BM * BM + i];                  // :117
  }
#pragma unroll
  for (int i = 0; i < BM; ++i) { hv[i] = P_.h[n * BM + i]; m[i] = 0.0; }   // :116
  const int off = P_.off[n];
  const int b = P_.off[n + 1] - off;

  double y_nx = St.y[0];
  double tt_nx = kal ? St.ttau[n] : 0.0;
  double tn_nx = kal ? St.tnu[n] : 0.0;

  for (long long k = 0; k < T; ++k) {
    const double y = y_nx;
    const double tt_ld = tt_nx, tn_ld = tn_nx;
    if (k + 1 < T) {
      y_nx = St.y[k + 1];
      if (kal) { tt_nx = St.ttau[(k + 1) * M + n]; tn_nx = St.tnu[(k + 1) * M + n]; }
    }
    const bool obs = !isnan(y);                          // :135 (uniform over the CTA)
    const bool do_mom = obs && (mom_all || k == T - 1);  // :141
    double fmu = 0.0, HPH = 0.0, Wv[BM], hP[BM];
    if (kal) {
      if (k > 0) {                                       // :129-132
        double t[BM], AP[BM * BM];
#pragma unroll
        for (int i = 0; i < BM; ++i) {
          double s = 0.0;
#pragma unroll
          for (int j = 0; j < BM; ++j) s = fma(A[i + j * BM], m[j], s);
          t[i] = s;
        }
#pragma unroll
        for (int i = 0; i < BM; ++i) m[i] = t[i];
        mat_mul<BM>(A, P, AP);
        mat_mul_bt_add<BM>(AP, A, Q, P);
      }
      if (obs) {
#pragma unroll
        for (int i = 0; i < BM; ++i) {
          fmu = fma(hv[i], m[i], fmu);
          double w = 0.0, g = 0.0;
#pragma unroll
          for (int j = 0; j < BM; ++j) {
            w = fma(P[i + j * BM], hv[j], w);            // W = P*H'
            g = fma(hv[j], P[j + i * BM], g);            // H*P
          }
          Wv[i] = w; hP[i] = g;
        }
#pragma unroll
        for (int i = 0; i < BM; ++i) HPH = fma(hP[i], hv[i], HPH);   // diag(H*P*H')
        if (nlz && active && !(HPH > 0.0)) atomicCAS(St.status, 0, 2);   // `keyboard` trap (:408-410)
        if (do_mom && active) { s_mu[lane] = fmu; s_s2[lane] = HPH; }
      }
    }
    if (do_mom) {
      __syncthreads();
      mom_cta<DPT, SINGLE>(mp, th, 1.0, y, s_mu, s_s2, s_part, s_fin);
    }
    if (kal) {
      if (obs) {
        double tt = tt_ld, tn = tn_ld;
        if (do_mom) {
          double Z, d1, d2;
          mom_cta_result<DPT>(mp, pep, s_fin, n, Z, d1, d2);
          const double den = 1.0 + d2 * HPH;
          tt = (1.0 - ep_damp) * tt + ep_damp * (-d2 / den);                      // :147
          tn = (1.0 - ep_damp) * tn + ep_damp * ((d1 - fmu * d2) / den);          // :148
          if (!nlz) tt = fmax(tt, 0.0);                                           // :151
          if (lane == 0) St.lZ[k] = log(Z);
          if (active && !nlz) St.R[k * M + n] = 1.0 / tt;                         // :154
        }
        if (nlz) tt = fmax(tt, 0.0);                                              // :425
        if (active) { St.ttau[k * M + n] = tt; St.tnu[k * M + n] = tn; }

        const bool at_bound = (tt == 0.0);
        const bool zform = nlz ? (__any_sync(0xffffffffu, active && at_bound) != 0) : at_bound;
        if (zform) {                                      // :162-169 / :428-433
          const double z = tt * HPH + 1.0;
          const double gk = tt / z;
          const double v = (tt * fmu - tn) / z;
#pragma unroll
          for (int i = 0; i < BM; ++i) m[i] = fma(-Wv[i], v, m[i]);
#pragma unroll
          for (int j = 0; j < BM; ++j)
#pragma unroll
            for (int i = 0; i < BM; ++i) P[i + j * BM] = fma(-(Wv[i] * gk), Wv[j], P[i + j * BM]);
        } else {                                          // :171-176 / :435-438
          const double g = 1.0 / (HPH + 1.0 / tt);
          const double v = tn / tt - fmu;
#pragma unroll
          for (int i = 0; i < BM; ++i) m[i] = fma(Wv[i] * g, v, m[i]);
#pragma unroll
          for (int j = 0; j < BM; ++j)
#pragma unroll
            for (int i = 0; i < BM; ++i) P[i + j * BM] = fma(-(Wv[i] * g), hP[j], P[i + j * BM]);   // P - K*H*P
        }
      }
      if (active && store) {                              // :181-182
#pragma unroll
        for (int i = 0; i < BM; ++i) if (i < b) St.MS[k * P_.n + off + i] = m[i];
        double* dst = St.PS + ((size_t)k * M + n) * BM * BM;
#pragma unroll
        for (int i = 0; i < BM * BM; ++i) dst[i] = P[i];
        if (k == T - 1) {
          double e = 0.0, v = 0.0;
#pragma unroll
          for (int i = 0; i < BM; ++i) {
            e = fma(hv[i], m[i], e);
            double g = 0.0;
#pragma unroll
            for (int j = 0; j < BM; ++j) g = fma(hv[j], P[j + i * BM], g);
            v = fma(g, hv[i], v);
          }
          St.E[k * M + n] = e;
          St.V[k * M + n] = v;
        }
      }
    }
  }
}

}  // namespace nsagp
