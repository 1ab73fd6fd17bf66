// Sequential (ADF) filter passes, one CTA per signal.
//
// References: matlab/ihgp_ep_modulator_nmf.m:233-310 (infinite-horizon filter),
// matlab/gf_ep_modulator_nmf.m:126-184 / :400-447 (full-state filter).
//
// The pass is a nonlinear recurrence in time, so one step's latency is the whole
// cost.  Roles inside the CTA (momcta.cuh):
//   warp 0 ("Kalman warp"), lane n < M : predict / update of latent block n, in registers;
//   warps 1.. ("moment warps")         : the sigma-point moment matching, 4 threads per point.
// Per step the Kalman warp publishes the cavity (predicted marginal) in shared
// memory and arrives at a named barrier; the moment warps wait there, integrate,
// leave their partial sums in shared memory and arrive at a second barrier the
// Kalman warp waits on.  While the moment warps work, the Kalman warp issues the
// previous step's global stores and its logarithm / division for the outputs that
// are not part of the recurrence (lZ, R).  On the critical path the site update and
// the Kalman gain are fused into two reciprocals (momcta.cuh: adf_site_from_sums,
// and the "z-form" of the update, gf_ep_modulator_nmf.m:162-169, which equals the
// gain form :171-176 algebraically), the steady-state tables sit in shared memory
// and are indexed by a threshold search on ttau itself (no division), trying the
// previous step's row first.
#pragma once
#include "common.cuh"
#include "lookup.cuh"
#include "mom.cuh"
#include "momcta.cuh"
#include "ihgp.cuh"
#include "gfep.cuh"

namespace nsagp {

constexpr int kAdfMaxMomThreads = 352;     // 11 moment warps + the Kalman warp = 384 threads

// idx = #{i : thr[i] <= 1/tt} = #{i : tt <= cthr[i]} (cthr descending, nr-1 entries; host-built so
// that the decision is bit-identical to the reference's min(abs(r-R)) with R = 1./ttau).
// In the ADF sweep the sites of consecutive steps are unrelated (old sites are zero), so R jumps by decades from
// step to step and a window around the previous row misses most of the time.  The row is guessed from the bits of
// ttau instead: the high word of a double is a piecewise-linear log2 (error <= 0.086), the grid is log-spaced, so
// guess = ga + gb * hi32(ttau) is within one row of the answer; six thresholds around the guess are loaded at once
// and counted (the array carries kCthrPad sentinels on either side: no bounds checks).  A grid that is not
// log-spaced only costs the binary search.
struct TtauGuess { double a, b; };

__device__ __forceinline__ TtauGuess make_ttau_guess(const double* r, int nr) {
  // row position of R on the grid: (log10 R - log10 r0) * scale; R = 1/ttau; log2(ttau) ~ hi32/2^20 - 1023 + 0.043;
  // #{thresholds <= R} ~ floor(position + 0.49) for arithmetic mid-point thresholds on a log-spaced grid
  const double scale = (double)(nr - 1) / (log10(r[nr - 1]) - log10(r[0]));
  const double l2 = 0.30102999566398120 * scale;
  TtauGuess g;
  g.b = -l2 / 1048576.0;
  g.a = -log10(r[0]) * scale + (1023.0 - 0.043) * l2 + 0.49;
  return g;
}

__device__ __forceinline__ int lookup_by_ttau(const double* cthr, int nr, double tt, const TtauGuess& g) {
  // clamped to the table: R below / above the grid then finds its answer (row 0 / nr-1) in the middle of the window
  const int guess = min(max(__double2int_rd(fma((double)__double2hiint(tt), g.b, g.a)), 0), nr - 1);
  const int w0 = guess - 3;
  const double* w = cthr + w0;
  int cnt = 0;
#pragma unroll
  for (int j = 0; j < 6; ++j) cnt += (tt <= w[j]) ? 1 : 0;
  if (cnt > 0 && cnt < 6) return w0 + cnt;   // the change from "<=" to ">" lies inside the window
  int a = 0, b = nr - 1;
  while (a < b) {
    const int mid = (a + b) >> 1;
    if (tt <= cthr[mid]) a = mid + 1; else b = mid;
  }
  return a;
}

// Shared-memory layout (doubles) of the CTA kernels.
struct AdfSmem {
  int logtab, mom, wn, xn, q, cthr, hph, wtab, sdt, rs2t, total;
  __host__ __device__ AdfSmem(int nmw, int NVP, int S, int M, int N, int nr, int BM, bool tables, bool fullstate) {
    int o = 0;
    logtab = o; o += kLogTabDoubles;                   // fastmath.cuh: log_ge1_tab (first: 16-byte aligned)
    mom = o; o += 72 + nmw * 4 * NVP;
    wn = o; o += S;
    xn = o; o += kNP * S;
    q = o; o += fullstate ? M * BM * BM : 0;
    cthr = o; o += tables ? nr - 1 + 2 * kCthrPad : 0;
    hph = o; o += tables ? M * (nr + 1) : 0;
    wtab = o; o += tables ? M * (nr + 1) * BM : 0;
    sdt = o; o += tables ? N * (nr + 1) : 0;
    rs2t = o; o += tables ? N * (nr + 1) : 0;
    total = o;
  }
};

// Parallel-in-time ADF with burn-in overlap (opt-in, approximate, error-reported).  The window
// [w0, w1) is cut into chunks of chunk_len steps, one CTA (blockIdx.y) per chunk.  A chunk that does not start at
// step 0 starts `burn` steps early from the stationary prior (m = 0, P = Pinf / the Pinf table row), runs the
// reference's literal recursion through the burn-in WITHOUT storing anything, and stores from its own first step
// on.  The filter forgets its initial state at the rate of the slowest latent, so the deviation from
// the exact sequential pass decays with `burn`; it is measured, not assumed: the state a chunk holds at the end of
// its burn-in is recorded in bstate [chunk][n] and compared with what the preceding chunk stored for that step.
// Cold start only (old sites are zero, as in every call of the reference): the burn-in never reads the site arrays,
// which the preceding chunk is writing at that moment.  chunk_len = 0: the exact single pass.
struct AdfPar {
  long long w0, w1, chunk_len, burn;
  double* bstate;
};

struct AdfWindow {
  long long ks, kw0, kw1;       // first step executed, first step stored, end
  bool chunked;
  __device__ __forceinline__ AdfWindow(const AdfPar& par, long long k0, long long k1) {
    chunked = par.chunk_len > 0;
    if (!chunked) { ks = k0; kw0 = k0; kw1 = k1; return; }
    kw0 = par.w0 + (long long)blockIdx.y * par.chunk_len;
    kw1 = kw0 + par.chunk_len < par.w1 ? kw0 + par.chunk_len : par.w1;
    ks = kw0 - par.burn > 0 ? kw0 - par.burn : 0;
  }
};

// max over the chunk boundaries inside one launch of |state at the end of chunk c's burn-in - what chunk c-1 stored for
// that step| and of |stored mean|: diag[0], diag[1] (bit patterns of non-negative doubles, atomicMax).
__global__ void adf_mismatch_kernel(const DevProblem* __restrict__ probs, const DevState* __restrict__ states, AdfPar par,
                                    int nch, unsigned long long* __restrict__ diag) {
  const DevProblem& P = probs[blockIdx.x];
  const DevState& St = states[blockIdx.x];
  double dm = 0.0, mm = 0.0;
  for (int i = threadIdx.x; i < (nch - 1) * P.n; i += blockDim.x) {
    const int c = 1 + i / P.n, j = i - (c - 1) * P.n;
    const long long kw0 = par.w0 + (long long)c * par.chunk_len;
    if (kw0 >= par.w1 || kw0 - par.burn <= 0) continue;       // (a chunk that reaches back to step 0 is exact)
    const double ref = St.MS[(kw0 - 1) * P.n + j];
    dm = fmax(dm, fabs(par.bstate[((size_t)blockIdx.x * nch + c) * P.n + j] - ref));
    mm = fmax(mm, fabs(ref));
  }
  atomic_max_nonneg(diag, dm);
  atomic_max_nonneg(diag + 1, mm);
}

// Loop of a moment thread: steps k0..k1-1, moments wherever the Kalman warp asks.
template <int DPT, bool SINGLE>
__device__ __forceinline__ void adf_moment_loop(const MomParams& mp, const double* __restrict__ yv, long long T,
                                                long long k0, long long k1, int mom_all, bool skip_nan,
                                                double* s_mom, const double* s_logtab, int mtid, int nmt, int nthreads) {
  MomCtaThread<DPT> th;
  th.init(mp, mtid);
  th.ltab = (unsigned)__cvta_generic_to_shared(s_logtab);
  const double noise = mp.sn2;                        // alpha = 1 in the filter (:256 / :144)
  double* s_part = s_mom + MomCta<DPT>::kCav;
  double y_nx = yv[k0];
  for (long long k = k0; k < k1; ++k) {
    const double y = y_nx;
    const bool do_mom = (mom_all || k == T - 1) && !(skip_nan && isnan(y));
    if (do_mom) {
      named_bar_sync(kBarCavity, nthreads);
      mom_cta_points<DPT, SINGLE>(mp, th, noise, y, s_mom, s_part, mtid, nmt);
      named_bar_arrive(kBarSums, nthreads);
    }
    if (k + 1 < k1) y_nx = yv[k + 1];      // lands while the Kalman warp works
  }
}

// ----------------------------------------------------------------- IHGP
// Steps k0..k1-1.  mom_all: moment matching at every step (first pass) or only at
// k == T-1.  running: the _constraints nlZ variant's running site vectors.
// TABS: steady-state tables in shared memory (the host checks that they fit).
template <int DPT, int BM, bool SINGLE, bool TABS>
__global__ void __launch_bounds__(32 + kAdfMaxMomThreads)
ihgp_adf_cta_kernel(const DevProblem* __restrict__ probs, const DevState* __restrict__ states,
                    long long T, long long k0_, long long k1_, int mom_all, double ep_damp, int running, AdfPar par) {
  constexpr int NVP = MomCta<DPT>::NVP;
  const AdfWindow win(par, k0_, k1_);
  const long long k0 = win.ks, k1 = win.kw1, kw0 = win.kw0;
  if (kw0 >= k1) return;
  const DevProblem& P = probs[blockIdx.x];
  const DevState& St = states[blockIdx.x];
  const int tid = threadIdx.x, lane = tid & 31;
  const int nthreads = blockDim.x, nmt = nthreads - 32, nmw = nmt >> 5;
  const int M = P.M, nr = P.nr, D = P.D;

  extern __shared__ double sm[];
  const AdfSmem L(nmw, NVP, P.S, M, P.N, nr, BM, TABS, false);
  double* s_mom = sm + L.mom;
  double* s_wn = sm + L.wn;
  double* s_xn = sm + L.xn;
  for (int i = tid; i < P.S; i += nthreads) s_wn[i] = P.wn[i];
  for (int i = tid; i < kNP * P.S; i += nthreads) s_xn[i] = P.xn[i];
  log_tab_fill(sm + L.logtab, tid, nthreads);
  // steady-state tables: shared memory when they fit (TABS), else straight from HBM/L2
  const double* cthr = TABS ? sm + L.cthr + kCthrPad : P.cthr;
  const double* hphtab = TABS ? sm + L.hph : P.HPHtab;
  const double* wtab = TABS ? sm + L.wtab : P.Wtab;
  const double* sdtab = TABS ? sm + L.sdt : P.SDtab;
  const double* rs2tab = TABS ? sm + L.rs2t : P.RS2tab;
  if (TABS) {
    for (int i = tid; i < nr - 1 + 2 * kCthrPad; i += nthreads) sm[L.cthr + i] = P.cthr[i - kCthrPad];
    for (int i = tid; i < M * (nr + 1); i += nthreads) sm[L.hph + i] = P.HPHtab[i];
    for (int i = tid; i < M * (nr + 1) * BM; i += nthreads) sm[L.wtab + i] = P.Wtab[i];
    for (int i = tid; i < P.N * (nr + 1); i += nthreads) { sm[L.sdt + i] = P.SDtab[i]; sm[L.rs2t + i] = P.RS2tab[i]; }
  }
  __syncthreads();
  const MomParams mp = make_mom_params(P, P.W, s_wn, s_xn);

  if (tid >= 32) {
    adf_moment_loop<DPT, SINGLE>(mp, St.y, T, k0, k1, mom_all, true, s_mom, sm + L.logtab, tid - 32, nmt, nthreads);
    return;
  }

  // ------------------------------------------------------------ Kalman warp
  const bool active = lane < M;
  const int n = active ? lane : M - 1;
  const bool is_mod = active && n >= D;
  const double pep = pep_const(mp.kind, mp.sn2, 1.0);     // alpha = 1 in the filter (:256)
  MomCtaSumAddr<DPT> sum_addr;
  sum_addr.init(s_mom + MomCta<DPT>::kCav, n, D);
  double A[BM * BM], hA[BM], hv[BM], m[BM];
#pragma unroll
  for (int i = 0; i < BM * BM; ++i) A[i] = P.A[n * BM * BM + i];
#pragma unroll
  for (int i = 0; i < BM; ++i) { hA[i] = P.hA[n * BM + i]; hv[i] = P.h[n * BM + i]; }
  const int off = P.off[n];
  const int b = P.off[n + 1] - off;

  const TtauGuess guess = make_ttau_guess(P.r, nr);
  int idx;
  if (k0 == 0) {
    idx = nr;                                 // PP = Pinf at the first step (:246)
#pragma unroll
    for (int i = 0; i < BM; ++i) m[i] = St.mcarry[n * BM + i];
  } else if (win.chunked) {
    idx = nr;                                 // burn-in from the stationary prior
#pragma unroll
    for (int i = 0; i < BM; ++i) m[i] = 0.0;
  } else {
#pragma unroll
    for (int i = 0; i < BM; ++i) m[i] = (i < b) ? St.MS[(k0 - 1) * P.n + off + i] : 0.0;
    const double ttp = fmax(St.ttau[(k0 - 1) * M + n], 0.0);
    const double Rp = (ttp == 0.0) ? INFINITY : St.R[(k0 - 1) * M + n];
    idx = lookup_filter(P.r, P.thr, nr, Rp);
  }
  // one-step-ahead loads of what does not depend on the recurrence
  const bool ld_sites = !win.chunked;        // (chunked: cold start, old sites are zero and must not be read)
  double y_nx = St.y[k0];
  double tt_nx = ld_sites ? St.ttau[k0 * M + n] : 0.0;
  double tn_nx = ld_sites ? St.tnu[k0 * M + n] : 0.0;
  double R_nx = mom_all ? 0.0 : St.R[k0 * M + n];
  // outputs of the previous step, stored while the moment warps work
  bool pend = false, pend_mom = false;
  long long pk = 0;
  double p_tt = 0.0, p_tn = 0.0, p_ttraw = 0.0, p_R = 0.0, p_Z = 1.0;

  for (long long k = k0; k < k1; ++k) {
    const double tt_ld = tt_nx, tn_ld = tn_nx, R_ld = R_nx;
    const bool do_mom = mom_all || k == T - 1;
    // A missing sample (y = NaN) is not skipped by the reference (:253-271 has no isnan test):
    // its moments are NaN, so ttau = max(NaN, 0) = 0, tnu = NaN and lZ = log(pEP * jitter).
    // The moment warps never see it; the same values are produced here.
    const bool y_nan = isnan(y_nx);
    const bool ask = do_mom && !y_nan;
    double Am[BM], Wv[BM];
    double fmu = 0.0;
#pragma unroll
    for (int i = 0; i < BM; ++i) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < BM; ++j) acc = fma(A[i + j * BM], m[j], acc);
      Am[i] = acc;
      fmu = fma(hA[i], m[i], fmu);            // fmu = (H*A)*m (:250)
    }
    const double HPH = hphtab[(size_t)n * (nr + 1) + idx];
    if (ask) {
      if (active) {
        s_mom[lane] = fmu; s_mom[32 + lane] = HPH;
        if (is_mod) {
          s_mom[64 + n - D] = sdtab[(size_t)(n - D) * (nr + 1) + idx];
          s_mom[68 + n - D] = rs2tab[(size_t)(n - D) * (nr + 1) + idx];
        }
      }
      __syncwarp();
      named_bar_arrive(kBarCavity, nthreads);
    }
    {
      const double* wrow = wtab + ((size_t)n * (nr + 1) + idx) * BM;
#pragma unroll
      for (int i = 0; i < BM; ++i) Wv[i] = wrow[i];
    }
    // ---- off the critical path: next step's loads, previous step's outputs ----
    if (k + 1 < k1) {
      y_nx = St.y[k + 1];
      if (ld_sites) {
        tt_nx = St.ttau[(k + 1) * M + n];
        tn_nx = St.tnu[(k + 1) * M + n];
      }
      if (!mom_all) R_nx = St.R[(k + 1) * M + n];
    }
    if (pend && pk < kw0) {
      if (pk == kw0 - 1 && active && par.bstate) {                              // end of the burn-in: kept for the mismatch check
#pragma unroll
        for (int i = 0; i < BM; ++i) if (i < b) par.bstate[((size_t)blockIdx.x * gridDim.y + blockIdx.y) * P.n + off + i] = m[i];
      }
    } else if (pend) {
      double Rk = p_R;
      if (pend_mom) {
        Rk = 1.0 / p_ttraw;                                                     // :269 (before the clamp)
        if (lane == 0) St.lZ[pk] = log(pep * p_Z);
      }
      if (p_tt == 0.0) Rk = INFINITY;                                           // :287
      if (active) {
        St.ttau[pk * M + n] = p_tt; St.tnu[pk * M + n] = p_tn; St.R[pk * M + n] = Rk;
#pragma unroll
        for (int i = 0; i < BM; ++i) if (i < b) St.MS[pk * P.n + off + i] = m[i];
      }
    }
    // ---- this step ----
    double tt, tn, ttraw, Zm = 1.0;
    if (do_mom) {
      double tt_new = NAN, tn_new = NAN;
      Zm = kJitter;
      if (ask) {
        named_bar_sync(kBarSums, nthreads);
        double Zs, r1, r2;
        mom_cta_sums<DPT>(sum_addr, nmw, Zs, r1, r2);
        adf_site_from_sums(Zs, r1, r2, fmu, HPH, tt_new, tn_new, Zm);
      }
      const double tt_old = running ? p_tt : tt_ld;
      const double tn_old = running ? p_tn : tn_ld;
      ttraw = fma(ep_damp, tt_new, (1.0 - ep_damp) * tt_old);                   // :265
      tn = fma(ep_damp, tn_new, (1.0 - ep_damp) * tn_old);                      // :266
    } else {
      ttraw = tt_ld; tn = tn_ld;
    }
    tt = fmax(ttraw, 0.0);                    // NaN -> 0, as MATLAB max (:274)
    if (tt == 0.0) {
#pragma unroll
      for (int i = 0; i < BM; ++i) m[i] = Am[i];                                // :287-289
      idx = 0;                                // R = Inf: every distance is Inf, min() returns the first index (:239)
    } else {
      // (A - K h A) m + K ys with K = W/(HPH + 1/ttau), ys = tnu/ttau  (:291-300), as one reciprocal
      const double c = (tn - tt * fmu) * rcp_fast(fma(tt, HPH, 1.0));
#pragma unroll
      for (int i = 0; i < BM; ++i) m[i] = fma(Wv[i], c, Am[i]);
      if (do_mom) {
        idx = (tt > 2e-12) ? lookup_by_ttau(cthr, nr, tt, guess) : lookup_filter(P.r, P.thr, nr, 1.0 / tt);
      } else {
        idx = lookup_filter(P.r, P.thr, nr, R_ld);
      }
    }
    pend = true; pend_mom = do_mom; pk = k;
    p_tt = tt; p_tn = tn; p_ttraw = ttraw; p_R = R_ld; p_Z = Zm;
  }
  if (pend) {
    double Rk = p_R;
    if (pend_mom) {
      Rk = 1.0 / p_ttraw;
      if (lane == 0) St.lZ[pk] = log(pep * p_Z);
    }
    if (p_tt == 0.0) Rk = INFINITY;
    if (active) {
      St.ttau[pk * M + n] = p_tt; St.tnu[pk * M + n] = p_tn; St.R[pk * M + n] = Rk;
#pragma unroll
      for (int i = 0; i < BM; ++i) if (i < b) St.MS[pk * P.n + off + i] = m[i];
      if (pk == T - 1) {
        double e = 0.0;
#pragma unroll
        for (int i = 0; i < BM; ++i) e = fma(hv[i], m[i], e);
        St.E[pk * M + n] = e;
      }
    }
  }
}

// ------------------------------------------------------------ full-state EP
// One CTA per signal, steps k0..T-1 (gf_ep_modulator_nmf.m:126-184; nlZ mode :400-447).
// mom_all: moment matching at every observed step (first EP iteration) or only at
// k == T-1.  nlz: nlZ-mode rules (clamp at every step, :425).  The measurement
// update is written in the z-form for every site (:162-169 / :428-433), which is
// algebraically the gain form (:171-176 / :435-438) and covers ttau = 0 without a branch.
template <int DPT, int BM, bool SINGLE>
__global__ void __launch_bounds__(32 + kAdfMaxMomThreads)
gfep_filter_cta_kernel(const DevProblem* __restrict__ probs, const DevState* __restrict__ states, long long T,
                       long long k0_, int mom_all, double ep_damp, int nlz, int store, AdfPar par) {
  constexpr int NVP = MomCta<DPT>::NVP;
  const AdfWindow win(par, k0_, T);
  const long long k0 = win.ks, k1 = win.kw1, kw0 = win.kw0;
  if (kw0 >= k1) return;
  const DevProblem& P_ = probs[blockIdx.x];
  const DevState& St = states[blockIdx.x];
  const int tid = threadIdx.x, lane = tid & 31;
  const int nthreads = blockDim.x, nmt = nthreads - 32, nmw = nmt >> 5;
  const int M = P_.M, D = P_.D;

  extern __shared__ double sm[];
  const AdfSmem L(nmw, NVP, P_.S, M, P_.N, 0, BM, false, true);
  double* s_mom = sm + L.mom;
  double* s_wn = sm + L.wn;
  double* s_xn = sm + L.xn;
  double* s_q = sm + L.q;
  for (int i = tid; i < P_.S; i += nthreads) s_wn[i] = P_.wn[i];
  for (int i = tid; i < kNP * P_.S; i += nthreads) s_xn[i] = P_.xn[i];
  for (int i = tid; i < M * BM * BM; i += nthreads) s_q[i] = P_.Q[i];
  log_tab_fill(sm + L.logtab, tid, nthreads);
  __syncthreads();
  const MomParams mp = make_mom_params(P_, P_.W, s_wn, s_xn);

  if (tid >= 32) {
    adf_moment_loop<DPT, SINGLE>(mp, St.y, T, k0, k1, mom_all, true, s_mom, sm + L.logtab, tid - 32, nmt, nthreads);
    return;
  }

  // ------------------------------------------------------------ Kalman warp
  const bool active = lane < M;
  const int n = active ? lane : M - 1;
  const bool is_mod = active && n >= D;
  const double pep = pep_const(mp.kind, mp.sn2, 1.0);     // alpha = 1 (:144)
  MomCtaSumAddr<DPT> sum_addr;
  sum_addr.init(s_mom + MomCta<DPT>::kCav, n, D);
  const double* Qn = s_q + n * BM * BM;
  double A[BM * BM], hv[BM], hA[BM], m[BM], P[BM * BM];
#pragma unroll
  for (int i = 0; i < BM * BM; ++i) { A[i] = P_.A[n * BM * BM + i]; P[i] = P_.Pinf[n * BM * BM + i]; }   // :117
#pragma unroll
  for (int i = 0; i < BM; ++i) { hv[i] = P_.h[n * BM + i]; hA[i] = P_.hA[n * BM + i]; m[i] = 0.0; }       // :116
  double hQh = 0.0;
#pragma unroll
  for (int i = 0; i < BM; ++i) {
    double w = 0.0;
#pragma unroll
    for (int j2 = 0; j2 < BM; ++j2) w = fma(Qn[i + j2 * BM], hv[j2], w);
    hQh = fma(hv[i], w, hQh);
  }
  const int off = P_.off[n];
  const int b = P_.off[n + 1] - off;

  if (k0 > 0 && !win.chunked) {                        // continue from the stored estimate of step k0-1
                                                       // (chunked: burn-in from the stationary prior (0, Pinf))
#pragma unroll
    for (int i = 0; i < BM; ++i) m[i] = (i < b) ? St.MS[(k0 - 1) * P_.n + off + i] : 0.0;
    const double* src = St.PS + ((size_t)(k0 - 1) * M + n) * BM * BM;
#pragma unroll
    for (int i = 0; i < BM * BM; ++i) P[i] = src[i];
  }
  const bool ld_sites = !win.chunked;                  // (chunked: cold start, old sites are zero and must not be read)
  double y_nx = St.y[k0];
  double tt_nx = ld_sites ? St.ttau[k0 * M + n] : 0.0;
  double tn_nx = ld_sites ? St.tnu[k0 * M + n] : 0.0;
  bool pend = false, pend_obs = false, pend_mom = false;
  long long pk = 0;
  double p_tt = 0.0, p_tn = 0.0, p_Z = 1.0;

  auto flush = [&](bool last) {
    if (pk < kw0) {                                                               // burn-in: nothing is stored
      if (pk == kw0 - 1 && active && par.bstate) {
#pragma unroll
        for (int i = 0; i < BM; ++i) if (i < b) par.bstate[((size_t)blockIdx.x * gridDim.y + blockIdx.y) * P_.n + off + i] = m[i];
      }
      return;
    }
    // outputs of step pk: (m, P) still hold its posterior (:181-182)
    if (pend_mom) {
      if (lane == 0) St.lZ[pk] = log(pep * p_Z);
      if (active && !nlz) St.R[pk * M + n] = 1.0 / p_tt;                          // :154
    }
    if (active && pend_obs) { St.ttau[pk * M + n] = p_tt; St.tnu[pk * M + n] = p_tn; }
    if (active && store) {
#pragma unroll
      for (int i = 0; i < BM; ++i) if (i < b) St.MS[pk * P_.n + off + i] = m[i];
      double* dst = St.PS + ((size_t)pk * M + n) * BM * BM;
#pragma unroll
      for (int i = 0; i < BM * BM; ++i) dst[i] = P[i];
      if (last) {
        double e = 0.0, v = 0.0;
#pragma unroll
        for (int i = 0; i < BM; ++i) {
          e = fma(hv[i], m[i], e);
          double g = 0.0;
#pragma unroll
          for (int j = 0; j < BM; ++j) g = fma(hv[j], P[j + i * BM], g);
          v = fma(g, hv[i], v);
        }
        St.E[pk * M + n] = e;
        St.V[pk * M + n] = v;
      }
    }
  };

  for (long long k = k0; k < k1; ++k) {
    const double y = y_nx;
    const double tt_ld = tt_nx, tn_ld = tn_nx;
    const bool obs = !isnan(y);                          // :135 (uniform over the CTA)
    const bool do_mom = obs && (mom_all || k == T - 1);  // :141
    // Cavity of step k straight from the posterior of step k-1, in O(b^2): with hA = h A,
    //   fmu = h A m,   h (A P A' + Q) h' = hA P hA' + h Q h'
    // so the moment warps can start before the O(b^3) prediction of the covariance, which
    // this warp then computes while they integrate.
    double fmu = 0.0, HPH = 0.0;
    if (obs) {
      if (k > 0) {
        double v = hQh;
#pragma unroll
        for (int i = 0; i < BM; ++i) {
          fmu = fma(hA[i], m[i], fmu);
          double w = 0.0;
#pragma unroll
          for (int j2 = 0; j2 < BM; ++j2) w = fma(P[i + j2 * BM], hA[j2], w);
          v = fma(hA[i], w, v);
        }
        HPH = v;
      } else {
#pragma unroll
        for (int i = 0; i < BM; ++i) {
          fmu = fma(hv[i], m[i], fmu);
          double w = 0.0;
#pragma unroll
          for (int j2 = 0; j2 < BM; ++j2) w = fma(P[i + j2 * BM], hv[j2], w);
          HPH = fma(hv[i], w, HPH);
        }
      }
      if (do_mom) {
        if (active) {
          s_mom[lane] = fmu; s_mom[32 + lane] = HPH;
          if (is_mod) { s_mom[64 + n - D] = sqrt_fast(HPH); s_mom[68 + n - D] = rcp_fast(HPH); }
        }
        __syncwarp();
        named_bar_arrive(kBarCavity, nthreads);
      }
      if (nlz && active && !(HPH > 0.0)) atomicCAS(St.status, 0, 2);   // `keyboard` trap (:408-410)
    }
    // ---- off the critical path: next step's loads, previous step's outputs ----
    if (k + 1 < k1) {
      y_nx = St.y[k + 1];
      if (ld_sites) {
        tt_nx = St.ttau[(k + 1) * M + n];
        tn_nx = St.tnu[(k + 1) * M + n];
      }
    }
    if (pend) flush(false);
    // predicted moments of step k (:129-132), kept apart from the posterior of step k-1
    double mq[BM], Pq[BM * BM], Wv[BM], hP[BM];
    if (k > 0) {
      double AP[BM * BM], Qr[BM * BM];
#pragma unroll
      for (int i = 0; i < BM * BM; ++i) Qr[i] = Qn[i];
#pragma unroll
      for (int i = 0; i < BM; ++i) {
        double s = 0.0;
#pragma unroll
        for (int j2 = 0; j2 < BM; ++j2) s = fma(A[i + j2 * BM], m[j2], s);
        mq[i] = s;
      }
      mat_mul<BM>(A, P, AP);
      mat_mul_bt_add<BM>(AP, A, Qr, Pq);
    } else {
#pragma unroll
      for (int i = 0; i < BM; ++i) mq[i] = m[i];
#pragma unroll
      for (int i = 0; i < BM * BM; ++i) Pq[i] = P[i];
    }
    if (obs) {
#pragma unroll
      for (int i = 0; i < BM; ++i) {
        double w = 0.0, g = 0.0;
#pragma unroll
        for (int j2 = 0; j2 < BM; ++j2) {
          w = fma(Pq[i + j2 * BM], hv[j2], w);           // W = P*H'
          g = fma(hv[j2], Pq[j2 + i * BM], g);           // H*P
        }
        Wv[i] = w; hP[i] = g;
      }
    }
    double tt = tt_ld, tn = tn_ld, Zm = 1.0;
    if (obs) {
      if (do_mom) {
        named_bar_sync(kBarSums, nthreads);
        double Zs, r1, r2, tt_new, tn_new;
        mom_cta_sums<DPT>(sum_addr, nmw, Zs, r1, r2);
        adf_site_from_sums(Zs, r1, r2, fmu, HPH, tt_new, tn_new, Zm);
        tt = fma(ep_damp, tt_new, (1.0 - ep_damp) * tt_ld);                       // :147
        tn = fma(ep_damp, tn_new, (1.0 - ep_damp) * tn_ld);                       // :148
        if (!nlz) tt = fmax(tt, 0.0);                                             // :151
      }
      if (nlz) tt = fmax(tt, 0.0);                                                // :425
      const double rz = rcp_fast(fma(tt, HPH, 1.0));
      const double c = (tn - tt * fmu) * rz;
      const double g = tt * rz;
#pragma unroll
      for (int i = 0; i < BM; ++i) m[i] = fma(Wv[i], c, mq[i]);
#pragma unroll
      for (int j = 0; j < BM; ++j)
#pragma unroll
        for (int i = 0; i < BM; ++i) P[i + j * BM] = fma(-(Wv[i] * g), hP[j], Pq[i + j * BM]);
    } else {
#pragma unroll
      for (int i = 0; i < BM; ++i) m[i] = mq[i];
#pragma unroll
      for (int i = 0; i < BM * BM; ++i) P[i] = Pq[i];
    }
    pend = true; pend_obs = obs; pend_mom = do_mom; pk = k;
    p_tt = tt; p_tn = tn; p_Z = Zm;
  }
  if (pend) flush(pk == T - 1);
}

}  // namespace nsagp
