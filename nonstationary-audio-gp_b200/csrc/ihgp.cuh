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
#include "scan.cuh"

namespace nsagp {


__device__ __forceinline__ MomParams make_mom_params(const DevProblem& P, const double* W,
                                                     const double* wn, const double* xn) {
  MomParams mp;
  mp.D = P.D; mp.N = P.N; mp.S = P.S; mp.kind = P.lik_kind;
  mp.sn2 = P.sn2; mp.shift = P.link_shift;
  mp.W = W; mp.wn = wn; mp.xn = xn;
  return mp;
}

// doubles of shared memory the thread-form site update needs beyond its [4][M][TPB] staging and the likelihood
// tables: the per-thread link table and the (byte) index map
__host__ __device__ inline int site_tab_doubles(int ndist, int S, int TPB) {
  return ndist > 0 ? 2 * kNP * ndist * TPB + (kNP * S + 7) / 8 : 0;
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
// ind = #{i : thr[i] <= R} (thr ascending, nthr entries).  The row is guessed from the bits of R (the high word of a
// double is a piecewise-linear log2, the grid is log-spaced: the guess is within one row) and confirmed by counting
// six thresholds around it, loaded at once -- one round trip to L1 instead of the eight dependent ones of a binary
// search, and no dependence on the previous step's row (sites of neighbouring steps can differ by decades early in
// EP).  thr carries kCthrPad sentinels on either side; a grid that is not log-spaced only costs the binary search.
__device__ __forceinline__ int count_le_guess(const DevProblem& P, int nthr, double R) {
  const double* __restrict__ thr = P.thr;
  const int guess = min(max(__double2int_rd(fma((double)__double2hiint(R), P.rg_b, P.rg_a)), 0), nthr);
  const int w0 = guess - 3;
  int cnt = 0;
#pragma unroll
  for (int j = 0; j < 6; ++j) cnt += (thr[w0 + j] <= R) ? 1 : 0;
  if (cnt > 0 && cnt < 6) return w0 + cnt;
  return nearest_by_threshold(thr, nthr + 1, R);
}

__device__ __forceinline__ int lookup_filter_hint(const DevProblem& P, double R) {
  if (!(R > 0.0)) return 0;
  if (isinf(R)) return 0;
  if (R >= kLookupBig) return nearest_bruteforce(P.r, P.nr, R);
  return count_le_guess(P, P.nr - 1, R);
}

__device__ __forceinline__ int lookup_smoother_hint(const DevProblem& P, double R) {
  if (isinf(R)) return P.nr - 1;
  if (!(R > 0.0)) return 0;
  if (R >= kLookupBig) return nearest_bruteforce(P.r, P.nr, R);
  return count_le_guess(P, P.nr - 1, R);
}

template <int BM>
struct Affine {
  double F[BM * BM];     // column-major
  double c[BM];
};

template <int BM>
struct MeanState {
  double m[BM];
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

// What the two IHGP mean recursions share (scan.cuh's Elem interface).
// Hint the lines an element will read at step k towards L1: the element inputs (sites, means) do not depend on the
// scan's running state, but every step of the walk stores (MS, E), so the compiler cannot hoist the next step's
// loads itself and each step would wait for three dependent L2 round trips.
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// BMS: leading dimension of the per-latent blocks in HBM (the plan pads every block to max(bz, bg)); BM: size the
// element computes with -- the block size of the latent family the thread belongs to (scan.cuh runs the subband and
// the modulator family side by side), so a 2 x 2 subband block next to 3 x 3 modulators costs 2 x 2 arithmetic.
template <int BMS, int BM>
struct AffineElemBase {
  using Map = Affine<BM>;
  using State = MeanState<BM>;
  static constexpr int kMapDoubles = BMS * BMS + BMS;     // slot sizes are the same for both families
  static constexpr int kStateDoubles = BMS;
  static constexpr bool kTwoTiles = BMS <= 4;       // scan.cuh ScanBounds: two CTA tiles per SM (register cap 102)
  __device__ static __forceinline__ void compose(Map& acc, const Map& e) { affine_compose<BM>(acc, e); }
  __device__ static __forceinline__ void apply(const Map& e, State& s) { affine_apply<BM>(e, s.m); }
  __device__ static __forceinline__ void store_map(const Map& e, double* dst) {
#pragma unroll
    for (int i = 0; i < BM * BM; ++i) dst[i] = e.F[i];
#pragma unroll
    for (int i = 0; i < BM; ++i) dst[BM * BM + i] = e.c[i];
  }
  __device__ static __forceinline__ void load_map(Map& e, const double* src) {
#pragma unroll
    for (int i = 0; i < BM * BM; ++i) e.F[i] = src[i];
#pragma unroll
    for (int i = 0; i < BM; ++i) e.c[i] = src[BM * BM + i];
  }
  __device__ static __forceinline__ void store_state(const State& s, double* dst) {
#pragma unroll
    for (int i = 0; i < BM; ++i) dst[i] = s.m[i];
  }
  __device__ static __forceinline__ void load_state(State& s, const double* src) {
#pragma unroll
    for (int i = 0; i < BM; ++i) s.m[i] = src[i];
  }
};

// Frozen-site filter step k as an affine map of the block mean (:280-304).
template <int BMS, int BM = BMS>
struct FilterElem : AffineElemBase<BMS, BM> {
  using Map = Affine<BM>;
  using State = MeanState<BM>;
  const DevProblem& P; const DevState& St; int n, off, b;
  double A[BM * BM], hA[BM];
  __device__ __forceinline__ void begin_apply() {}
  __device__ FilterElem(const DevProblem& P_, const DevState& S_, int n_, const ScanArgs&) : P(P_), St(S_), n(n_) {
    off = P.off[n]; b = P.off[n + 1] - off;
#pragma unroll
    for (int i = 0; i < BM * BM; ++i) A[i] = P.A[n * BMS * BMS + (i % BM) + (i / BM) * BMS];
#pragma unroll
    for (int i = 0; i < BM; ++i) hA[i] = P.hA[n * BMS + i];
  }
  // rows [k_lo, k_hi) of the inputs towards L2 (scan.cuh: scan_tile_prefetch); the look-up reads one step back
  __device__ static __forceinline__ void prefetch_rows(const DevProblem& P, const DevState& St, long long k_lo, long long k_hi, bool) {
    const long long k0 = k_lo > 0 ? k_lo - 1 : 0;
    const size_t bytes = (size_t)(k_hi - k0) * P.M * sizeof(double);
    l2_prefetch_bulk(St.ttau + k0 * P.M, bytes);
    l2_prefetch_bulk(St.tnu + k0 * P.M, bytes);
    l2_prefetch_bulk(St.R + k0 * P.M, bytes);
  }
  // Inputs of step k: the sites of steps k and k-1 (the look-up uses R(:,k-1), :239); table row of the look-up.
  // Ring depth 1 (the inputs of step s + 1 are requested before step s is processed): measured against 2, 3 and 4
  // (profiles/r3c_ring.jsonl, r3d_ring.jsonl: 0.75 / 0.77 / 0.91 / 1.02 ms per 10^6 steps) -- every ring slot costs 10
  // registers the walk does not have, and with the tile's rows on their way to L2 one step of lead is enough.
  static constexpr int kPrefetch = 1;
  struct In { double tt, tn, R, ttp, Rp; };
  struct Tab { double w[BM], HPH; };
  __device__ __forceinline__ void load(long long k, In& in) const {
    const int M = P.M;
    in.tt = St.ttau[k * M + n]; in.tn = St.tnu[k * M + n]; in.R = St.R[k * M + n];
    in.ttp = k > 0 ? St.ttau[(k - 1) * M + n] : 0.0;
    in.Rp = k > 0 ? St.R[(k - 1) * M + n] : 0.0;
  }
  __device__ __forceinline__ void lookup(long long k, const In& in, Tab& tb) const {
    const int nr = P.nr;
    int idx = nr;
    if (k > 0) {
      const double ttp = fmax(in.ttp, 0.0);
      const double Rp = (ttp == 0.0) ? INFINITY : in.Rp;
      idx = lookup_filter_hint(P, Rp);
    }
    const double* wrow = P.Wtab + ((size_t)n * (nr + 1) + idx) * BMS;
#pragma unroll
    for (int i = 0; i < BM; ++i) tb.w[i] = wrow[i];
    tb.HPH = P.HPHtab[(size_t)n * (nr + 1) + idx];
  }
  // commit: write back the clamp / Inf marking of the reference's filter loop (:274,:287)
  __device__ __forceinline__ void get_impl(long long k, const In& in, const Tab& tb, Map& e, bool commit) {
    const int M = P.M;
    const double tt = fmax(in.tt, 0.0);                                        // :274
    if (tt == 0.0) {
#pragma unroll
      for (int i = 0; i < BM * BM; ++i) e.F[i] = A[i];
#pragma unroll
      for (int i = 0; i < BM; ++i) e.c[i] = 0.0;
      if (commit) { St.ttau[k * M + n] = 0.0; St.R[k * M + n] = INFINITY; }
      return;
    }
    const double g = 1.0 / (tb.HPH + in.R);
    const double ys = in.tn / tt;
#pragma unroll
    for (int i = 0; i < BM; ++i) {
      const double Ki = tb.w[i] * g;
      e.c[i] = Ki * ys;
#pragma unroll
      for (int j = 0; j < BM; ++j) e.F[i + j * BM] = fma(-Ki, hA[j], A[i + j * BM]);   // A - K (h A)
    }
  }
  __device__ __forceinline__ void get(long long k, const In& in, const Tab& tb, Map& e) { get_impl(k, in, tb, e, false); }
  __device__ __forceinline__ void step(long long k, const In& in, const Tab& tb, State& s) {
    Map e;
    get_impl(k, in, tb, e, true);
    affine_apply<BM>(e, s.m);
#pragma unroll
    for (int i = 0; i < BM; ++i) if (i < b) St.MS[k * P.n + off + i] = s.m[i];
  }
  __device__ __forceinline__ void init(State& s, int, long long) {            // m carried from the previous smoother pass
#pragma unroll
    for (int i = 0; i < BM; ++i) s.m[i] = St.mcarry[n * BMS + i];
  }
  __device__ __forceinline__ void store_final(const State&) {}
  __device__ __forceinline__ void finish_reduce() {}
  __device__ __forceinline__ void finish_apply() {}
};

// RTS mean step k (:379-394): m_k = MS_k + G_k (m_{k+1} - A MS_k).
template <int BMS, int BM = BMS>
struct SmootherElem : AffineElemBase<BMS, BM> {
  using Map = Affine<BM>;
  using State = MeanState<BM>;
  const DevProblem& P; const DevState& St; int n, off, b;
  double A[BM * BM], hv[BM];
  double mdM;
  bool applying = false;
  __device__ __forceinline__ void begin_apply() { applying = true; }
  __device__ SmootherElem(const DevProblem& P_, const DevState& S_, int n_, const ScanArgs&) : P(P_), St(S_), n(n_), mdM(0.0) {
    off = P.off[n]; b = P.off[n + 1] - off;
#pragma unroll
    for (int i = 0; i < BM * BM; ++i) A[i] = P.A[n * BMS * BMS + (i % BM) + (i / BM) * BMS];
#pragma unroll
    for (int i = 0; i < BM; ++i) hv[i] = P.h[n * BMS + i];
  }
  __device__ static __forceinline__ void prefetch_rows(const DevProblem& P, const DevState& St, long long k_lo, long long k_hi, bool apply) {
    l2_prefetch_bulk(St.R + k_lo * P.M, (size_t)(k_hi - k_lo) * P.M * sizeof(double));
    l2_prefetch_bulk(St.MS + k_lo * P.n, (size_t)(k_hi - k_lo) * P.n * sizeof(double));
    if (apply) l2_prefetch_bulk(St.E + k_lo * P.M, (size_t)(k_hi - k_lo) * P.M * sizeof(double));
  }
  // Inputs of step k: R(:,k) (look-up of the smoother gain), the filtered mean, H*MS of the previous iteration.
  static constexpr int kPrefetch = 1;       // (see FilterElem)
  struct In { double R, ms[BM], Eold; };
  struct Tab { double G[BM * BM]; int idx; };
  __device__ __forceinline__ void load(long long k, In& in) const {
    in.R = St.R[k * P.M + n];
#pragma unroll
    for (int i = 0; i < BM; ++i) in.ms[i] = (i < b) ? St.MS[k * P.n + off + i] : 0.0;
    in.Eold = applying ? St.E[k * P.M + n] : 0.0;
  }
  __device__ __forceinline__ void lookup(long long, const In& in, Tab& tb) const {
    tb.idx = lookup_smoother_hint(P, in.R);
    const double* G = P.Gtab + ((size_t)n * P.nr + tb.idx) * BMS * BMS;
#pragma unroll
    for (int i = 0; i < BM * BM; ++i) tb.G[i] = G[(i % BM) + (i / BM) * BMS];
  }
  __device__ __forceinline__ void get_impl(long long k, const In& in, const Tab& tb, Map& e, bool commit) {
    double t[BM];
#pragma unroll
    for (int i = 0; i < BM; ++i) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < BM; ++j) acc = fma(A[i + j * BM], in.ms[j], acc);
      t[i] = acc;
    }
#pragma unroll
    for (int i = 0; i < BM * BM; ++i) e.F[i] = tb.G[i];
#pragma unroll
    for (int i = 0; i < BM; ++i) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < BM; ++j) acc = fma(e.F[i + j * BM], t[j], acc);
      e.c[i] = in.ms[i] - acc;
    }
    if (commit && k == 0) {
      // marginal variance of the last look-up: the reference's Varft (:492) and maxDiffP (:444)
      const double vm = P.vmtab[(size_t)n * P.nr + tb.idx];
      atomic_max_nonneg(St.maxdiff + 1, fabs(St.vm0[n] - vm));
      St.vm0[n] = vm;
    }
  }
  __device__ __forceinline__ void get(long long k, const In& in, const Tab& tb, Map& e) { get_impl(k, in, tb, e, false); }
  __device__ __forceinline__ void step(long long k, const In& in, const Tab& tb, State& s) {
    Map e;
    get_impl(k, in, tb, e, true);
    affine_apply<BM>(e, s.m);
    double ev = 0.0;
#pragma unroll
    for (int i = 0; i < BM; ++i) {
      if (i < b) St.MS[k * P.n + off + i] = s.m[i];
      ev = fma(hv[i], s.m[i], ev);
    }
    // maxDiffM (:440): against H*MS of the previous iteration's smoother
    mdM = fmax(mdM, fabs(in.Eold - ev));
    St.E[k * P.M + n] = ev;
  }
  __device__ __forceinline__ void init(State& s, int, long long kinit) {      // m starts at the filtered mean of the last step
#pragma unroll
    for (int i = 0; i < BM; ++i) s.m[i] = (i < b) ? St.MS[kinit * P.n + off + i] : 0.0;
  }
  __device__ __forceinline__ void store_final(const State&) {}
  __device__ __forceinline__ void finish_reduce() {}
  __device__ __forceinline__ void finish_apply() { atomic_max_nonneg(St.maxdiff, mdM); }
};

// -------------------------------------------------------- site update (EP)
// One thread per time step k in [k0, k1) (the whole pass: [0, T-1)): cavity from the smoothed marginal,
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
                   long long k0, long long k1, double alpha, double ep_damp, int write_lZ, int clamp_R) {
  const DevProblem& P = probs[blockIdx.y];
  const DevState& St = states[blockIdx.y];
  const int tid = threadIdx.x;
  const long long k = k0 + (long long)blockIdx.x * TPB + tid;     // steps k0..k1-1 (k1 <= T-1)
  const int M = P.M, nr = P.nr;

  extern __shared__ double sm[];
  double* s_mu = sm;                       // [M][TPB]
  double* s_s2 = s_mu + M * TPB;           // [M][TPB]
  double* s_d1 = s_s2 + M * TPB;
  double* s_d2 = s_d1 + M * TPB;
  double* s_W = s_d2 + M * TPB;            // [DP*kNP]
  double* s_wn = s_W + DP * kNP;
  double* s_xn = s_wn + P.S;
  double* s_tab = s_xn + kNP * P.S;        // [kNP][ndist][2][TPB] link table of each thread's step
  unsigned char* s_xi = reinterpret_cast<unsigned char*>(s_tab + 2 * kNP * P.ndist * TPB);
  for (int i = tid; i < DP * kNP; i += TPB) s_W[i] = P.W[i];
  for (int i = tid; i < P.S; i += TPB) s_wn[i] = P.wn[i];
  for (int i = tid; i < kNP * P.S; i += TPB) s_xn[i] = P.xn[i];
  if (P.ndist > 0)
    for (int i = tid; i < kNP * P.S; i += TPB) s_xi[i] = P.xidx[i];
  __syncthreads();
  if (k >= k1) return;
  const double y = St.y[k];
  if (isnan(y)) {                          // ihgp :398, gf_ep :237
    if (!FULL && write_lZ) St.lZ[k] = 0.0; // IHGP accumulates a scalar: no term for this step
    return;
  }
  MomParams mp = make_mom_params(P, s_W, s_wn, s_xn);
  mp.nd = P.ndist; mp.xd = P.xdist; mp.xi = s_xi;
  for (int n = 0; n < M; ++n) {
    double vm;
    if (FULL) {
      vm = St.V[k * M + n];
    } else {
      const int idx = lookup_smoother_hint(P, St.R[k * M + n]);
      vm = P.vmtab[(size_t)n * nr + idx];
    }
    const double mm = St.E[k * M + n];
    const double vcav = 1.0 / (1.0 / vm - alpha * St.ttau[k * M + n]);      // ihgp :407
    const double mcav = vcav * (mm / vm - alpha * St.tnu[k * M + n]);       // ihgp :408
    s_mu[n * TPB + tid] = mcav;
    s_s2[n * TPB + tid] = vcav;
  }
  const double lz = mom_thread<DP>(mp, alpha, y, s_mu + tid, s_s2 + tid, TPB, s_d1 + tid, s_d2 + tid, s_tab + tid);
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

// Deterministic sum of lZ[k0..k0+n) into out[slot] (negated if neg): one CTA per signal.
__global__ void __launch_bounds__(1024)
sum_kernel(const DevState* __restrict__ states, long long k0, long long n, double* __restrict__ out, int stride, int slot,
           int neg) {
  const double* x = states[blockIdx.x].lZ + k0;
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
