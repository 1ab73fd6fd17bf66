// Full-state Power-EP kernels (gf_ep_modulator_nmf).
//
// Reference recursion: matlab/gf_ep_modulator_nmf.m:126-184 (filter), :207-274
// (RTS smoother + site update), :384-522 (nlZ mode).  The reference's n-by-n
// covariance is block diagonal with one b-by-b block per latent and stays so
// under its own update, hence the state here is M independent
// (m_b, P_b) pairs coupled only through the moment matching.
//   * filter passes run sequentially, one warp per signal, lanes over blocks
//     (and over sigma points inside mom_warp);
//   * the RTS smoother is linear once the filtered estimates exist: a chunked
//     three-phase scan over (E, g, L) elements (Sarkka & Garcia-Fernandez 2021),
//     where phase 3 re-applies the reference's literal step inside each chunk;
//   * the smoother-side site update is independent per time step (siteupdate.cuh).
#pragma once
#include "common.cuh"
#include "mom.cuh"
#include "scan.cuh"

namespace nsagp {

// C = A * B (b-by-b, column-major, padded to BM)
template <int BM>
__device__ __forceinline__ void mat_mul(const double* A, const double* B, double* C) {
#pragma unroll
  for (int j = 0; j < BM; ++j)
#pragma unroll
    for (int i = 0; i < BM; ++i) {
      double s = 0.0;
#pragma unroll
      for (int l = 0; l < BM; ++l) s = fma(A[i + l * BM], B[l + j * BM], s);
      C[i + j * BM] = s;
    }
}

// C = A * B' + Q
template <int BM>
__device__ __forceinline__ void mat_mul_bt_add(const double* A, const double* B, const double* Q, double* C) {
#pragma unroll
  for (int j = 0; j < BM; ++j)
#pragma unroll
    for (int i = 0; i < BM; ++i) {
      double s = 0.0;
#pragma unroll
      for (int l = 0; l < BM; ++l) s = fma(A[i + l * BM], B[j + l * BM], s);
      C[i + j * BM] = s + Q[i + j * BM];
    }
}

// ---------------------------------------------------------------- filter pass
// One warp per signal, steps 0..T-1.  mom_all: moment matching at every observed
// step (first EP iteration) or only at k == T-1.  nlz: nlZ-mode update rules
// (clamp at every step, all-sites z-form when any site is at the bound, :424-439).
template <int DP, int BM>
__global__ void __launch_bounds__(32)
gfep_filter_kernel(const DevProblem* __restrict__ probs, const DevState* __restrict__ states, long long T,
                   long long k0, int mom_all, double ep_damp, int nlz, int store) {
  const DevProblem& P_ = probs[blockIdx.x];
  const DevState& St = states[blockIdx.x];
  const int lane = threadIdx.x;
  const int M = P_.M;
  const bool active = lane < M;
  const int n = active ? lane : M - 1;

  extern __shared__ double sm[];
  double* s_mu = sm;
  double* s_s2 = sm + 32;
  double* s_W = sm + 64;
  double* s_wn = s_W + DP * kNP;
  double* s_xn = s_wn + P_.S;
  for (int i = lane; i < DP * kNP; i += 32) s_W[i] = P_.W[i];
  for (int i = lane; i < P_.S; i += 32) s_wn[i] = P_.wn[i];
  for (int i = lane; i < kNP * P_.S; i += 32) s_xn[i] = P_.xn[i];
  __syncwarp();
  MomParams mp;
  mp.D = P_.D; mp.N = P_.N; mp.S = P_.S; mp.kind = P_.lik_kind; mp.sn2 = P_.sn2; mp.shift = P_.link_shift;
  mp.W = s_W; mp.wn = s_wn; mp.xn = s_xn;

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
  if (k0 > 0) {                                        // continue from the stored estimate of step k0-1
#pragma unroll
    for (int i = 0; i < BM; ++i) m[i] = (i < b) ? St.MS[(k0 - 1) * P_.n + off + i] : 0.0;
    const double* src = St.PS + ((size_t)(k0 - 1) * M + n) * BM * BM;
#pragma unroll
    for (int i = 0; i < BM * BM; ++i) P[i] = src[i];
  }

  for (long long k = k0; k < T; ++k) {
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
    const double y = St.y[k];
    if (!isnan(y)) {                                   // :135
      double fmu = 0.0, HPH = 0.0, Wv[BM], hP[BM];
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

      double tt = St.ttau[k * M + n], tn = St.tnu[k * M + n];
      if (mom_all || k == T - 1) {                     // :141
        if (active) { s_mu[lane] = fmu; s_s2[lane] = HPH; }
        __syncwarp();
        double d1, d2;
        const double lz = mom_warp<DP>(mp, 1.0, y, s_mu, s_s2, lane, d1, d2);   // alpha = 1 (:144)
        __syncwarp();
        const double den = 1.0 + d2 * HPH;
        tt = (1.0 - ep_damp) * tt + ep_damp * (-d2 / den);                      // :147
        tn = (1.0 - ep_damp) * tn + ep_damp * ((d1 - fmu * d2) / den);          // :148
        if (!nlz) tt = fmax(tt, 0.0);                                           // :151
        if (lane == 0) St.lZ[k] = lz;
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

// ------------------------------------------------------------- RTS smoother
template <int BM>
struct RtsMap {            // x_k = E x_{k+1} + g ;  P_k = E P_{k+1} E' + L
  double E[BM * BM], g[BM], L[BM * BM];
};

// Per-(block, step) smoother quantities from the filtered estimate
// (gf_ep_modulator_nmf.m:210-226): PSkp = A PSk A' + Q, G = PSk A' / PSkp.
template <int BM>
struct RtsStep {
  double G[BM * BM], PSkp[BM * BM], PSk[BM * BM], ms[BM], Ams[BM];
  bool ok;
};

// BMS: leading dimension of the blocks in HBM; BM: the size this latent family computes with (ihgp.cuh: AffineElemBase).
template <int BMS, int BM>
struct RtsElem {
  const DevProblem& P; const DevState& St; int n, off, b;
  double A[BM * BM], Q[BM * BM], hv[BM];
  __device__ RtsElem(const DevProblem& P_, const DevState& S_, int n_) : P(P_), St(S_), n(n_) {
    off = P.off[n]; b = P.off[n + 1] - off;
#pragma unroll
    for (int i = 0; i < BM * BM; ++i) {
      A[i] = P.A[n * BMS * BMS + (i % BM) + (i / BM) * BMS];
      Q[i] = P.Q[n * BMS * BMS + (i % BM) + (i / BM) * BMS];
    }
#pragma unroll
    for (int i = 0; i < BM; ++i) hv[i] = P.h[n * BMS + i];
  }
  // the filtered estimate of step k (the only HBM reads of a step: scan.cuh requests them steps ahead)
  __device__ __forceinline__ void load(long long k, double (&PSk)[BM * BM], double (&ms)[BM]) const {
    const double* src = St.PS + ((size_t)k * P.M + n) * BMS * BMS;
#pragma unroll
    for (int i = 0; i < BM * BM; ++i) PSk[i] = src[(i % BM) + (i / BM) * BMS];
#pragma unroll
    for (int i = 0; i < BM; ++i) ms[i] = (i < b) ? St.MS[k * P.n + off + i] : 0.0;
  }
  __device__ __forceinline__ void step(long long k, RtsStep<BM>& s) const {
    load(k, s.PSk, s.ms);
    compute(s);
  }
  __device__ __forceinline__ void compute(RtsStep<BM>& s) const {
    double AP[BM * BM];
    mat_mul<BM>(A, s.PSk, AP);
    mat_mul_bt_add<BM>(AP, A, Q, s.PSkp);               // :213
    // Cholesky PSkp = L L' (lower), padding rows/cols treated as identity (:216)
    double L[BM * BM], Linv[BM];
    s.ok = true;
#pragma unroll
    for (int j = 0; j < BM; ++j) {
      double d = (j < b) ? s.PSkp[j + j * BM] : 1.0;
#pragma unroll
      for (int l = 0; l < BM; ++l) if (l < j) d = fma(-L[j + l * BM], L[j + l * BM], d);
      if (!(d > 0.0) || !(d < 1e300)) { s.ok = false; d = 1.0; }
      // 1/sqrt(d) once per pivot, straight-line (fastmath.cuh, <= 1 ulp): the triangular solves below multiply by it
      // instead of dividing 6 BM times by L(j,j) -- the divisions were half of this step's instructions
      const double inv = rsqrt_fast(d);
      const double ljj = d * inv;
      Linv[j] = inv;
      L[j + j * BM] = ljj;
#pragma unroll
      for (int i = 0; i < BM; ++i) {
        if (i > j) {
          double v = (i < b && j < b) ? s.PSkp[i + j * BM] : 0.0;
#pragma unroll
          for (int l = 0; l < BM; ++l) if (l < j) v = fma(-L[i + l * BM], L[j + l * BM], v);
          L[i + j * BM] = v * inv;
        } else if (i < j) {
          L[i + j * BM] = 0.0;
        }
      }
    }
    // G = PSk*A'/L'/L  (:226): row-wise triangular solves
    double Bm[BM * BM];
#pragma unroll
    for (int j = 0; j < BM; ++j)
#pragma unroll
      for (int i = 0; i < BM; ++i) {
        double v = 0.0;
#pragma unroll
        for (int l = 0; l < BM; ++l) v = fma(s.PSk[i + l * BM], A[j + l * BM], v);   // PSk*A'
        Bm[i + j * BM] = v;
      }
#pragma unroll
    for (int i = 0; i < BM; ++i) {
      double x[BM];
#pragma unroll
      for (int j = 0; j < BM; ++j) {                    // x L' = B_i  (forward)
        double v = Bm[i + j * BM];
#pragma unroll
        for (int l = 0; l < BM; ++l) if (l < j) v = fma(-x[l], L[j + l * BM], v);
        x[j] = v * Linv[j];
      }
#pragma unroll
      for (int jj = 0; jj < BM; ++jj) {                 // g L = x  (backward)
        const int j = BM - 1 - jj;
        double v = x[j];
#pragma unroll
        for (int l = 0; l < BM; ++l) if (l > j) v = fma(-s.G[i + l * BM], L[l + j * BM], v);
        s.G[i + j * BM] = v * Linv[j];
      }
    }
#pragma unroll
    for (int i = 0; i < BM; ++i) {
      double v = 0.0;
#pragma unroll
      for (int j = 0; j < BM; ++j) v = fma(A[i + j * BM], s.ms[j], v);
      s.Ams[i] = v;
    }
  }
  // the step as an associative element: E = G, g = ms - G A ms, L = PSk - G PSkp G'
  __device__ __forceinline__ void as_map(const RtsStep<BM>& s, RtsMap<BM>& e) const {
    double T1[BM * BM];
    mat_mul<BM>(s.G, s.PSkp, T1);
#pragma unroll
    for (int j = 0; j < BM; ++j)
#pragma unroll
      for (int i = 0; i < BM; ++i) {
        double v = 0.0;
#pragma unroll
        for (int l = 0; l < BM; ++l) v = fma(T1[i + l * BM], s.G[j + l * BM], v);
        e.L[i + j * BM] = s.PSk[i + j * BM] - v;
        e.E[i + j * BM] = s.G[i + j * BM];
      }
#pragma unroll
    for (int i = 0; i < BM; ++i) {
      double v = 0.0;
#pragma unroll
      for (int j = 0; j < BM; ++j) v = fma(s.G[i + j * BM], s.Ams[j], v);
      e.g[i] = s.ms[i] - v;
    }
  }
  // the reference's literal update (:229-230): m = MSk + G (m - A MSk); P = PSk + G (P - PSkp) G'
  __device__ __forceinline__ void apply_step(const RtsStep<BM>& s, double (&m)[BM], double (&Pm)[BM * BM]) const {
    double dm[BM], Dp[BM * BM], T1[BM * BM];
#pragma unroll
    for (int i = 0; i < BM; ++i) dm[i] = m[i] - s.Ams[i];
#pragma unroll
    for (int i = 0; i < BM; ++i) {
      double v = s.ms[i];
#pragma unroll
      for (int j = 0; j < BM; ++j) v = fma(s.G[i + j * BM], dm[j], v);
      m[i] = v;
    }
#pragma unroll
    for (int i = 0; i < BM * BM; ++i) Dp[i] = Pm[i] - s.PSkp[i];
    mat_mul<BM>(s.G, Dp, T1);
#pragma unroll
    for (int j = 0; j < BM; ++j)
#pragma unroll
      for (int i = 0; i < BM; ++i) {
        double v = s.PSk[i + j * BM];
#pragma unroll
        for (int l = 0; l < BM; ++l) v = fma(T1[i + l * BM], s.G[j + l * BM], v);
        Pm[i + j * BM] = v;
      }
  }
};

template <int BM>
__device__ __forceinline__ void rts_compose(RtsMap<BM>& acc, const RtsMap<BM>& e) {
  // acc <- e o acc
  double E[BM * BM], T1[BM * BM], g[BM];
  mat_mul<BM>(e.E, acc.E, E);
  mat_mul<BM>(e.E, acc.L, T1);
#pragma unroll
  for (int i = 0; i < BM; ++i) {
    double v = e.g[i];
#pragma unroll
    for (int j = 0; j < BM; ++j) v = fma(e.E[i + j * BM], acc.g[j], v);
    g[i] = v;
  }
#pragma unroll
  for (int j = 0; j < BM; ++j)
#pragma unroll
    for (int i = 0; i < BM; ++i) {
      double v = e.L[i + j * BM];
#pragma unroll
      for (int l = 0; l < BM; ++l) v = fma(T1[i + l * BM], e.E[j + l * BM], v);
      acc.L[i + j * BM] = v;
    }
#pragma unroll
  for (int i = 0; i < BM * BM; ++i) acc.E[i] = E[i];
#pragma unroll
  for (int i = 0; i < BM; ++i) acc.g[i] = g[i];
}

template <int BM>
__device__ __forceinline__ void rts_apply_map(const RtsMap<BM>& e, double (&m)[BM], double (&Pm)[BM * BM]) {
  double r[BM], T1[BM * BM];
#pragma unroll
  for (int i = 0; i < BM; ++i) {
    double v = e.g[i];
#pragma unroll
    for (int j = 0; j < BM; ++j) v = fma(e.E[i + j * BM], m[j], v);
    r[i] = v;
  }
#pragma unroll
  for (int i = 0; i < BM; ++i) m[i] = r[i];
  mat_mul<BM>(e.E, Pm, T1);
#pragma unroll
  for (int j = 0; j < BM; ++j)
#pragma unroll
    for (int i = 0; i < BM; ++i) {
      double v = e.L[i + j * BM];
#pragma unroll
      for (int l = 0; l < BM; ++l) v = fma(T1[i + l * BM], e.E[j + l * BM], v);
      Pm[i + j * BM] = v;
    }
}

// scan.cuh element of the RTS smoother: processing order s = 0..T-2 is time k = T-2-s.
// Phase 3 applies the reference's literal step (:229-230), overwrites the stored estimates,
// emits the marginals and the convergence diagnostics (:231-234, :271-272).
template <int BM>
struct RtsState {
  double m[BM], P[BM * BM];
};

template <int BMS, int BM = BMS>
struct RtsScanElem {
  using Map = RtsMap<BM>;
  using State = RtsState<BM>;
  static constexpr int kMapDoubles = 2 * BMS * BMS + BMS;       // slot sizes are the same for both families
  static constexpr int kStateDoubles = BMS * BMS + BMS;
  RtsElem<BMS, BM> el;
  bool ok;
  double mdM, mdP;
  __device__ RtsScanElem(const DevProblem& P_, const DevState& S_, int n_, const ScanArgs&) : el(P_, S_, n_), ok(true), mdM(0.0), mdP(0.0) {}
  __device__ static __forceinline__ void compose(Map& acc, const Map& e) { rts_compose<BM>(acc, e); }
  __device__ static __forceinline__ void apply(const Map& e, State& s) { rts_apply_map<BM>(e, s.m, s.P); }
  __device__ static __forceinline__ void store_map(const Map& e, double* dst) {
#pragma unroll
    for (int i = 0; i < BM * BM; ++i) { dst[i] = e.E[i]; dst[BM * BM + BM + i] = e.L[i]; }
#pragma unroll
    for (int i = 0; i < BM; ++i) dst[BM * BM + i] = e.g[i];
  }
  __device__ static __forceinline__ void load_map(Map& e, const double* src) {
#pragma unroll
    for (int i = 0; i < BM * BM; ++i) { e.E[i] = src[i]; e.L[i] = src[BM * BM + BM + i]; }
#pragma unroll
    for (int i = 0; i < BM; ++i) e.g[i] = src[BM * BM + i];
  }
  __device__ static __forceinline__ void store_state(const State& s, double* dst) {
#pragma unroll
    for (int i = 0; i < BM; ++i) dst[i] = s.m[i];
#pragma unroll
    for (int i = 0; i < BM * BM; ++i) dst[BM + i] = s.P[i];
  }
  __device__ static __forceinline__ void load_state(State& s, const double* src) {
#pragma unroll
    for (int i = 0; i < BM; ++i) s.m[i] = src[i];
#pragma unroll
    for (int i = 0; i < BM * BM; ++i) s.P[i] = src[BM + i];
  }
  // Inputs of a step -- the filtered covariance block and mean, and (apply phase) the marginals of the previous EP
  // iteration for the max-diff diagnostics -- are requested kPrefetch steps ahead (scan.cuh: scan_walk): the apply
  // phase was bound by two exposed HBM round trips per step (issue slots 2-3 % busy, profiles/r2h_*).
  static constexpr bool kTwoTiles = false;
  static constexpr int kPrefetch = BM <= 2 ? 3 : 2;
  // rows [k_lo, k_hi) of the inputs towards L2 (scan.cuh: scan_tile_prefetch)
  __device__ static __forceinline__ void prefetch_rows(const DevProblem& P, const DevState& St, long long k_lo, long long k_hi, bool apply) {
    const size_t rows = (size_t)(k_hi - k_lo);
    l2_prefetch_bulk(St.PS + (size_t)k_lo * P.M * BMS * BMS, rows * P.M * BMS * BMS * sizeof(double));
    l2_prefetch_bulk(St.MS + k_lo * P.n, rows * P.n * sizeof(double));
    if (apply) {
      l2_prefetch_bulk(St.E + k_lo * P.M, rows * P.M * sizeof(double));
      l2_prefetch_bulk(St.V + k_lo * P.M, rows * P.M * sizeof(double));
    }
  }
  struct In { double PSk[BM * BM], ms[BM], Eold, Vold; };
  struct Tab {};
  bool applying = false;
  __device__ __forceinline__ void begin_apply() { applying = true; }
  __device__ __forceinline__ void load(long long k, In& in) const {
    el.load(k, in.PSk, in.ms);
    in.Eold = applying ? el.St.E[k * el.P.M + el.n] : 0.0;
    in.Vold = applying ? el.St.V[k * el.P.M + el.n] : 0.0;
  }
  __device__ __forceinline__ void lookup(long long, const In&, Tab&) const {}
  __device__ __forceinline__ void fill(const In& in, RtsStep<BM>& st) const {
#pragma unroll
    for (int i = 0; i < BM * BM; ++i) st.PSk[i] = in.PSk[i];
#pragma unroll
    for (int i = 0; i < BM; ++i) st.ms[i] = in.ms[i];
  }
  __device__ __forceinline__ void get(long long, const In& in, const Tab&, Map& e) {
    RtsStep<BM> st;
    fill(in, st);
    el.compute(st);
    ok = ok && st.ok;
    el.as_map(st, e);
  }
  __device__ __forceinline__ void step(long long k, const In& in, const Tab&, State& s) {
    RtsStep<BM> st;
    fill(in, st);
    el.compute(st);
    el.apply_step(st, s.m, s.P);
    const DevProblem& P = el.P; const DevState& St = el.St; const int n = el.n;
#pragma unroll
    for (int i = 0; i < BM; ++i) if (i < el.b) St.MS[k * P.n + el.off + i] = s.m[i];
    // the whole padded block (zeros outside the BM x BM corner): full 32-byte sectors, no read-modify-write in L2
    double* dst = St.PS + ((size_t)k * P.M + n) * BMS * BMS;
#pragma unroll
    for (int i = 0; i < BMS * BMS; ++i) dst[i] = ((i % BMS) < BM && (i / BMS) < BM) ? s.P[(i % BMS) + (i / BMS) * BM] : 0.0;
    double e = 0.0, v = 0.0;
#pragma unroll
    for (int i = 0; i < BM; ++i) {
      e = fma(el.hv[i], s.m[i], e);
      double g = 0.0;
#pragma unroll
      for (int j = 0; j < BM; ++j) g = fma(el.hv[j], s.P[j + i * BM], g);
      v = fma(g, el.hv[i], v);
    }
    mdM = fmax(mdM, fabs(in.Eold - e));
    mdP = fmax(mdP, fabs(in.Vold - v));
    St.E[k * P.M + n] = e;
    St.V[k * P.M + n] = v;
  }
  // the smoother starts from the filtered estimate of the last step, k = kinit = T-1
  __device__ __forceinline__ void init(State& s, int, long long kinit) {
    const DevProblem& P = el.P; const DevState& St = el.St;
#pragma unroll
    for (int i = 0; i < BM; ++i) s.m[i] = (i < el.b) ? St.MS[kinit * P.n + el.off + i] : 0.0;
    const double* src = St.PS + ((size_t)kinit * P.M + el.n) * BMS * BMS;
#pragma unroll
    for (int i = 0; i < BM * BM; ++i) s.P[i] = src[(i % BM) + (i / BM) * BMS];
  }
  __device__ __forceinline__ void store_final(const State&) {}
  __device__ __forceinline__ void finish_reduce() { if (!ok) atomicCAS(el.St.status, 0, 1); }
  __device__ __forceinline__ void finish_apply() {
    atomic_max_nonneg(el.St.maxdiff, mdM);
    atomic_max_nonneg(el.St.maxdiff + 1, mdP);
  }
};

// ------------------------------------------------------- frozen-site filter scan
// Once the sites are frozen (every filter pass after the first, except its last step), the
// block filter of gf_ep_modulator_nmf.m:126-184 is a linear Kalman filter with per-step
// pseudo-observations (precision ttau_k, natural mean tnu_k), and its steps are the elements
// a_k = (A, b, C, eta, J) of Sarkka & Garcia-Fernandez (2021), written here on the natural
// parameters so that ttau = 0 (site at the bound, or a missing sample) needs no special case:
//   z = ttau h Q h' + 1
//   A_k = (I - Q h' h ttau / z) A      b_k = Q h' tnu / z      C_k = Q - Q h' h Q ttau / z
//   eta_k = A' h' tnu / z              J_k = A' h' h A ttau / z
// The first step has no prediction: it is the element (0, m_post, P_post, 0, 0) of the prior.
template <int BM>
struct KfMap {
  double A[BM * BM], b[BM], C[BM * BM], eta[BM], J[BM * BM];
};

// X <- inverse of X (BM-by-BM, column-major), Gauss-Jordan with partial pivoting done by
// conditional row swaps so that every index is a compile-time constant (registers, no local memory).
template <int BM>
__device__ __forceinline__ void small_inverse(double (&X)[BM * BM], double (&Y)[BM * BM]) {
#pragma unroll
  for (int i = 0; i < BM * BM; ++i) Y[i] = ((i % BM) == (i / BM)) ? 1.0 : 0.0;
#pragma unroll
  for (int j = 0; j < BM; ++j) {
#pragma unroll
    for (int r = j + 1; r < BM; ++r) {
      const bool sw = fabs(X[r + j * BM]) > fabs(X[j + j * BM]);
#pragma unroll
      for (int c = 0; c < BM; ++c) {
        const double xa = X[j + c * BM], xb = X[r + c * BM];
        X[j + c * BM] = sw ? xb : xa; X[r + c * BM] = sw ? xa : xb;
        const double ya = Y[j + c * BM], yb = Y[r + c * BM];
        Y[j + c * BM] = sw ? yb : ya; Y[r + c * BM] = sw ? ya : yb;
      }
    }
    const double piv = 1.0 / X[j + j * BM];
#pragma unroll
    for (int c = 0; c < BM; ++c) { X[j + c * BM] *= piv; Y[j + c * BM] *= piv; }
#pragma unroll
    for (int r = 0; r < BM; ++r) {
      if (r != j) {
        const double f = X[r + j * BM];
#pragma unroll
        for (int c = 0; c < BM; ++c) {
          X[r + c * BM] = fma(-f, X[j + c * BM], X[r + c * BM]);
          Y[r + c * BM] = fma(-f, Y[j + c * BM], Y[r + c * BM]);
        }
      }
    }
  }
}

// y = X v ; y = X' v
template <int BM>
__device__ __forceinline__ void mat_vec(const double* X, const double* v, double* y) {
#pragma unroll
  for (int i = 0; i < BM; ++i) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < BM; ++j) s = fma(X[i + j * BM], v[j], s);
    y[i] = s;
  }
}
template <int BM>
__device__ __forceinline__ void mat_t_vec(const double* X, const double* v, double* y) {
#pragma unroll
  for (int i = 0; i < BM; ++i) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < BM; ++j) s = fma(X[j + i * BM], v[j], s);
    y[i] = s;
  }
}
// C = A' * B
template <int BM>
__device__ __forceinline__ void mat_t_mul(const double* A, const double* B, double* C) {
#pragma unroll
  for (int j = 0; j < BM; ++j)
#pragma unroll
    for (int i = 0; i < BM; ++i) {
      double s = 0.0;
#pragma unroll
      for (int l = 0; l < BM; ++l) s = fma(A[l + i * BM], B[l + j * BM], s);
      C[i + j * BM] = s;
    }
}

// acc <- e o acc  (acc earlier in time, e later): Sarkka & Garcia-Fernandez (2021), eq. for a_i (x) a_j
template <int BM>
__device__ __forceinline__ void kf_compose(KfMap<BM>& acc, const KfMap<BM>& e) {
  double X[BM * BM], Mx[BM * BM], T1[BM * BM], T2[BM * BM];
  mat_mul<BM>(acc.C, e.J, X);                               // I + C_i J_j
#pragma unroll
  for (int i = 0; i < BM; ++i) X[i + i * BM] += 1.0;
  small_inverse<BM>(X, Mx);
  double v[BM], u[BM], w[BM];
  mat_vec<BM>(acc.C, e.eta, v);                             // b_i + C_i eta_j
#pragma unroll
  for (int i = 0; i < BM; ++i) v[i] += acc.b[i];
  mat_vec<BM>(Mx, v, u);
  double bn[BM];
  mat_vec<BM>(e.A, u, bn);
  mat_vec<BM>(e.J, acc.b, v);                               // eta_j - J_j b_i
#pragma unroll
  for (int i = 0; i < BM; ++i) v[i] = e.eta[i] - v[i];
  mat_t_vec<BM>(Mx, v, w);                                  // (I + J_j C_i)^-1 = Mx'
  double en[BM];
  mat_t_vec<BM>(acc.A, w, en);
  mat_mul<BM>(Mx, acc.C, T1);                               // C_ij = A_j Mx C_i A_j' + C_j
  mat_mul<BM>(e.A, T1, T2);
  double Cn[BM * BM];
  mat_mul_bt_add<BM>(T2, e.A, e.C, Cn);
  mat_mul<BM>(e.J, acc.A, T1);                              // J_ij = A_i' Mx' J_j A_i + J_i
  mat_t_mul<BM>(Mx, T1, T2);
  double Jn[BM * BM];
  mat_t_mul<BM>(acc.A, T2, Jn);
  mat_mul<BM>(Mx, acc.A, T1);                               // A_ij = A_j Mx A_i
  mat_mul<BM>(e.A, T1, T2);
#pragma unroll
  for (int i = 0; i < BM * BM; ++i) { acc.A[i] = T2[i]; acc.C[i] = Cn[i]; acc.J[i] = Jn[i] + acc.J[i]; }
#pragma unroll
  for (int i = 0; i < BM; ++i) { acc.b[i] = bn[i] + e.b[i]; acc.eta[i] = en[i] + acc.eta[i]; }
}

// (m, P) <- the filtered moments after the steps of e, given the moments (m, P) before them.
template <int BM>
__device__ __forceinline__ void kf_apply(const KfMap<BM>& e, double (&m)[BM], double (&Pm)[BM * BM]) {
  double X[BM * BM], Mx[BM * BM], T1[BM * BM], T2[BM * BM], v[BM], u[BM];
  mat_mul<BM>(Pm, e.J, X);
#pragma unroll
  for (int i = 0; i < BM; ++i) X[i + i * BM] += 1.0;
  small_inverse<BM>(X, Mx);
  mat_vec<BM>(Pm, e.eta, v);
#pragma unroll
  for (int i = 0; i < BM; ++i) v[i] += m[i];
  mat_vec<BM>(Mx, v, u);
  mat_vec<BM>(e.A, u, v);
#pragma unroll
  for (int i = 0; i < BM; ++i) m[i] = v[i] + e.b[i];
  mat_mul<BM>(Mx, Pm, T1);
  mat_mul<BM>(e.A, T1, T2);
  mat_mul_bt_add<BM>(T2, e.A, e.C, Pm);
}

template <int BMS, int BM = BMS>
struct KfScanElem {
  using Map = KfMap<BM>;
  using State = RtsState<BM>;
  static constexpr int kMapDoubles = 3 * BMS * BMS + 2 * BMS;   // slot sizes are the same for both families
  static constexpr int kStateDoubles = BMS * BMS + BMS;
  const DevProblem& P; const DevState& St; int n, off, b; bool nlz;
  double A[BM * BM], Q[BM * BM], hv[BM], hA[BM], Qh[BM], hQh;
  __device__ KfScanElem(const DevProblem& P_, const DevState& S_, int n_, const ScanArgs& a)
      : P(P_), St(S_), n(n_), nlz((a.flags & 1) != 0) {
    off = P.off[n]; b = P.off[n + 1] - off;
#pragma unroll
    for (int i = 0; i < BM * BM; ++i) {
      A[i] = P.A[n * BMS * BMS + (i % BM) + (i / BM) * BMS];
      Q[i] = P.Q[n * BMS * BMS + (i % BM) + (i / BM) * BMS];
    }
#pragma unroll
    for (int i = 0; i < BM; ++i) { hv[i] = P.h[n * BMS + i]; hA[i] = P.hA[n * BMS + i]; }
    mat_vec<BM>(Q, hv, Qh);
    hQh = 0.0;
#pragma unroll
    for (int i = 0; i < BM; ++i) hQh = fma(hv[i], Qh[i], hQh);
  }
  __device__ static __forceinline__ void compose(Map& acc, const Map& e) { kf_compose<BM>(acc, e); }
  __device__ static __forceinline__ void apply(const Map& e, State& s) { kf_apply<BM>(e, s.m, s.P); }
  __device__ static __forceinline__ void store_map(const Map& e, double* d) {
#pragma unroll
    for (int i = 0; i < BM * BM; ++i) { d[i] = e.A[i]; d[BM * BM + i] = e.C[i]; d[2 * BM * BM + i] = e.J[i]; }
#pragma unroll
    for (int i = 0; i < BM; ++i) { d[3 * BM * BM + i] = e.b[i]; d[3 * BM * BM + BM + i] = e.eta[i]; }
  }
  __device__ static __forceinline__ void load_map(Map& e, const double* d) {
#pragma unroll
    for (int i = 0; i < BM * BM; ++i) { e.A[i] = d[i]; e.C[i] = d[BM * BM + i]; e.J[i] = d[2 * BM * BM + i]; }
#pragma unroll
    for (int i = 0; i < BM; ++i) { e.b[i] = d[3 * BM * BM + i]; e.eta[i] = d[3 * BM * BM + BM + i]; }
  }
  __device__ static __forceinline__ void store_state(const State& s, double* dst) { RtsScanElem<BMS, BM>::store_state(s, dst); }
  __device__ static __forceinline__ void load_state(State& s, const double* src) { RtsScanElem<BMS, BM>::load_state(s, src); }
  // sites of step k as the frozen pass sees them (:159-176; nlZ mode clamps at every step, :425)
  __device__ __forceinline__ bool sites(long long k, double& tt, double& tn) const {
    if (isnan(St.y[k])) { tt = 0.0; tn = 0.0; return false; }                  // :135 no update at all
    tt = St.ttau[k * P.M + n]; tn = St.tnu[k * P.M + n];
    if (nlz) tt = fmax(tt, 0.0);
    return true;
  }
  __device__ static __forceinline__ void prefetch_rows(const DevProblem& P, const DevState& St, long long k_lo, long long k_hi, bool) {
    const size_t bytes = (size_t)(k_hi - k_lo) * P.M * sizeof(double);
    l2_prefetch_bulk(St.ttau + k_lo * P.M, bytes);
    l2_prefetch_bulk(St.tnu + k_lo * P.M, bytes);
  }
  // (no register-pipelined inputs for this element: scan.cuh scan_walk)
  static constexpr bool kTwoTiles = false;
  static constexpr int kPrefetch = 1;
  struct In {};
  struct Tab {};
  __device__ __forceinline__ void load(long long, In&) const {}
  __device__ __forceinline__ void lookup(long long, const In&, Tab&) const {}
  __device__ __forceinline__ void get(long long k, const In&, const Tab&, Map& e) { get(k, e); }
  __device__ __forceinline__ void step(long long k, const In&, const Tab&, State& s) { step(k, s); }
  __device__ __forceinline__ void begin_apply() {}
  __device__ __forceinline__ void get(long long k, Map& e) {
    double tt, tn;
    sites(k, tt, tn);
    if (k == 0) {                                           // prior (0, Pinf), update only (:116-117,:129)
      double W0[BM], s0 = 0.0;
      const double* Pi = P.Pinf + n * BMS * BMS;
#pragma unroll
      for (int i = 0; i < BM; ++i) {
        double w = 0.0;
#pragma unroll
        for (int j = 0; j < BM; ++j) w = fma(Pi[i + j * BMS], hv[j], w);
        W0[i] = w;
      }
#pragma unroll
      for (int i = 0; i < BM; ++i) s0 = fma(hv[i], W0[i], s0);
      const double rz = 1.0 / fma(tt, s0, 1.0);
#pragma unroll
      for (int j = 0; j < BM; ++j)
#pragma unroll
        for (int i = 0; i < BM; ++i) {
          e.A[i + j * BM] = 0.0; e.J[i + j * BM] = 0.0;
          e.C[i + j * BM] = fma(-(W0[i] * (tt * rz)), W0[j], Pi[i + j * BMS]);
        }
#pragma unroll
      for (int i = 0; i < BM; ++i) { e.b[i] = W0[i] * (tn * rz); e.eta[i] = 0.0; }
      return;
    }
    const double rz = 1.0 / fma(tt, hQh, 1.0);
    const double g = tt * rz, c = tn * rz;
#pragma unroll
    for (int j = 0; j < BM; ++j)
#pragma unroll
      for (int i = 0; i < BM; ++i) {
        e.A[i + j * BM] = fma(-(Qh[i] * g), hA[j], A[i + j * BM]);
        e.C[i + j * BM] = fma(-(Qh[i] * g), Qh[j], Q[i + j * BM]);
        e.J[i + j * BM] = (hA[i] * g) * hA[j];
      }
#pragma unroll
    for (int i = 0; i < BM; ++i) { e.b[i] = Qh[i] * c; e.eta[i] = hA[i] * c; }
  }
  // the reference's literal step: predict (k > 0), z-form update, store (:129-182)
  __device__ __forceinline__ void step(long long k, State& s) {
    if (k > 0) {
      double t[BM], AP[BM * BM];
      mat_vec<BM>(A, s.m, t);
#pragma unroll
      for (int i = 0; i < BM; ++i) s.m[i] = t[i];
      mat_mul<BM>(A, s.P, AP);
      mat_mul_bt_add<BM>(AP, A, Q, s.P);
    }
    double tt, tn;
    if (sites(k, tt, tn)) {
      double fmu = 0.0, HPH = 0.0, Wv[BM], hP[BM];
#pragma unroll
      for (int i = 0; i < BM; ++i) {
        fmu = fma(hv[i], s.m[i], fmu);
        double w = 0.0, g = 0.0;
#pragma unroll
        for (int j = 0; j < BM; ++j) {
          w = fma(s.P[i + j * BM], hv[j], w);
          g = fma(hv[j], s.P[j + i * BM], g);
        }
        Wv[i] = w; hP[i] = g;
      }
#pragma unroll
      for (int i = 0; i < BM; ++i) HPH = fma(hP[i], hv[i], HPH);
      if (nlz && !(HPH > 0.0)) atomicCAS(St.status, 0, 2);                     // `keyboard` trap (:408-410)
      if (nlz) St.ttau[k * P.M + n] = tt;                                      // the clamp is stored (:425)
      const double rz = 1.0 / fma(tt, HPH, 1.0);
      const double c = (tn - tt * fmu) * rz, g = tt * rz;
#pragma unroll
      for (int i = 0; i < BM; ++i) s.m[i] = fma(Wv[i], c, s.m[i]);
#pragma unroll
      for (int j = 0; j < BM; ++j)
#pragma unroll
        for (int i = 0; i < BM; ++i) s.P[i + j * BM] = fma(-(Wv[i] * g), hP[j], s.P[i + j * BM]);
    }
#pragma unroll
    for (int i = 0; i < BM; ++i) if (i < b) St.MS[k * P.n + off + i] = s.m[i];
    // the whole padded block (zeros outside the BM x BM corner): full 32-byte sectors, no read-modify-write in L2
    double* dst = St.PS + ((size_t)k * P.M + n) * BMS * BMS;
#pragma unroll
    for (int i = 0; i < BMS * BMS; ++i) dst[i] = ((i % BMS) < BM && (i / BMS) < BM) ? s.P[(i % BMS) + (i / BMS) * BM] : 0.0;
  }
  __device__ __forceinline__ void init(State& s, int, long long) {            // (0, Pinf); step 0 does not predict
#pragma unroll
    for (int i = 0; i < BM; ++i) s.m[i] = 0.0;
#pragma unroll
    for (int i = 0; i < BM * BM; ++i) s.P[i] = P.Pinf[n * BMS * BMS + (i % BM) + (i / BM) * BMS];
  }
  __device__ __forceinline__ void store_final(const State&) {}
  __device__ __forceinline__ void finish_reduce() {}
  __device__ __forceinline__ void finish_apply() {}
};

}  // namespace nsagp
