// Energy of the extended Kalman filter AND its analytic gradient by the sensitivity equations:
// matlab/gf_giekf_modulator_nmf.m:296-437 with GradObj = 'on' (the only analytic-gradient path of the reference;
// SURVEY.md section 8 row a13).
//
// The reference carries, next to (m, P), one pair (dm_j, dP_j) per hyper-parameter j and pushes each through the
// derivative of the prediction (:344-369) and of the update (:405-423) with dense n x n products (nparam * O(n^3)
// per step).  Two facts make the work per parameter O(n^2 b) instead:
//   * every parameter belongs to ONE latent, so dA_j and dQ_j are a single b x b diagonal block (ss_modulators_nmf.m
//     builds dF / dPinf block by block); the prediction of dP_j is then, tile by tile over pairs of latents (I, J),
//       dP_IJ <- (A_I dP_IJ + [I = l] dA P_IJ) A_J' + [J = l] (A_I P_IJ) dA' + [I = J = l] dQ ;
//   * the measurement is scalar, so K, dK are vectors and the update of dP_j is a symmetric rank-2 correction
//       dP <- dP - S (dK K' + K dK') - dS K K' .
// Parallelisation: the parameters are independent given (m, P), so ONE CTA PER PARAMETER runs the whole recursion
// with P and dP_j in its shared memory (2 n^2 doubles), recomputing the shared filter quantities itself -- no
// traffic and no synchronisation between CTAs; nparam = 1 + 3D + 2N <= 148 CTAs is one wave on the GPU.  CTA 0
// also reports the energy.  As in the filter kernel (ekf.cuh) the rank corrections of a step are applied when the
// next step's tile pass loads the tile ("pending update"), which saves one sweep over the two matrices.
//
// The Jacobian / Hessian of h(x) = (H_z x)' W softplus(H_g x) are chained through H.  The reference scatters them to
// the columns with sum(H,1) == 1 (:452, :470), which is the same thing while the observed components of H are 1 and
// ill-defined once balancing (:78) has rescaled them.
#pragma once
#include "ekf.cuh"

namespace nsagp {

constexpr int kEgMaxThreads = 320;
template <int BM> struct EgCfg { static constexpr int TH = BM <= 4 ? 320 : 256; };     // register budget of the tile pass

struct EkfGradArgs {
  EkfArgs ekf;             // model, y, T, sigma2; edata / status used, MS / PS not
  int nparam;
  const int* latent;       // [nparam] latent whose block the parameter lives in, -1 = none (the noise variance)
  const double* dA;        // [nparam][BM*BM] block of d expm(F) / d theta_j     (:328-338, lower-left of AA)
  const double* dQ;        // [nparam][BM*BM] block of dQ_j                      (:362-364, constant in time)
  const double* dP0;       // [nparam][BM*BM] block of dPinf_j                   (:314)
  const double* dR;        // [nparam]                                           (:96)
  double* gdata;           // [nparam] gradient BEFORE the log-scale factor of :432-433
  double* dP_hbm;          // [nparam][n*n] or null: dP_j kept in HBM / L2 when 2 n^2 doubles exceed the shared memory (n > ~115)
};

inline size_t ekf_grad_smem_doubles(int n, int M, int BM, int D, int N, bool dp_in_smem = true) {
  const int parts = std::max(1, kEgMaxThreads / n);
  return (dp_in_smem ? 2 : 1) * (size_t)n * n + 2 * (size_t)M * BM * BM + 2 * BM * BM + (size_t)M * BM + (size_t)D * N + 13 * (size_t)n +
         3 * (size_t)parts * n + 5 * (size_t)M + 5 * (size_t)N + 5 * 10 + 8;
}

// One (latent I, latent J) tile of exact shape NI x NJ, I >= J: finish the pending rank corrections of the last update,
// predict P and dP in registers, write the tile and (off the diagonal) its transpose.
template <int NI, int NJ, int BM>
__device__ __forceinline__ void eg_tile(double* P, double* dP, int n, int oi, int oj, const double* Ai, const double* Aj,
                                        const double* Qi, const double* sdA, const double* sdQ, bool li, bool lj,
                                        const double* Ks, const double* Kv, const double* dKv, const double* Kd, bool mirror) {
  double X[NI * NJ], DX[NI * NJ], T1[NI * NJ], T2[NI * NJ];
  double ksi[NI], dki[NI], kdi[NI];
#pragma unroll
  for (int r = 0; r < NI; ++r) { ksi[r] = Ks[oi + r]; dki[r] = dKv[oi + r]; kdi[r] = Kd[oi + r]; }
#pragma unroll
  for (int c = 0; c < NJ; ++c) {
    const double ksl = Ks[oj + c], kvl = Kv[oj + c], dkl = dKv[oj + c];
#pragma unroll
    for (int r = 0; r < NI; ++r) {
      const int idx = (oi + r) + (oj + c) * n;
      X[r + c * NI] = fma(-ksi[r], kvl, P[idx]);                                            // P - K S K' (:427)
      DX[r + c * NI] = dP[idx] - fma(dki[r], ksl, fma(ksi[r], dkl, kdi[r] * kvl));          // :419-420
    }
  }
#pragma unroll
  for (int c = 0; c < NJ; ++c)
#pragma unroll
    for (int r = 0; r < NI; ++r) {
      double s1 = 0.0, s2 = 0.0;
#pragma unroll
      for (int l = 0; l < NI; ++l) { s1 = fma(Ai[r + l * BM], X[l + c * NI], s1); s2 = fma(Ai[r + l * BM], DX[l + c * NI], s2); }
      T1[r + c * NI] = s1; T2[r + c * NI] = s2;
    }
  if (li) {                                                                                 // dA P A' (:366)
#pragma unroll
    for (int c = 0; c < NJ; ++c)
#pragma unroll
      for (int r = 0; r < NI; ++r) {
        double s2 = T2[r + c * NI];
#pragma unroll
        for (int l = 0; l < NI; ++l) s2 = fma(sdA[r + l * BM], X[l + c * NI], s2);
        T2[r + c * NI] = s2;
      }
  }
#pragma unroll
  for (int c = 0; c < NJ; ++c)
#pragma unroll
    for (int r = 0; r < NI; ++r) {
      double s1 = Qi ? Qi[r + c * BM] : 0.0, s2 = (li && lj) ? sdQ[r + c * BM] : 0.0;
#pragma unroll
      for (int l = 0; l < NJ; ++l) { s1 = fma(T1[r + l * NI], Aj[c + l * BM], s1); s2 = fma(T2[r + l * NI], Aj[c + l * BM], s2); }
      if (lj) {                                                                             // (dA P A')' (:367)
#pragma unroll
        for (int l = 0; l < NJ; ++l) s2 = fma(T1[r + l * NI], sdA[c + l * BM], s2);
      }
      const int idx = (oi + r) + (oj + c) * n;
      P[idx] = s1; dP[idx] = s2;
      if (mirror) { const int idt = (oj + c) + (oi + r) * n; P[idt] = s1; dP[idt] = s2; }
    }
}

// BZ, BG > 0: the model's exact block sizes -- tiles of compile-time shape, only the tiles on and below the block
// diagonal computed (P and dP_j are symmetric) and mirrored; BZ = BG = 0: any block sizes <= BM, padded tiles with
// run-time guards, all M^2 tiles.
template <int BZ, int BG, int BM>
__global__ void __launch_bounds__(EgCfg<BM>::TH, 1) giekf_grad_kernel(EkfGradArgs g) {
  const EkfArgs& a = g.ekf;
  const int tid = threadIdx.x, nth = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int D = a.D, N = a.N, M = a.M, n = a.n;
  const int j = blockIdx.x, ell = g.latent[j];
  const double dRj = g.dR[j];
  const int parts = max(1, nth / n);
  extern __shared__ __align__(16) double sm[];
  double* P = sm;                              // [n*n] column-major
  // dP_j [n*n]: shared memory, or (state dimensions whose 2 n^2 doubles do not fit) this parameter's slice of an HBM
  // scratch -- 150 KB at n = 137, re-read every step, so it lives in L2; same code, slower
  double* dP = g.dP_hbm ? g.dP_hbm + (size_t)j * n * n : P + (size_t)n * n;
  double* sA = P + (size_t)(g.dP_hbm ? 1 : 2) * n * n;   // [M][BM*BM]
  double* sQ = sA + M * BM * BM;
  double* sdA = sQ + M * BM * BM;              // [BM*BM]
  double* sdQ = sdA + BM * BM;
  double* sh = sdQ + BM * BM;                  // [M][BM]
  double* sW = sh + M * BM;                    // [D][N] row-major
  double* mb = sW + D * N;                     // [2][n]   mean, double-buffered over the prediction
  double* dmb = mb + 2 * n;                    // [2][n]
  double* JH = dmb + 2 * n;                    // [n]  Jacobian row (:377)
  double* cv = JH + n;                         // [n]  dm' * d2h (:408)
  double* Ph = cv + n;                         // [n]  P JH'  (= K S)
  double* qv = Ph + n;                         // [n]  dP JH'
  double* Pc = qv + n;                         // [n]  P cv'
  double* Kv = Pc + n;                         // [n]  K          \  rank corrections of the last update, applied by the
  double* Ks = Kv + n;                         // [n]  K S        |  next tile pass
  double* dKv = Ks + n;                        // [n]  dK         |
  double* Kd = dKv + n;                        // [n]  dS K       /
  double* part = Kd + n;                       // [parts][3][n]
  double* fv = part + 3 * parts * n;           // [M]  H m
  double* uv = fv + M;                         // [M]  H dm
  double* coef = uv + M;                       // [M]  dh / d(latent value)
  double* rv = coef + M;                       // [M]  (d2h u) per latent
  double* mup = rv + M;                        // [M]  terms of h(m), first D used
  double* spv = mup + M;                       // [N]  link
  double* dlv = spv + N;                       // [N]  dlink
  double* d2v = dlv + N;                       // [N]  d2link
  double* zWv = d2v + N;                       // [N]  z'W
  double* uWv = zWv + N;                       // [N]  (H_z dm)'W
  double* red = uWv + N;                       // [5][10]
  __shared__ int s_blk[160];
  __shared__ int s_off[kMaxSites + 1];
  __shared__ unsigned short s_pair[(kMaxSites * (kMaxSites + 1)) / 2 + 96];
  // pair list of the exact-shape path: [g,g lower incl. diagonal | g,z | z,z lower incl. diagonal], every class
  // starting at a multiple of 32 so that no warp mixes tile shapes (0xffff = padding entry)
  const int nzz = D * (D + 1) / 2, ngz = N * D, ngg = N * (N + 1) / 2;
  const int o_gz = (ngg + 31) & ~31, o_zz = o_gz + ((ngz + 31) & ~31);
  const int npairs = BZ > 0 ? o_zz + nzz : M * M;
  if (BZ > 0) {
    for (int p = tid; p < npairs; p += nth) {
      int bi = 255, bj = 255;
      if (p < ngg) { int ii = 0; while ((ii + 1) * (ii + 2) / 2 <= p) ++ii; bi = D + ii; bj = D + p - ii * (ii + 1) / 2; }
      else if (p >= o_gz && p < o_gz + ngz) { const int u = p - o_gz; bi = D + u / D; bj = u % D; }
      else if (p >= o_zz) { const int u = p - o_zz; bi = 0; while ((bi + 1) * (bi + 2) / 2 <= u) ++bi; bj = u - bi * (bi + 1) / 2; }
      s_pair[p] = (unsigned short)(bi | (bj << 8));
    }
  }

  for (int i = tid; i < n * n; i += nth) { P[i] = a.Pinf[i]; dP[i] = 0.0; }    // :312-314
  for (int i = tid; i < M * BM * BM; i += nth) { sA[i] = a.A[i]; sQ[i] = a.Q[i]; }
  for (int i = tid; i < BM * BM; i += nth) { sdA[i] = g.dA[(size_t)j * BM * BM + i]; sdQ[i] = g.dQ[(size_t)j * BM * BM + i]; }
  for (int i = tid; i < M * BM; i += nth) sh[i] = a.h[i];
  for (int i = tid; i < D * N; i += nth) sW[i] = a.W[i];
  for (int i = tid; i < 2 * n; i += nth) { mb[i] = 0.0; dmb[i] = 0.0; }
  for (int i = tid; i < n; i += nth) { Kv[i] = 0.0; Ks[i] = 0.0; dKv[i] = 0.0; Kd[i] = 0.0; }
  for (int b = tid; b <= M; b += nth) s_off[b] = a.off[b];
  for (int b = tid; b < M; b += nth)
    for (int i = a.off[b]; i < a.off[b + 1]; ++i) s_blk[i] = b;
  __syncthreads();
  if (ell >= 0) {
    const int o = s_off[ell], nb = s_off[ell + 1] - o;
    for (int i = tid; i < nb * nb; i += nth) dP[(o + i % nb) + (size_t)(o + i / nb) * n] = g.dP0[(size_t)j * BM * BM + (i % nb) + (i / nb) * BM];
  }
  int my_b = 0, my_o = 0;
  double my_h = 0.0;
  if (tid < n) { my_b = s_blk[tid]; my_o = s_off[my_b]; my_h = sh[my_b * BM + (tid - my_o)]; }
  const int row = tid % n, prt = tid / n;
  double* m = mb;
  double* m2 = mb + n;
  double* dm = dmb;
  double* dm2 = dmb + n;
  double e_acc = 0.0, g_acc = 0.0;
  bool bad = false;
  __syncthreads();

  for (long long k = 0; k < a.T; ++k) {
    // ---- prediction (:344-373): tiles of P and dP (pending rank corrections folded into the load), mean and dm
    if (BZ > 0) {
#pragma unroll 1
      for (int p = tid; p < npairs; p += nth) {
        const int pr = s_pair[p], bi = pr & 255, bj = pr >> 8;
        if (bi == 255) continue;
        const int oi = s_off[bi], oj = s_off[bj];
        const double* Ai = sA + bi * BM * BM;
        const double* Aj = sA + bj * BM * BM;
        const double* Qi = (bi == bj) ? sQ + bi * BM * BM : nullptr;
        const bool li = bi == ell, lj = bj == ell;
        if (p < o_gz) eg_tile<(BG > 0 ? BG : 1), (BG > 0 ? BG : 1), BM>(P, dP, n, oi, oj, Ai, Aj, Qi, sdA, sdQ, li, lj, Ks, Kv, dKv, Kd, bi != bj);
        else if (p < o_zz) eg_tile<(BG > 0 ? BG : 1), (BZ > 0 ? BZ : 1), BM>(P, dP, n, oi, oj, Ai, Aj, Qi, sdA, sdQ, li, lj, Ks, Kv, dKv, Kd, true);
        else eg_tile<(BZ > 0 ? BZ : 1), (BZ > 0 ? BZ : 1), BM>(P, dP, n, oi, oj, Ai, Aj, Qi, sdA, sdQ, li, lj, Ks, Kv, dKv, Kd, bi != bj);
      }
    } else
#pragma unroll 1
    for (int p = tid; p < M * M; p += nth) {
      const int bi = p % M, bj = p / M;
      const int oi = s_off[bi], ni = s_off[bi + 1] - oi;
      const int oj = s_off[bj], nj = s_off[bj + 1] - oj;
      const double* Ai = sA + bi * BM * BM;
      const double* Aj = sA + bj * BM * BM;
      const bool li = bi == ell, lj = bj == ell;
      double accP[BM * BM], accD[BM * BM];
#pragma unroll
      for (int i = 0; i < BM * BM; ++i) {
        accP[i] = (bi == bj) ? sQ[bi * BM * BM + i] : 0.0;
        accD[i] = (li && lj) ? sdQ[i] : 0.0;
      }
#pragma unroll
      for (int l = 0; l < BM; ++l) {
        if (l < nj) {
          double x[BM], dx[BM], t1[BM], t2[BM];
          const double ksl = Ks[oj + l], kvl = Kv[oj + l], dkl = dKv[oj + l];
#pragma unroll
          for (int r = 0; r < BM; ++r) {
            if (r < ni) {
              const size_t idx = (oi + r) + (size_t)(oj + l) * n;
              x[r] = fma(-Ks[oi + r], kvl, P[idx]);                                          // P - K S K' (:427)
              dx[r] = dP[idx] - fma(dKv[oi + r], ksl, fma(Ks[oi + r], dkl, Kd[oi + r] * kvl)); // :419-420
            } else {
              x[r] = 0.0; dx[r] = 0.0;
            }
          }
#pragma unroll
          for (int r = 0; r < BM; ++r) {
            double s1 = 0.0, s2 = 0.0;
#pragma unroll
            for (int s = 0; s < BM; ++s) { s1 = fma(Ai[r + s * BM], x[s], s1); s2 = fma(Ai[r + s * BM], dx[s], s2); }
            if (li) {
#pragma unroll
              for (int s = 0; s < BM; ++s) s2 = fma(sdA[r + s * BM], x[s], s2);
            }
            t1[r] = s1; t2[r] = s2;
          }
#pragma unroll
          for (int c = 0; c < BM; ++c) {
            const double ajc = Aj[c + l * BM], dac = lj ? sdA[c + l * BM] : 0.0;
#pragma unroll
            for (int r = 0; r < BM; ++r) {
              accP[r + c * BM] = fma(t1[r], ajc, accP[r + c * BM]);
              accD[r + c * BM] = fma(t2[r], ajc, fma(t1[r], dac, accD[r + c * BM]));
            }
          }
        }
      }
#pragma unroll
      for (int c = 0; c < BM; ++c)
#pragma unroll
        for (int r = 0; r < BM; ++r)
          if (r < ni && c < nj) {
            const size_t idx = (oi + r) + (size_t)(oj + c) * n;
            P[idx] = accP[r + c * BM];
            dP[idx] = accD[r + c * BM];
          }
    }
    if (tid < n) {                                                             // :347-350 (AA * [m; dm])
      const int nb = s_off[my_b + 1] - my_o, r = tid - my_o;
      double mv = 0.0, dv = 0.0;
#pragma unroll 1
      for (int c = 0; c < nb; ++c) {
        const double av = sA[my_b * BM * BM + r + c * BM];
        mv = fma(av, m[my_o + c], mv);
        dv = fma(av, dm[my_o + c], dv);
        if (my_b == ell) dv = fma(sdA[r + c * BM], m[my_o + c], dv);
      }
      m2[tid] = mv; dm2[tid] = dv;
    }
    { double* t = m; m = m2; m2 = t; t = dm; dm = dm2; dm2 = t; }
    __syncthreads();
    // ---- measurement model (:376-378)
    if (tid < M) {
      const int o = s_off[tid], nb = s_off[tid + 1] - o;
      double f = 0.0, u = 0.0;
      for (int c = 0; c < nb; ++c) { f = fma(sh[tid * BM + c], m[o + c], f); u = fma(sh[tid * BM + c], dm[o + c], u); }
      fv[tid] = f; uv[tid] = u;
    }
    __syncthreads();
    if (tid < N) {
      const double eg = exp_fast(fv[D + tid]);
      const double dl = eg * rcp_fast(eg + 1.0);
      spv[tid] = log_ge1_fast(1.0 + eg);
      dlv[tid] = dl;
      d2v[tid] = dl * (1.0 - dl);
      double zw = 0.0, uw = 0.0;
      for (int d = 0; d < D; ++d) { zw = fma(fv[d], sW[d * N + tid], zw); uw = fma(uv[d], sW[d * N + tid], uw); }
      zWv[tid] = zw; uWv[tid] = uw;
    }
    __syncthreads();
    if (tid < M) {
      if (tid < D) {
        double wl = 0.0, r = 0.0;
        for (int q = 0; q < N; ++q) { wl = fma(sW[tid * N + q], spv[q], wl); r = fma(sW[tid * N + q] * dlv[q], uv[D + q], r); }
        coef[tid] = wl; rv[tid] = r; mup[tid] = fv[tid] * wl;
      } else {
        const int q = tid - D;
        coef[tid] = zWv[q] * dlv[q];
        rv[tid] = fma(dlv[q], uWv[q], zWv[q] * d2v[q] * uv[tid]);
      }
    }
    __syncthreads();
    if (tid < n) { JH[tid] = coef[my_b] * my_h; cv[tid] = rv[my_b] * my_h; }
    __syncthreads();
    // ---- P JH', dP JH', P cv'
    if (prt < parts) {
      double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll 4
      for (int c = prt; c < n; c += parts) {
        const double pv = P[row + (size_t)c * n];
        s0 = fma(pv, JH[c], s0);
        s1 = fma(dP[row + (size_t)c * n], JH[c], s1);
        s2 = fma(pv, cv[c], s2);
      }
      part[(prt * 3 + 0) * n + row] = s0; part[(prt * 3 + 1) * n + row] = s1; part[(prt * 3 + 2) * n + row] = s2;
    }
    __syncthreads();
    double r5[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    double ph = 0.0, q_ = 0.0, pc = 0.0;
    if (tid < n) {
      for (int pp = 0; pp < parts; ++pp) { ph += part[(pp * 3 + 0) * n + tid]; q_ += part[(pp * 3 + 1) * n + tid]; pc += part[(pp * 3 + 2) * n + tid]; }
      r5[0] = JH[tid] * ph;              // JH P JH'
      r5[1] = cv[tid] * ph;              // dmdJH P JH'
      r5[2] = JH[tid] * q_;              // JH dP JH'
      r5[3] = JH[tid] * dm[tid];         // JH dm
    }
    if (tid < D) r5[4] = mup[tid];       // h(m)
#pragma unroll
    for (int i = 0; i < 5; ++i) {
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) r5[i] += __shfl_xor_sync(0xffffffffu, r5[i], o);
      if (lane == 0) red[i * 10 + warp] = r5[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      double t = 0.0;
      for (int w = 0; w < (nth >> 5); ++w) t += red[i * 10 + w];
      r5[i] = t;
    }
    const double S = r5[0] + a.sigma2;                                         // :381
    const double dS = 2.0 * r5[1] + r5[2] + dRj;                               // :411
    const double v = a.y[k] - r5[4];                                           // :400
    const double hdm = r5[3];
    if (!(S > 0.0)) bad = true;                                                // :384-395
    const double iS = 1.0 / S, vtiS = v * iS;
    if (tid == 0) {
      g_acc += 0.5 * iS * dS - hdm * vtiS - 0.5 * vtiS * dS * vtiS;            // :414-418
      e_acc += 0.5 * log(2.0 * 3.14159265358979323846) + log(sqrt(S)) + 0.5 * vtiS * v;   // :426
    }
    if (tid < n) {
      const double K = ph * iS;                                                // :398-399
      const double dK = (q_ + pc - K * dS) * iS;                               // :421
      dm[tid] = dm[tid] + dK * v - K * hdm;                                    // :422
      m[tid] = fma(K, v, m[tid]);                                              // :429
      Kv[tid] = K; Ks[tid] = K * S; dKv[tid] = dK; Kd[tid] = dS * K;
    }
    __syncthreads();
  }
  if (tid == 0) {
    const bool nan_out = bad || isnan(e_acc);
    g.gdata[j] = nan_out ? NAN : g_acc;
    if (j == 0) { a.edata[0] = nan_out ? NAN : e_acc; if (nan_out) atomicCAS(a.status, 0, 3); }
  }
}

}  // namespace nsagp
