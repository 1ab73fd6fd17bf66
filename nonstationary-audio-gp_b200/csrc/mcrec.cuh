// Monte-Carlo reconstruction of the signal and of the NMF components from the posterior marginals
// (matlab/demo_toy_modulators_nmf.m:119-165; "sqrt" model: experiments/missing_data_music.m:138-176):
//   sub_samp(i,k,v) = Eft(i,k)   + sqrt(Varft(i,k))   * z          i = 1..D
//   mod_samp(j,k,v) = Eft(D+j,k) + sqrt(Varft(D+j,k)) * z          j = 1..N
//   Eft_mod(j,k), Varft_mod(j,k) = mean / var over v of link(mod_samp(j,k,v))
//   sig_samp(k,v) = sum_d (W link(mod_samp(:,k,v)))_d [or its sqrt] * sub_samp(d,k,v);  Esig, Vsig = mean / var over v
// Independent per time step: one thread per step streams its s samples (the explicit normal draws Z are read
// coalesced, T fastest), or generates them with a counter-based Philox4x32-10 + Box-Muller generator, in which
// case no memory is read beyond the 2M marginals of the step.  Mean and variance by Welford's update
// (the reference's two-pass var agrees to rounding).
#pragma once
#include "common.cuh"
#include "fastmath.cuh"

namespace nsagp {

struct McArgs {
  int D, N, M;
  long long T;
  int s;
  int sqrt_model;
  double link_shift;
  const double* Eft;       // [T][M]  (MATLAB M-by-T column-major)
  const double* Varft;     // [T][M]
  const double* W;         // [D][N] row-major
  const double* Z;         // [M][s][T] standard normal draws (MATLAB T-by-s-by-M), or nullptr: generate
  unsigned long long seed;
  double* Esig; double* Vsig;          // [T]
  double* Emod; double* Vmod;          // [T][N]  (MATLAB N-by-T)
};

__device__ __forceinline__ void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1,
                                              unsigned out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned long long p0 = (unsigned long long)0xD2511F53u * c0, p1 = (unsigned long long)0xCD9E8D57u * c2;
    const unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0, n1 = (unsigned)p1, n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1, n3 = (unsigned)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// two independent standard normals from one Philox block (Box-Muller on 53-bit / 32-bit uniforms)
__device__ __forceinline__ void normal_pair(unsigned long long seed, long long k, int v, int pair, double& z0, double& z1) {
  unsigned r[4];
  philox4x32_10((unsigned)k, (unsigned)((unsigned long long)k >> 32), (unsigned)v, (unsigned)pair, (unsigned)seed,
                (unsigned)(seed >> 32), r);
  const unsigned long long bits = ((unsigned long long)r[0] << 21) ^ (unsigned long long)(r[1] >> 11);      // 53 bits
  const double u = ((double)bits + 0.5) * (1.0 / 9007199254740992.0);                                        // (0, 1)
  const double ang = ((double)r[2] * 4294967296.0 + (double)r[3]) * (1.0 / 18446744073709551616.0);          // [0, 1)
  const double rad = sqrt(-2.0 * log(u));
  double sn, cs;
  sincospi(2.0 * ang, &sn, &cs);
  z0 = rad * cs; z1 = rad * sn;
}

constexpr int kMcMaxN = 8;

template <int DMAX>
__global__ void __launch_bounds__(128)
mc_reconstruct_kernel(const McArgs a) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  __shared__ double sW[DMAX * kMcMaxN];
  for (int i = threadIdx.x; i < a.D * a.N; i += blockDim.x) sW[i] = a.W[i];
  __syncthreads();
  if (k >= a.T) return;
  const int D = a.D, N = a.N, M = a.M, s = a.s;
  double ez[DMAX], sz[DMAX], eg[kMcMaxN], sg[kMcMaxN];
#pragma unroll
  for (int d = 0; d < DMAX; ++d)
    if (d < D) { ez[d] = a.Eft[k * M + d]; sz[d] = sqrt(a.Varft[k * M + d]); }
#pragma unroll
  for (int j = 0; j < kMcMaxN; ++j)
    if (j < N) { eg[j] = a.Eft[k * M + D + j]; sg[j] = sqrt(a.Varft[k * M + D + j]); }
  double mean_s = 0.0, m2_s = 0.0, mean_l[kMcMaxN], m2_l[kMcMaxN];
#pragma unroll
  for (int j = 0; j < kMcMaxN; ++j) { mean_l[j] = 0.0; m2_l[j] = 0.0; }
  const size_t TS = (size_t)a.T * s;
  for (int v = 0; v < s; ++v) {
    const double inv = 1.0 / (double)(v + 1);
    double lk[kMcMaxN];
    double zbuf = 0.0;
#pragma unroll
    for (int j = 0; j < kMcMaxN; ++j)
      if (j < N) {
        double z;
        if (a.Z) z = a.Z[(size_t)(D + j) * TS + (size_t)v * a.T + k];
        else { if (j & 1) z = zbuf; else normal_pair(a.seed, k, v, j >> 1, z, zbuf); }      // one Philox block per two latents
        const double g = fma(sg[j], z, eg[j]);
        lk[j] = log(1.0 + exp(g - a.link_shift));                       // link = @(g) log(1+exp(g-c))
        const double dl = lk[j] - mean_l[j];
        mean_l[j] = fma(dl, inv, mean_l[j]);
        m2_l[j] = fma(dl, lk[j] - mean_l[j], m2_l[j]);
      }
    double sig = 0.0;
#pragma unroll
    for (int d = 0; d < DMAX; ++d)
      if (d < D) {
        double z;
        if (a.Z) z = a.Z[(size_t)d * TS + (size_t)v * a.T + k];
        else { if (d & 1) z = zbuf; else normal_pair(a.seed, k, v, 0x10000 + (d >> 1), z, zbuf); }
        double wl = 0.0;
#pragma unroll
        for (int j = 0; j < kMcMaxN; ++j)
          if (j < N) wl = fma(sW[d * N + j], lk[j], wl);
        if (a.sqrt_model) wl = sqrt(wl);
        sig = fma(wl, fma(sz[d], z, ez[d]), sig);
      }
    const double ds = sig - mean_s;
    mean_s = fma(ds, inv, mean_s);
    m2_s = fma(ds, sig - mean_s, m2_s);
  }
  const double nm1 = s > 1 ? 1.0 / (double)(s - 1) : 0.0;               // var(x, 0, dim): N-1 normalisation (0 for s = 1)
  a.Esig[k] = mean_s; a.Vsig[k] = m2_s * nm1;
#pragma unroll
  for (int j = 0; j < kMcMaxN; ++j)
    if (j < N) { a.Emod[k * N + j] = mean_l[j]; a.Vmod[k * N + j] = m2_l[j] * nm1; }
}

}  // namespace nsagp
