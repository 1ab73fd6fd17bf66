// Common definitions for the nsagp sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace nsagp {

constexpr int kMaxBlock = 8;     // largest per-latent state block (matern72 subband = 8)
constexpr int kMaxSites = 64;    // D + N
constexpr int kWarp = 32;
constexpr double kInvSqrt2Pi = 0.39894228040143267793994605993438;
constexpr double kJitter = 1e-10;          // likModulatorNMFPower.m:28
constexpr int kCthrPad = 8;                // sentinels on either side of the ttau thresholds (window look-up)
constexpr int kMaxDist = 8;                // distinct sigma-point coordinates per modulator the link table holds
constexpr double kLookupBig = 1e12;        // above this the nearest-neighbour search is done by brute force

// Per-problem constant data resident in HBM (built once by the plan).
// All per-block arrays are padded to BM = max(bz,bg) so one code path serves
// both block families; padding is exact zeros.
struct DevProblem {
  int D, N, M, BM, n;            // sites, padded block size, true state dim
  int bz, bg;
  int S;                         // sigma points
  int nr;                        // table rows (IHGP); 0 for the full-state path
  int lik_kind;                  // 0 power, 1 precalc(sqrt)
  double sn2, link_shift;
  const int* off;                // [M+1] start of each block in the n-vector
  const double* A;               // [M][BM*BM] column-major, zero padded
  const double* Q;               // [M][BM*BM]
  const double* Pinf;            // [M][BM*BM]
  const double* h;               // [M][BM]
  const double* hA;              // [M][BM]   h*A
  const double* W;               // [DP][NP] row-major padded (DP = 16/32, NP = 4)
  const double* wn;              // [S]
  const double* xn;              // [NP][S]
  // Distinct abscissae per modulator: the coordinates of a fully symmetric or tensor rule take few distinct values
  // (0, +-u, +-v for utp_ws(9, N)), so the link function is evaluated ndist times per modulator and step instead of S
  // times (same arithmetic on the same inputs: bit-identical).  ndist = 0: not available (more than kMaxDist values).
  int ndist;
  const double* xdist;           // [NP][ndist]
  const unsigned char* xidx;     // [NP][S] index into xdist
  // IHGP tables
  const double* r;               // [nr]
  const double* thr;             // [nr-1] decision thresholds of the nearest-neighbour search; kCthrPad sentinels
                                 // (-Inf below index 0, +Inf above nr-2) surround the array
  double rg_a, rg_b;             // row guess from the high word of R: floor(rg_a + rg_b * hi32(R)) (log-spaced grid)
  const double* Wtab;            // [M][nr+1][BM]  PP*h'  (row nr: Pinf)
  const double* HPHtab;          // [M][nr+1]      h*PP*h'
  const double* Gtab;            // [M][nr][BM*BM] smoother gain
  const double* vmtab;           // [M][nr]        h*Ps*h'
  const double* cthr;            // [nr-1] the same thresholds expressed on ttau: 1/ttau >= thr[i]  <=>  ttau <= cthr[i];
                                 // kCthrPad sentinels (+Inf below index 0, -Inf above nr-2) surround the array
  const double* SDtab;           // [N][nr+1]      sqrt(h*PP*h') of the modulator blocks
  const double* RS2tab;          // [N][nr+1]      1/(h*PP*h')
};

// Per-problem mutable state in HBM.  Site arrays are time-major, M contiguous
// doubles per step: exactly MATLAB's M-by-T column-major layout.
struct DevState {
  const double* y;               // [T]
  double* ttau; double* tnu; double* R;   // [T][M]
  double* MS;                    // [T][n]   filtered, then smoothed means
  double* E;                     // [T][M]   h*m of the last smoother pass (Eft)
  double* V;                     // [T][M]   h*P*h' (full-state path: Varft)
  double* lZ;                    // [T]
  double* PS;                    // [T][M][BM*BM] full-state path covariances (filtered, then smoothed)
  double* mcarry;                // [M][BM] initial mean of the next filter pass
  double* vm0;                   // [M] marginal variance at k=0 of the last smoother pass (IHGP Varft)
  unsigned long long* maxdiff;   // [2] bit patterns of non-negative doubles (maxDiffM, maxDiffP)
  unsigned long long* negcav;    // [1]
  int* status;                   // [1] sticky error flag (first failure wins)
};

__device__ __forceinline__ void atomic_max_nonneg(unsigned long long* addr, double v) {
  // valid for v >= 0 (bit patterns of non-negative doubles are ordered); NaN is ignored
  if (v >= 0.0) atomicMax(addr, (unsigned long long)__double_as_longlong(v));
}

}  // namespace nsagp
