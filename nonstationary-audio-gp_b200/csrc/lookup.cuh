// Nearest-neighbour look-up of the steady-state tables by equivalent noise R.
//
// The reference does  [~,ind] = min(abs(r-R))  over the 200-point grid
// (ihgp_ep_modulator_nmf.m:239 filter, :379-380 smoother): nearest in *linear*
// distance, first index on ties, R = Inf -> index 1 in the filter (all distances
// Inf) but the last index in the smoother.  The same decision is reproduced
// exactly with precomputed thresholds: thr[i] is the smallest double R for which
// the argmin is > i (found on the host by bisection over bit patterns with the
// same floating-point expression), so ind = #{i : thr[i] <= R} -- a binary
// search instead of 200 subtractions.  For R >= 1e12, where rounding of r-R
// starts to produce ties between grid points, the search is done literally.
#pragma once
#include "common.cuh"

namespace nsagp {

__host__ __device__ inline int nearest_bruteforce(const double* r, int nr, double R) {
  int best = 0;
  double bd = fabs(r[0] - R);
  for (int i = 1; i < nr; ++i) {
    const double d = fabs(r[i] - R);
    if (d < bd) { bd = d; best = i; }   // strict: first index wins ties; NaN never wins
  }
  return best;
}

__device__ __forceinline__ int nearest_by_threshold(const double* thr, int nr, double R) {
  // number of thresholds <= R  (thr ascending, nr-1 entries)
  int lo = 0, hi = nr - 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (thr[mid] <= R) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// Filter rule (ihgp_ep_modulator_nmf.m:239).
__device__ __forceinline__ int lookup_filter(const double* r, const double* thr, int nr, double R) {
  if (!(R > 0.0)) return 0;             // R <= 0 or NaN: distances increase with r (or are all NaN)
  if (isinf(R)) return 0;               // all distances Inf -> first index
  if (R >= kLookupBig) return nearest_bruteforce(r, nr, R);
  return nearest_by_threshold(thr, nr, R);
}

// Smoother rule (ihgp_ep_modulator_nmf.m:379-380): Inf of either sign -> last row.
__device__ __forceinline__ int lookup_smoother(const double* r, const double* thr, int nr, double R) {
  if (isinf(R)) return nr - 1;
  if (!(R > 0.0)) return 0;
  if (R >= kLookupBig) return nearest_bruteforce(r, nr, R);
  return nearest_by_threshold(thr, nr, R);
}

}  // namespace nsagp
