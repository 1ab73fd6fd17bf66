// Three-phase chunked scan over time for the passes that are linear once the EP sites
// are frozen (every IHGP filter pass >= 2, every RTS smoother pass).
//
// An element type `Elem` describes one block's recursion:
//   Map   : the associative element of one step (affine map of the mean; (E, g, L) of the
//           RTS smoother, Sarkka & Garcia-Fernandez 2021), with compose / apply;
//   State : what is carried (mean, or mean + covariance);
//   step  : the reference's LITERAL step applied to a State, emitting the outputs.
// Steps are numbered 0..nsteps-1 in PROCESSING order; step s is time k = kfirst + dir*s.
//
//   phase 1  scan_reduce : thread (block n, chunk c) composes the kSteps maps of its chunk; the
//            CTA's chunks are then composed in order by warp 0 (from shared memory) into one
//            CTA aggregate.  Chunk aggregates and CTA aggregates go to HBM.
//   phase 2  scan_carry  : one warp per signal walks the CTA aggregates (prefetched ahead),
//            recording the state entering every CTA tile.  This is the only sequential part:
//            nsteps / (kSteps * CH) applications -- and the only part that would cross GPUs
//            when a signal is time-chunked over ranks (the carry exchange).
//   phase 3  scan_apply  : warp 0 turns the entering state + chunk aggregates into the state
//            entering every chunk (shared memory); every thread then re-applies the literal
//            steps of its chunk and emits.  Only the chunk-entry states are re-associated.
//
// Thread layout: threadIdx.x = latent block n (coalesced M-contiguous site loads; blockDim.x = M, so no idle lanes),
// threadIdx.y = chunk within the CTA tile (CH chunks), grid = (tiles, signals).
#pragma once
#include "common.cuh"

namespace nsagp {

constexpr int kScanSteps = 32;     // time steps composed by one thread

struct ScanArgs {
  long long kfirst, nsteps;
  int dir;                         // +1 forward in time, -1 backward
  int init;                        // Elem-specific initial-state selector
  long long kinit;
  int CH;                          // chunks per CTA tile (blockDim.y)
  int flags;                       // Elem-specific (bit 0: nlZ-mode rules of the full-state filter)
  int nprev;                       // carry only: aggregates of preceding shards (other GPUs) placed before the tiles
};

__host__ __device__ inline long long scan_num_chunks(long long nsteps) { return (nsteps + kScanSteps - 1) / kScanSteps; }
__host__ __device__ inline long long scan_num_tiles(long long nsteps, int CH) {
  const long long per = (long long)kScanSteps * CH;
  return (nsteps + per - 1) / per;
}

template <class Elem>
__global__ void scan_reduce_kernel(const DevProblem* __restrict__ probs, const DevState* __restrict__ states,
                                   ScanArgs a, double* __restrict__ chunk_buf, double* __restrict__ tile_buf) {
  using Map = typename Elem::Map;
  constexpr int W = Elem::kMapDoubles;
  const DevProblem& P = probs[blockIdx.y];
  const DevState& St = states[blockIdx.y];
  const int n = threadIdx.x, c = threadIdx.y, M = P.M, CH = a.CH;
  extern __shared__ double sm[];                 // [CH][M][W]
  const long long nchunks = scan_num_chunks(a.nsteps);
  const long long ntiles = scan_num_tiles(a.nsteps, CH);
  const long long chunk = (long long)blockIdx.x * CH + c;
  const bool live = n < M && chunk < nchunks;
  if (n < M) {
    Map acc;
    if (live) {
      Elem el(P, St, n, a);
      Map e;
      const long long s0 = chunk * kScanSteps;
      const long long s1 = (s0 + kScanSteps < a.nsteps) ? s0 + kScanSteps : a.nsteps;
      if (s0 + 1 < a.nsteps) el.prefetch(a.kfirst + a.dir * (s0 + 1));
      el.get(a.kfirst + a.dir * s0, acc);
      for (long long s = s0 + 1; s < s1; ++s) {
        if (s + 2 < a.nsteps) el.prefetch(a.kfirst + a.dir * (s + 2));
        el.get(a.kfirst + a.dir * s, e);
        Elem::compose(acc, e);
      }
      el.finish_reduce();
      Elem::store_map(acc, chunk_buf + (((size_t)blockIdx.y * nchunks + chunk) * M + n) * W);
      Elem::store_map(acc, sm + ((size_t)c * M + n) * W);
    }
  }
  __syncthreads();
  if (c == 0 && n < M) {
    // compose this tile's chunks in processing order
    const long long first = (long long)blockIdx.x * CH;
    const int cnt = (int)((nchunks - first < CH) ? nchunks - first : CH);
    Map acc, e;
    Elem::load_map(acc, sm + (size_t)n * W);
    for (int j = 1; j < cnt; ++j) {
      Elem::load_map(e, sm + ((size_t)j * M + n) * W);
      Elem::compose(acc, e);
    }
    Elem::store_map(acc, tile_buf + (((size_t)blockIdx.y * ntiles + blockIdx.x) * M + n) * W);
  }
}

// One CTA per signal.  tile_start[(b, tile, n)] = state entering the tile.  The walk over the
// tile aggregates is the sequential part of the scan, so its loads must not sit on the chain:
// all threads stage a batch of aggregates in shared memory (coalesced), then warp 0 walks it.
constexpr int kCarryThreads = 256;

template <class Elem>
__global__ void __launch_bounds__(kCarryThreads)
scan_carry_kernel(const DevProblem* __restrict__ probs, const DevState* __restrict__ states, ScanArgs a,
                  const double* __restrict__ tile_buf, double* __restrict__ tile_start, int batch) {
  using Map = typename Elem::Map;
  using State = typename Elem::State;
  constexpr int W = Elem::kMapDoubles, SW = Elem::kStateDoubles;
  const DevProblem& P = probs[blockIdx.x];
  const DevState& St = states[blockIdx.x];
  const int tid = threadIdx.x, M = P.M;
  extern __shared__ double sm[];                 // [batch][M][W]
  // When a signal is time-chunked over GPUs, the aggregates of the shards that come earlier in
  // processing order sit in the nprev slots before tile_buf (single-problem plans only): the
  // walk starts from the global initial state and passes through them first.
  const long long ntiles = scan_num_tiles(a.nsteps, a.CH) + a.nprev;
  const double* src = tile_buf + (size_t)blockIdx.x * ntiles * M * W - (size_t)a.nprev * M * W;
  double* dst = tile_start + (size_t)blockIdx.x * ntiles * M * SW - (size_t)a.nprev * M * SW;
  const bool walker = tid < M;
  Elem el(P, St, walker ? tid : 0, a);
  State s;
  if (walker) el.init(s, a.init, a.kinit);
  for (long long t0 = 0; t0 < ntiles; t0 += batch) {
    const int cnt = (int)((ntiles - t0 < batch) ? ntiles - t0 : batch);
    const size_t words = (size_t)cnt * M * W;
    for (size_t i = tid; i < words; i += kCarryThreads) sm[i] = src[(size_t)t0 * M * W + i];
    __syncthreads();
    if (walker) {
      for (int j = 0; j < cnt; ++j) {
        Map e;
        Elem::load_map(e, sm + ((size_t)j * M + tid) * W);
        Elem::store_state(s, dst + ((size_t)(t0 + j) * M + tid) * SW);
        Elem::apply(e, s);
      }
    }
    __syncthreads();
  }
  if (walker) el.store_final(s);   // the state after the last step (used by a following tile / rank)
}

// The shard's single aggregate (composition of its tiles in processing order): what a GPU
// contributes to the carry exchange when a signal is time-chunked over ranks.
template <class Elem>
__global__ void __launch_bounds__(32)
scan_total_kernel(const DevProblem* __restrict__ probs, ScanArgs a, const double* __restrict__ tile_buf,
                  double* __restrict__ total) {
  using Map = typename Elem::Map;
  constexpr int W = Elem::kMapDoubles;
  const int n = threadIdx.x, M = probs[0].M;
  if (n >= M) return;
  const long long ntiles = scan_num_tiles(a.nsteps, a.CH);
  Map acc, e;
  Elem::load_map(acc, tile_buf + (size_t)n * W);
  for (long long t = 1; t < ntiles; ++t) {
    Elem::load_map(e, tile_buf + ((size_t)t * M + n) * W);
    Elem::compose(acc, e);
  }
  Elem::store_map(acc, total + (size_t)n * W);
}

template <class Elem>
__global__ void scan_apply_kernel(const DevProblem* __restrict__ probs, const DevState* __restrict__ states,
                                  ScanArgs a, const double* __restrict__ chunk_buf,
                                  const double* __restrict__ tile_start) {
  using Map = typename Elem::Map;
  using State = typename Elem::State;
  constexpr int W = Elem::kMapDoubles, SW = Elem::kStateDoubles;
  const DevProblem& P = probs[blockIdx.y];
  const DevState& St = states[blockIdx.y];
  const int n = threadIdx.x, c = threadIdx.y, M = P.M, CH = a.CH;
  extern __shared__ double sm[];                 // [CH][M][SW] states entering each chunk | [CH][M][W] chunk maps
  const long long nchunks = scan_num_chunks(a.nsteps);
  const long long ntiles = scan_num_tiles(a.nsteps, CH);
  const long long first = (long long)blockIdx.x * CH;
  const int cnt = (int)((nchunks - first < CH) ? nchunks - first : CH);
  double* s_state = sm;
  double* s_map = sm + (size_t)CH * M * SW;
  {
    // stage this tile's chunk aggregates (contiguous in HBM) so the walk below runs from shared memory
    const double* src = chunk_buf + ((size_t)blockIdx.y * nchunks + first) * M * W;
    const int nthreads = blockDim.x * blockDim.y, tid = c * blockDim.x + n;
    for (int i = tid; i < cnt * M * W; i += nthreads) s_map[i] = src[i];
  }
  __syncthreads();
  if (c == 0 && n < M) {
    State s;
    Elem::load_state(s, tile_start + (((size_t)blockIdx.y * ntiles + blockIdx.x) * M + n) * SW);
    for (int j = 0; j < cnt; ++j) {
      Map e;
      Elem::load_map(e, s_map + ((size_t)j * M + n) * W);
      Elem::store_state(s, s_state + ((size_t)j * M + n) * SW);
      Elem::apply(e, s);
    }
  }
  __syncthreads();
  const long long chunk = first + c;
  if (n >= M || chunk >= nchunks) return;
  Elem el(P, St, n, a);
  State s;
  Elem::load_state(s, s_state + ((size_t)c * M + n) * SW);
  const long long s0 = chunk * kScanSteps;
  const long long s1 = (s0 + kScanSteps < a.nsteps) ? s0 + kScanSteps : a.nsteps;
  if (s0 + 1 < a.nsteps) el.prefetch(a.kfirst + a.dir * (s0 + 1));
  for (long long t = s0; t < s1; ++t) {
    if (t + 2 < a.nsteps) el.prefetch(a.kfirst + a.dir * (t + 2));
    el.step(a.kfirst + a.dir * t, s);
  }
  el.finish_apply();
}

}  // namespace nsagp
