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

// Walk the steps s0..s1-1 of one chunk with the element's inputs software-pipelined through registers: the raw inputs
// of step s + kPrefetch (sites, means: loads that do not depend on anything) and the table rows of step s + 1 (loads that
// depend on that step's sites) are requested BEFORE step s is processed.  Without this every step of a thread waits for
// its own HBM round trip (and a dependent L2 round trip for the table row) -- 32 of them in sequence, which is what kept
// these passes at 7-10 % of the HBM roof (profiles/r1bl_scan_full.md: long_scoreboard 29-39 stalls per issue).
// Elements without pipelined loads declare empty In / Tab and kPrefetch = 1.
template <class Elem, class Body>
__device__ __forceinline__ void scan_walk(Elem& el, const ScanArgs& a, long long s0, long long s1, Body&& body) {
  using In = typename Elem::In;
  using Tab = typename Elem::Tab;
  constexpr int PF = Elem::kPrefetch;
  In ring[PF];
  Tab tb;
#pragma unroll
  for (int j = 0; j < PF; ++j)
    if (s0 + j < s1) el.load(a.kfirst + a.dir * (s0 + j), ring[j]);
  el.lookup(a.kfirst + a.dir * s0, ring[0], tb);
  for (long long sb = s0; sb < s1; sb += PF) {
#pragma unroll
    for (int j = 0; j < PF; ++j) {
      const long long st = sb + j;
      if (st < s1) {
        const In cur = ring[j];
        const Tab tcur = tb;
        if (st + PF < s1) el.load(a.kfirst + a.dir * (st + PF), ring[j]);
        if (st + 1 < s1) el.lookup(a.kfirst + a.dir * (st + 1), ring[(j + 1) % PF], tb);
        body(a.kfirst + a.dir * st, cur, tcur);
      }
    }
  }
}

// Launch bounds of the two streaming phases: elements whose walk fits ~100 registers ask for two CTA tiles per SM
// (kScanThreads2 threads each); the host caps the tile at that many threads (api.cu: scan_setup).
constexpr int kScanThreads2 = 320;
template <class Elem> struct ScanBounds {
  static constexpr int kThreads = Elem::kTwoTiles ? kScanThreads2 : 256;     // 256 x 1: the full-state elements keep their 190-255 registers
  static constexpr int kMinBlocks = Elem::kTwoTiles ? 2 : 1;
};

template <class Elem>
__global__ void __launch_bounds__(ScanBounds<Elem>::kThreads, ScanBounds<Elem>::kMinBlocks)
scan_reduce_kernel(const DevProblem* __restrict__ probs, const DevState* __restrict__ states,
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
      const long long s0 = chunk * kScanSteps;
      const long long s1 = (s0 + kScanSteps < a.nsteps) ? s0 + kScanSteps : a.nsteps;
      bool first = true;
      scan_walk(el, a, s0, s1, [&](long long k, const typename Elem::In& in, const typename Elem::Tab& tb) {
        if (first) { el.get(k, in, tb, acc); first = false; }
        else { Map e; el.get(k, in, tb, e); Elem::compose(acc, e); }
      });
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
                  const double* __restrict__ tile_buf, double* __restrict__ tile_start, int batch, long long nmaps) {
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
  // nmaps > 0: walk that many maps stored at tile_buf (the segment aggregates of the two-level carry) instead
  const long long ntiles = nmaps > 0 ? nmaps : scan_num_tiles(a.nsteps, a.CH) + a.nprev;
  const size_t back = nmaps > 0 ? 0 : (size_t)a.nprev;
  const double* src = tile_buf + (size_t)blockIdx.x * ntiles * M * W - back * M * W;
  double* dst = tile_start + (size_t)blockIdx.x * ntiles * M * SW - back * M * SW;
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

// Two-level carry for long signals: the walk above is one application per CTA tile, i.e. nsteps / (32 CH) dependent
// steps on ONE SM (0.4 ms per pass at 10^6 steps, 20-60 % of a frozen pass).  The tiles are cut into segments of
// seg_len; thread (block n, segment g) composes its segment's tile aggregates (carry_seg_reduce), scan_carry_kernel
// walks the few segment aggregates, and thread (n, g) replays its segment from the state entering it, recording the
// state entering every tile (carry_seg_apply): 3 sqrt(ntiles) dependent steps instead of ntiles.
// tile_buf / tile_start point at the FIRST map of the list (including the nprev shard aggregates in front).
template <class Elem>
__global__ void carry_seg_reduce_kernel(const DevProblem* __restrict__ probs, const double* __restrict__ tile_buf,
                                        double* __restrict__ seg_buf, long long ntiles, int seg_len) {
  using Map = typename Elem::Map;
  constexpr int W = Elem::kMapDoubles;
  const int M = probs[blockIdx.y].M, n = threadIdx.x;
  const long long g = (long long)blockIdx.x * blockDim.y + threadIdx.y;
  const long long nseg = (ntiles + seg_len - 1) / seg_len;
  if (n >= M || g >= nseg) return;
  const long long t0 = g * seg_len;
  const long long t1 = t0 + seg_len < ntiles ? t0 + seg_len : ntiles;
  const double* src = tile_buf + (size_t)blockIdx.y * ntiles * M * W;
  Map acc, e;
  Elem::load_map(acc, src + ((size_t)t0 * M + n) * W);
  for (long long t = t0 + 1; t < t1; ++t) {
    Elem::load_map(e, src + ((size_t)t * M + n) * W);
    Elem::compose(acc, e);
  }
  Elem::store_map(acc, seg_buf + (((size_t)blockIdx.y * nseg + g) * M + n) * W);
}

template <class Elem>
__global__ void carry_seg_apply_kernel(const DevProblem* __restrict__ probs, const double* __restrict__ tile_buf,
                                       const double* __restrict__ seg_start, double* __restrict__ tile_start,
                                       long long ntiles, int seg_len) {
  using Map = typename Elem::Map;
  using State = typename Elem::State;
  constexpr int W = Elem::kMapDoubles, SW = Elem::kStateDoubles;
  const int M = probs[blockIdx.y].M, n = threadIdx.x;
  const long long g = (long long)blockIdx.x * blockDim.y + threadIdx.y;
  const long long nseg = (ntiles + seg_len - 1) / seg_len;
  if (n >= M || g >= nseg) return;
  const long long t0 = g * seg_len;
  const long long t1 = t0 + seg_len < ntiles ? t0 + seg_len : ntiles;
  const double* src = tile_buf + (size_t)blockIdx.y * ntiles * M * W;
  double* dst = tile_start + (size_t)blockIdx.y * ntiles * M * SW;
  State s;
  Elem::load_state(s, seg_start + (((size_t)blockIdx.y * nseg + g) * M + n) * SW);
  for (long long t = t0; t < t1; ++t) {
    Map e;
    Elem::load_map(e, src + ((size_t)t * M + n) * W);
    Elem::store_state(s, dst + ((size_t)t * M + n) * SW);
    Elem::apply(e, s);
  }
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
__global__ void __launch_bounds__(ScanBounds<Elem>::kThreads, ScanBounds<Elem>::kMinBlocks)
scan_apply_kernel(const DevProblem* __restrict__ probs, const DevState* __restrict__ states,
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
  el.begin_apply();
  State s;
  Elem::load_state(s, s_state + ((size_t)c * M + n) * SW);
  const long long s0 = chunk * kScanSteps;
  const long long s1 = (s0 + kScanSteps < a.nsteps) ? s0 + kScanSteps : a.nsteps;
  scan_walk(el, a, s0, s1, [&](long long k, const typename Elem::In& in, const typename Elem::Tab& tb) { el.step(k, in, tb, s); });
  el.finish_apply();
}

}  // namespace nsagp
