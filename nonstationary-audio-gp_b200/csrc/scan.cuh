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

// ScanArgs::flags bit 8: at the start of a CTA tile one thread asks the memory system to bring the tile's input rows --
// contiguous in HBM, CH * 32 steps per array -- into L2 with bulk prefetches (cp.async.bulk.prefetch.L2).  The walks
// below are dependent chains of 32 steps per thread whose loads are issued a few steps ahead; with the rows already on
// their way to L2 a step waits for an L2 hit instead of an HBM round trip.  Elements opt in by defining prefetch_rows.
constexpr int kScanFlagL2Prefetch = 256;

__device__ __forceinline__ void l2_prefetch_bulk(const void* p, size_t bytes) {
  const unsigned long long a0 = (unsigned long long)p & ~15ull;
  const unsigned long long a1 = ((unsigned long long)p + bytes + 15ull) & ~15ull;
  for (unsigned long long a = a0; a < a1; a += 1u << 20) {             // (any size works; keep a request below 1 MiB)
    const unsigned n = (unsigned)((a1 - a < (1u << 20)) ? a1 - a : (1u << 20));
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"(n) : "memory");
  }
}

__device__ __forceinline__ void scan_tile_range(const ScanArgs& a, long long tile, long long& k_lo, long long& k_hi) {
  const long long s0 = tile * a.CH * kScanSteps;
  const long long s1 = (s0 + (long long)a.CH * kScanSteps < a.nsteps) ? s0 + (long long)a.CH * kScanSteps : a.nsteps;
  if (a.dir > 0) { k_lo = a.kfirst + s0; k_hi = a.kfirst + s1; }
  else { k_lo = a.kfirst - s1 + 1; k_hi = a.kfirst - s0 + 1; }
}

template <class Elem>
__device__ __forceinline__ void scan_tile_prefetch(const DevProblem& P, const DevState& St, const ScanArgs& a, bool apply) {
  if ((a.flags & kScanFlagL2Prefetch) && threadIdx.x == 0) {
    long long k_lo, k_hi;
    scan_tile_range(a, blockIdx.x, k_lo, k_hi);
    if (k_hi > k_lo) Elem::prefetch_rows(P, St, k_lo, k_hi, apply);
  }
}

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

// Two latent families: subbands (n < D, block size bz) and modulators (n >= D, block size bg).  Their element types
// share the slot sizes kMapDoubles / kStateDoubles (padded to max(bz, bg)), so the buffers have ONE layout; the two
// streaming phases are launched once per family (latents n0 .. n0+cnt-1) with that family's element type -- its own
// arithmetic size and its own register allocation (2 x 2 subband blocks next to 3 x 3 modulators: 100 instead of 190
// registers for 84 % of the threads).  The carry kernels handle both families in one launch.
struct ScanThread {
  int n, c;
  bool live;
  __device__ __forceinline__ ScanThread(int tid, int n0, int cnt, int CH) {
    c = tid / cnt;
    n = n0 + (tid - c * cnt);
    live = tid < CH * cnt;
  }
};

template <class Elem>
__device__ __forceinline__ void scan_reduce_thread(const DevProblem& P, const DevState& St, const ScanArgs& a, int n, int c,
                                                   long long chunk, double* chunk_slot, double* sm_slot) {
  using Map = typename Elem::Map;
  Map acc;
  Elem el(P, St, n, a);
  const long long s0 = chunk * kScanSteps;
  const long long s1 = (s0 + kScanSteps < a.nsteps) ? s0 + kScanSteps : a.nsteps;
  bool first = true;
  scan_walk(el, a, s0, s1, [&](long long k, const typename Elem::In& in, const typename Elem::Tab& tb) {
    if (first) { el.get(k, in, tb, acc); first = false; }
    else { Map e; el.get(k, in, tb, e); Elem::compose(acc, e); }
  });
  el.finish_reduce();
  Elem::store_map(acc, chunk_slot);
  Elem::store_map(acc, sm_slot);
}

template <class Elem>
__device__ __forceinline__ void scan_tile_compose_thread(const double* sm, int n, int M, int cnt, double* tile_slot) {
  using Map = typename Elem::Map;
  constexpr int W = Elem::kMapDoubles;
  Map acc, e;
  Elem::load_map(acc, sm + (size_t)n * W);
  for (int j = 1; j < cnt; ++j) {
    Elem::load_map(e, sm + ((size_t)j * M + n) * W);
    Elem::compose(acc, e);
  }
  Elem::store_map(acc, tile_slot);
}

template <class Elem>
__global__ void __launch_bounds__(ScanBounds<Elem>::kThreads, ScanBounds<Elem>::kMinBlocks)
scan_reduce_kernel(const DevProblem* __restrict__ probs, const DevState* __restrict__ states,
                   ScanArgs a, double* __restrict__ chunk_buf, double* __restrict__ tile_buf, int n0, int cnt_lat) {
  constexpr int W = Elem::kMapDoubles;
  const DevProblem& P = probs[blockIdx.y];
  const DevState& St = states[blockIdx.y];
  const int M = P.M, CH = a.CH;
  const ScanThread th(threadIdx.x, n0, cnt_lat, CH);
  const int n = th.n, c = th.c;
  scan_tile_prefetch<Elem>(P, St, a, false);
  extern __shared__ double sm[];                 // [CH][cnt_lat][W]: this family's latents only
  const long long nchunks = scan_num_chunks(a.nsteps);
  const long long ntiles = scan_num_tiles(a.nsteps, CH);
  const long long chunk = (long long)blockIdx.x * CH + c;
  if (th.live && chunk < nchunks)
    scan_reduce_thread<Elem>(P, St, a, n, c, chunk, chunk_buf + (((size_t)blockIdx.y * nchunks + chunk) * M + n) * W,
                             sm + ((size_t)c * cnt_lat + (n - n0)) * W);
  __syncthreads();
  if (th.live && c == 0) {
    // compose this tile's chunks in processing order
    const long long first = (long long)blockIdx.x * CH;
    const int cnt = (int)((nchunks - first < CH) ? nchunks - first : CH);
    scan_tile_compose_thread<Elem>(sm, n - n0, cnt_lat, cnt, tile_buf + (((size_t)blockIdx.y * ntiles + blockIdx.x) * M + n) * W);
  }
}

// Both latent families in ONE CTA tile, each with its own element type: threads [0, CH * D) (rounded up to whole warps)
// are the subband family, the threads after them the modulators, so no warp mixes the two walks.  Against one launch per
// family (above) the tile's rows of the site arrays -- M contiguous doubles per step, of which a family reads its share --
// are fetched from HBM once instead of once per family (the second family's sectors hit L1 / L2), and the small
// modulator-family CTAs (CH * N threads) no longer run on their own.  Shared memory: [CH][D][W] | [CH][N][W].
struct ScanThread2 {
  int n, c, n0, cnt, ftid, fthreads;
  bool live, fam_g;
  size_t sm_off;                   // offset of the family's region, in units of (its slot size) doubles per (chunk, latent)
  __device__ __forceinline__ ScanThread2(int tid, int D, int N, int CH) {
    const int zthreads = (CH * D + 31) & ~31;
    fam_g = tid >= zthreads;
    ftid = fam_g ? tid - zthreads : tid;
    n0 = fam_g ? D : 0;
    cnt = fam_g ? N : D;
    fthreads = fam_g ? (int)blockDim.x - zthreads : zthreads;
    c = ftid / cnt;
    n = n0 + (ftid - c * cnt);
    live = ftid < CH * cnt;
    sm_off = fam_g ? (size_t)CH * D : 0;
  }
};
__host__ __device__ inline int scan2_threads(int D, int N, int CH) { return ((CH * D + 31) & ~31) + CH * N; }

// TH: thread bound of the CTA tile.  The mean elements ask for two tiles per SM; with TH = 320 (16 chunks: 304 threads) that
// caps a thread at 96 registers and the walks spill 250-750 bytes, with TH = 256 (12 chunks: 228 threads) at 128 registers
// and the spills all but vanish -- fewer warps per SM against less local-memory traffic through an L1 that is the
// busiest unit of these kernels (profiles/r2n_scan.md).  Measured (profiles/r2q_ab.jsonl): 256 threads are 6-26 % faster at
// every length; the host picks them by default (nsagp_scan_tile).
template <class EZ, class EG, int TH = ScanBounds<EZ>::kThreads>
__global__ void __launch_bounds__(TH, ScanBounds<EZ>::kMinBlocks)
scan_reduce2_kernel(const DevProblem* __restrict__ probs, const DevState* __restrict__ states,
                    ScanArgs a, double* __restrict__ chunk_buf, double* __restrict__ tile_buf) {
  constexpr int W = EZ::kMapDoubles;
  const DevProblem& P = probs[blockIdx.y];
  const DevState& St = states[blockIdx.y];
  const int M = P.M, CH = a.CH;
  const ScanThread2 th(threadIdx.x, P.D, P.N, CH);
  scan_tile_prefetch<EZ>(P, St, a, false);
  extern __shared__ double sm[];
  double* smf = sm + th.sm_off * W;
  const long long nchunks = scan_num_chunks(a.nsteps);
  const long long ntiles = scan_num_tiles(a.nsteps, CH);
  const long long chunk = (long long)blockIdx.x * CH + th.c;
  if (th.live && chunk < nchunks) {
    double* cslot = chunk_buf + (((size_t)blockIdx.y * nchunks + chunk) * M + th.n) * W;
    double* sslot = smf + ((size_t)th.c * th.cnt + (th.n - th.n0)) * W;
    if (th.fam_g) scan_reduce_thread<EG>(P, St, a, th.n, th.c, chunk, cslot, sslot);
    else scan_reduce_thread<EZ>(P, St, a, th.n, th.c, chunk, cslot, sslot);
  }
  __syncthreads();
  if (th.live && th.c == 0) {
    const long long first = (long long)blockIdx.x * CH;
    const int cnt = (int)((nchunks - first < CH) ? nchunks - first : CH);
    double* tslot = tile_buf + (((size_t)blockIdx.y * ntiles + blockIdx.x) * M + th.n) * W;
    if (th.fam_g) scan_tile_compose_thread<EG>(smf, th.n - th.n0, th.cnt, cnt, tslot);
    else scan_tile_compose_thread<EZ>(smf, th.n - th.n0, th.cnt, cnt, tslot);
  }
}

// One CTA per signal.  tile_start[(b, tile, n)] = state entering the tile.  The walk over the
// tile aggregates is the sequential part of the scan, so its loads must not sit on the chain:
// all threads stage a batch of aggregates in shared memory (coalesced), then warp 0 walks it.
constexpr int kCarryThreads = 256;

template <class Elem>
__device__ __forceinline__ void scan_carry_walk(Elem& el, typename Elem::State& s, const double* sm, double* dst, int M, int tid,
                                                long long t0, int cnt) {
  using Map = typename Elem::Map;
  constexpr int W = Elem::kMapDoubles, SW = Elem::kStateDoubles;
  for (int j = 0; j < cnt; ++j) {
    Map e;
    Elem::load_map(e, sm + ((size_t)j * M + tid) * W);
    Elem::store_state(s, dst + ((size_t)(t0 + j) * M + tid) * SW);
    Elem::apply(e, s);
  }
}

template <class EZ, class EG>
__global__ void __launch_bounds__(kCarryThreads)
scan_carry_kernel(const DevProblem* __restrict__ probs, const DevState* __restrict__ states, ScanArgs a,
                  const double* __restrict__ tile_buf, double* __restrict__ tile_start, int batch, long long nmaps) {
  constexpr int W = EZ::kMapDoubles, SW = EZ::kStateDoubles;
  const DevProblem& P = probs[blockIdx.x];
  const DevState& St = states[blockIdx.x];
  const int tid = threadIdx.x, M = P.M, D = P.D;
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
  const bool fz = tid < D;
  EZ elz(P, St, (walker && fz) ? tid : 0, a);
  EG elg(P, St, (walker && !fz) ? tid : D, a);
  typename EZ::State sz;
  typename EG::State sg;
  if (walker) { if (fz) elz.init(sz, a.init, a.kinit); else elg.init(sg, a.init, a.kinit); }
  for (long long t0 = 0; t0 < ntiles; t0 += batch) {
    const int cnt = (int)((ntiles - t0 < batch) ? ntiles - t0 : batch);
    const size_t words = (size_t)cnt * M * W;
    for (size_t i = tid; i < words; i += kCarryThreads) sm[i] = src[(size_t)t0 * M * W + i];
    __syncthreads();
    if (walker) {
      if (fz) scan_carry_walk<EZ>(elz, sz, sm, dst, M, tid, t0, cnt);
      else scan_carry_walk<EG>(elg, sg, sm, dst, M, tid, t0, cnt);
    }
    __syncthreads();
  }
  if (walker) { if (fz) elz.store_final(sz); else elg.store_final(sg); }   // the state after the last step
}

// Two-level carry for long signals: the walk above is one application per CTA tile, i.e. nsteps / (32 CH) dependent
// steps on ONE SM (0.4 ms per pass at 10^6 steps, 20-60 % of a frozen pass).  The tiles are cut into segments of
// seg_len; thread (block n, segment g) composes its segment's tile aggregates (carry_seg_reduce), scan_carry_kernel
// walks the few segment aggregates, and thread (n, g) replays its segment from the state entering it, recording the
// state entering every tile (carry_seg_apply): 3 sqrt(ntiles) dependent steps instead of ntiles.
// tile_buf / tile_start point at the FIRST map of the list (including the nprev shard aggregates in front).
template <class Elem>
__device__ __forceinline__ void carry_seg_reduce_thread(const double* src, double* out, int M, int n, long long t0, long long t1) {
  using Map = typename Elem::Map;
  constexpr int W = Elem::kMapDoubles;
  Map acc, e;
  Elem::load_map(acc, src + ((size_t)t0 * M + n) * W);
  for (long long t = t0 + 1; t < t1; ++t) {
    Elem::load_map(e, src + ((size_t)t * M + n) * W);
    Elem::compose(acc, e);
  }
  Elem::store_map(acc, out);
}

template <class EZ, class EG>
__global__ void carry_seg_reduce_kernel(const DevProblem* __restrict__ probs, const double* __restrict__ tile_buf,
                                        double* __restrict__ seg_buf, long long ntiles, int seg_len) {
  constexpr int W = EZ::kMapDoubles;
  const int M = probs[blockIdx.y].M, D = probs[blockIdx.y].D, n = threadIdx.x;
  const long long g = (long long)blockIdx.x * blockDim.y + threadIdx.y;
  const long long nseg = (ntiles + seg_len - 1) / seg_len;
  if (n >= M || g >= nseg) return;
  const long long t0 = g * seg_len;
  const long long t1 = t0 + seg_len < ntiles ? t0 + seg_len : ntiles;
  const double* src = tile_buf + (size_t)blockIdx.y * ntiles * M * W;
  double* out = seg_buf + (((size_t)blockIdx.y * nseg + g) * M + n) * W;
  if (n < D) carry_seg_reduce_thread<EZ>(src, out, M, n, t0, t1);
  else carry_seg_reduce_thread<EG>(src, out, M, n, t0, t1);
}

template <class Elem>
__device__ __forceinline__ void carry_seg_apply_thread(const double* src, const double* start, double* dst, int M, int n,
                                                       long long t0, long long t1) {
  using Map = typename Elem::Map;
  using State = typename Elem::State;
  constexpr int W = Elem::kMapDoubles, SW = Elem::kStateDoubles;
  State s;
  Elem::load_state(s, start);
  for (long long t = t0; t < t1; ++t) {
    Map e;
    Elem::load_map(e, src + ((size_t)t * M + n) * W);
    Elem::store_state(s, dst + ((size_t)t * M + n) * SW);
    Elem::apply(e, s);
  }
}

template <class EZ, class EG>
__global__ void carry_seg_apply_kernel(const DevProblem* __restrict__ probs, const double* __restrict__ tile_buf,
                                       const double* __restrict__ seg_start, double* __restrict__ tile_start,
                                       long long ntiles, int seg_len) {
  constexpr int W = EZ::kMapDoubles, SW = EZ::kStateDoubles;
  const int M = probs[blockIdx.y].M, D = probs[blockIdx.y].D, n = threadIdx.x;
  const long long g = (long long)blockIdx.x * blockDim.y + threadIdx.y;
  const long long nseg = (ntiles + seg_len - 1) / seg_len;
  if (n >= M || g >= nseg) return;
  const long long t0 = g * seg_len;
  const long long t1 = t0 + seg_len < ntiles ? t0 + seg_len : ntiles;
  const double* src = tile_buf + (size_t)blockIdx.y * ntiles * M * W;
  double* dst = tile_start + (size_t)blockIdx.y * ntiles * M * SW;
  const double* start = seg_start + (((size_t)blockIdx.y * nseg + g) * M + n) * SW;
  if (n < D) carry_seg_apply_thread<EZ>(src, start, dst, M, n, t0, t1);
  else carry_seg_apply_thread<EG>(src, start, dst, M, n, t0, t1);
}

// The shard's single aggregate (composition of its tiles in processing order): what a GPU
// contributes to the carry exchange when a signal is time-chunked over ranks.
template <class EZ, class EG>
__global__ void __launch_bounds__(32)
scan_total_kernel(const DevProblem* __restrict__ probs, const double* __restrict__ maps, long long nmaps,
                  double* __restrict__ total) {
  constexpr int W = EZ::kMapDoubles;
  const int n = threadIdx.x, M = probs[0].M, D = probs[0].D;
  if (n >= M) return;
  if (n < D) carry_seg_reduce_thread<EZ>(maps, total + (size_t)n * W, M, n, 0, nmaps);
  else carry_seg_reduce_thread<EG>(maps, total + (size_t)n * W, M, n, 0, nmaps);
}

template <class Elem>
__device__ __forceinline__ void scan_apply_entry_thread(const double* tile_slot, const double* s_map, double* s_state, int n, int M,
                                                        int cnt) {
  using Map = typename Elem::Map;
  using State = typename Elem::State;
  constexpr int W = Elem::kMapDoubles, SW = Elem::kStateDoubles;
  State s;
  Elem::load_state(s, tile_slot);
  for (int j = 0; j < cnt; ++j) {
    Map e;
    Elem::load_map(e, s_map + ((size_t)j * M + n) * W);
    Elem::store_state(s, s_state + ((size_t)j * M + n) * SW);
    Elem::apply(e, s);
  }
}

template <class Elem>
__device__ __forceinline__ void scan_apply_thread(const DevProblem& P, const DevState& St, const ScanArgs& a, int n, long long chunk,
                                                  const double* state_slot) {
  using State = typename Elem::State;
  Elem el(P, St, n, a);
  el.begin_apply();
  State s;
  Elem::load_state(s, state_slot);
  const long long s0 = chunk * kScanSteps;
  const long long s1 = (s0 + kScanSteps < a.nsteps) ? s0 + kScanSteps : a.nsteps;
  scan_walk(el, a, s0, s1, [&](long long k, const typename Elem::In& in, const typename Elem::Tab& tb) { el.step(k, in, tb, s); });
  el.finish_apply();
}

template <class Elem>
__global__ void __launch_bounds__(ScanBounds<Elem>::kThreads, ScanBounds<Elem>::kMinBlocks)
scan_apply_kernel(const DevProblem* __restrict__ probs, const DevState* __restrict__ states,
                  ScanArgs a, const double* __restrict__ chunk_buf,
                  const double* __restrict__ tile_start, int n0, int cnt_lat) {
  constexpr int W = Elem::kMapDoubles, SW = Elem::kStateDoubles;
  const DevProblem& P = probs[blockIdx.y];
  const DevState& St = states[blockIdx.y];
  const int M = P.M, CH = a.CH;
  const ScanThread th(threadIdx.x, n0, cnt_lat, CH);
  const int n = th.n, c = th.c;
  scan_tile_prefetch<Elem>(P, St, a, true);
  extern __shared__ double sm[];                 // [CH][cnt_lat][SW] states entering each chunk | [CH][cnt_lat][W] chunk maps
  const long long nchunks = scan_num_chunks(a.nsteps);
  const long long ntiles = scan_num_tiles(a.nsteps, CH);
  const long long first = (long long)blockIdx.x * CH;
  const int cnt = (int)((nchunks - first < CH) ? nchunks - first : CH);
  double* s_state = sm;
  double* s_map = sm + (size_t)CH * cnt_lat * SW;
  {
    // stage this family's share of the tile's chunk aggregates (cnt_lat * W contiguous doubles per chunk)
    const double* src = chunk_buf + (((size_t)blockIdx.y * nchunks + first) * M + n0) * W;
    const int row = cnt_lat * W;
    for (int i = threadIdx.x; i < cnt * row; i += blockDim.x) {
      const int j = i / row;
      s_map[i] = src[(size_t)j * M * W + (i - j * row)];
    }
  }
  __syncthreads();
  if (th.live && c == 0)
    scan_apply_entry_thread<Elem>(tile_start + (((size_t)blockIdx.y * ntiles + blockIdx.x) * M + n) * SW, s_map, s_state, n - n0,
                                  cnt_lat, cnt);
  __syncthreads();
  const long long chunk = first + c;
  if (!th.live || chunk >= nchunks) return;
  scan_apply_thread<Elem>(P, St, a, n, chunk, s_state + ((size_t)c * cnt_lat + (n - n0)) * SW);
}

template <class EZ, class EG, int TH = ScanBounds<EZ>::kThreads>
__global__ void __launch_bounds__(TH, ScanBounds<EZ>::kMinBlocks)
scan_apply2_kernel(const DevProblem* __restrict__ probs, const DevState* __restrict__ states,
                   ScanArgs a, const double* __restrict__ chunk_buf, const double* __restrict__ tile_start) {
  constexpr int W = EZ::kMapDoubles, SW = EZ::kStateDoubles;
  const DevProblem& P = probs[blockIdx.y];
  const DevState& St = states[blockIdx.y];
  const int M = P.M, CH = a.CH;
  const ScanThread2 th(threadIdx.x, P.D, P.N, CH);
  scan_tile_prefetch<EZ>(P, St, a, true);
  extern __shared__ double sm[];                 // per family: [CH][cnt][SW] entering states | [CH][cnt][W] chunk maps
  const long long nchunks = scan_num_chunks(a.nsteps);
  const long long ntiles = scan_num_tiles(a.nsteps, CH);
  const long long first = (long long)blockIdx.x * CH;
  const int cnt = (int)((nchunks - first < CH) ? nchunks - first : CH);
  double* s_state = sm + th.sm_off * (SW + W);
  double* s_map = s_state + (size_t)CH * th.cnt * SW;
  {
    // every family stages its share of the tile's chunk aggregates (cnt * W contiguous doubles per chunk) with its own threads
    const double* src = chunk_buf + (((size_t)blockIdx.y * nchunks + first) * M + th.n0) * W;
    const int row = th.cnt * W;
    for (int i = th.ftid; i < cnt * row; i += th.fthreads) {
      const int j = i / row;
      s_map[i] = src[(size_t)j * M * W + (i - j * row)];
    }
  }
  __syncthreads();
  if (th.live && th.c == 0) {
    const double* tslot = tile_start + (((size_t)blockIdx.y * ntiles + blockIdx.x) * M + th.n) * SW;
    if (th.fam_g) scan_apply_entry_thread<EG>(tslot, s_map, s_state, th.n - th.n0, th.cnt, cnt);
    else scan_apply_entry_thread<EZ>(tslot, s_map, s_state, th.n - th.n0, th.cnt, cnt);
  }
  __syncthreads();
  const long long chunk = first + th.c;
  if (!th.live || chunk >= nchunks) return;
  const double* sslot = s_state + ((size_t)th.c * th.cnt + (th.n - th.n0)) * SW;
  if (th.fam_g) scan_apply_thread<EG>(P, St, a, th.n, chunk, sslot);
  else scan_apply_thread<EZ>(P, St, a, th.n, chunk, sslot);
}

}  // namespace nsagp
