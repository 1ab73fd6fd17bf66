// Device-side exchange between the GPUs that share ONE time-chunked signal (BASELINE.json
// north_star: "NCCL over NVLink only for the O(state^2) per-chunk carry exchange and the lZ reductions").
//
// The payloads are a few KB per pass (scan aggregates, a one-step halo of sites, end-point means, lZ partial
// sums), so the cost of an exchange is latency, not bandwidth.  A host-driven collective per payload costs a
// device->host copy, a stream synchronisation, the collective and a host->device copy (~0.4 ms measured in round 1,
// ~140 of them per EP run).  Here every rank owns a MAILBOX in its own HBM that its peers can write through
// NVLink peer access (cudaIpc handles between processes, plain pointers between emulated ranks of one process):
//
//   mailbox  [kSlots][world][slot_doubles]     record of rank `src` for exchange number `seq` at slot seq % kSlots
//   flags    [kSlots][world]                   = seq + 1 once that record is complete
//
// comm_exchange_kernel (one CTA, one launch per exchange): copy this rank's record into every peer's mailbox with
// plain stores, __threadfence_system(), publish the flags with a system-scope release store, then spin with
// acquire loads on the LOCAL flags until every peer's record of the same exchange number has arrived.  Afterwards
// the local mailbox slot holds all ranks' records and the consumer kernels (scan carry, site unpack, lZ sums) read
// them from HBM.  No host synchronisation, no collective library call, the whole EP schedule is enqueued once.
// A slot is reused kSlots exchanges later; a rank can only get that far ahead after every peer has LAUNCHED the
// exchange after the one whose records it still reads, and stream order puts that launch behind the readers.
#pragma once
#include "common.cuh"

namespace nsagp {

constexpr int kCommMaxWorld = 16;
constexpr int kCommSlots = 4;

struct CommDev {
  double* mailbox[kCommMaxWorld];               // peer-visible base pointers (index = rank)
  unsigned long long* flags[kCommMaxWorld];
  int rank, world;
  long long slot_doubles;
  int* err;                                     // local sticky error flag (1: a peer's record never arrived)
  long long timeout_cycles;
};

__device__ __forceinline__ double* comm_record(const CommDev& c, int owner, unsigned long long seq, int src) {
  return c.mailbox[owner] + (((seq % kCommSlots) * c.world) + src) * c.slot_doubles;
}

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// All-gather of one record per rank.  The local record must already sit in the local mailbox at
// comm_record(c, rank, seq, rank) (the producers write it there directly).
__global__ void __launch_bounds__(256)
comm_exchange_kernel(CommDev c, unsigned long long seq, int ndoubles) {
  const int tid = threadIdx.x;
  const double* src = comm_record(c, c.rank, seq, c.rank);
  for (int p = 0; p < c.world; ++p) {
    if (p == c.rank) continue;
    double* dst = comm_record(c, p, seq, c.rank);
    for (int i = tid; i < ndoubles; i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  if (tid < c.world && tid != c.rank) {
    __threadfence_system();
    st_release_sys(c.flags[tid] + (seq % kCommSlots) * c.world + c.rank, seq + 1);
  }
  if (tid < c.world && tid != c.rank) {
    const unsigned long long* f = c.flags[c.rank] + (seq % kCommSlots) * c.world + tid;
    const long long t0 = clock64();
    while (ld_acquire_sys(f) < seq + 1) {
      if (clock64() - t0 > c.timeout_cycles) { atomicExch(c.err, 1); break; }     // never hang the GPU
      __nanosleep(200);
    }
  }
  __syncthreads();
}

// ---- consumers --------------------------------------------------------------------------------------------
// Copy `nd` doubles at `offset` of the records of ranks first, first+step, ... (count of them) to dst, packed in that
// order: scan aggregates of the shards processed before this one (forward: ranks 0..rank-1 ascending; backward: ranks
// world-1..rank+1 descending), placed in front of the shard's own tile aggregates.
__global__ void comm_gather_kernel(CommDev c, unsigned long long seq, int offset, int nd, int first, int step, int count,
                                   double* __restrict__ dst) {
  const int total = count * nd;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int j = i / nd, w = i - j * nd;
    dst[i] = comm_record(c, c.rank, seq, first + j * step)[offset + w];
  }
}

// ---- record packers / unpackers of the EP schedule (api_comm.inc) -------------------------------------------------
// two plain copies in one launch (b may be null)
__global__ void x_copy2_kernel(double* __restrict__ da, const double* __restrict__ sa, int na, double* __restrict__ db,
                               const double* __restrict__ sb, int nb) {
  for (int i = threadIdx.x; i < na; i += blockDim.x) da[i] = sa[i];
  if (db && sb)
    for (int i = threadIdx.x; i < nb; i += blockDim.x) db[i] = sb[i];
}

// X0: [lZ partial | mean of the rank's last step]
__global__ void x0_pack_kernel(double* __restrict__ rec, const double* __restrict__ lz_partial, const double* __restrict__ m_last, int n) {
  if (threadIdx.x == 0) rec[0] = *lz_partial;
  for (int i = threadIdx.x; i < n; i += blockDim.x) rec[1 + i] = m_last[i];
}

// nlZ(1) = -sum of the partial sums in rank order; boundary mismatch of this rank's first chunk against the mean the
// preceding rank stored for the step before it (bstate0 = null: this rank's first chunk started from step 0, exact).
__global__ void x0_unpack_kernel(CommDev c, unsigned long long seq, double* __restrict__ nlz0, int n,
                                 const double* __restrict__ bstate0, unsigned long long* __restrict__ adfdiag) {
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int r = 0; r < c.world; ++r) s += comm_record(c, c.rank, seq, r)[0];
    *nlz0 = -s;
  }
  if (bstate0 && c.rank > 0) {
    const double* ref = comm_record(c, c.rank, seq, c.rank - 1) + 1;
    double dm = 0.0, mm = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { dm = fmax(dm, fabs(bstate0[i] - ref[i])); mm = fmax(mm, fabs(ref[i])); }
    atomic_max_nonneg(adfdiag, dm);
    atomic_max_nonneg(adfdiag + 1, mm);
  }
}

// X3: [0] lZ partial, [1..2] max-diff bit patterns, [3] unused, [4..4+3M) sites of the rank's last step,
//     [4+3M .. +M*BM) the smoothed mean of step 0 as padded blocks (rank 0: ms0 != null), [.. +M) vm0
__global__ void x3_pack_kernel(double* __restrict__ rec, const double* __restrict__ lz_partial,
                               const unsigned long long* __restrict__ maxdiff, const double* __restrict__ tt,
                               const double* __restrict__ tn, const double* __restrict__ R, int M,
                               const double* __restrict__ ms0, const int* __restrict__ off, int BM, const double* __restrict__ vm0) {
  const int t = threadIdx.x;
  if (t == 0) { rec[0] = lz_partial ? *lz_partial : 0.0; rec[3] = 0.0; }
  if (t < 2) rec[1 + t] = __longlong_as_double((long long)maxdiff[t]);
  if (t < M) {
    rec[4 + t] = tt[t]; rec[4 + M + t] = tn[t]; rec[4 + 2 * M + t] = R[t];
    rec[4 + 3 * M + M * BM + t] = vm0[t];
    const int o = off[t], b = off[t + 1] - o;
    for (int i = 0; i < BM; ++i) rec[4 + 3 * M + t * BM + i] = (ms0 && i < b) ? ms0[o + i] : 0.0;
  }
}

__global__ void x3_unpack_kernel(CommDev c, unsigned long long seq, double* __restrict__ nlz_slot,
                                 unsigned long long* __restrict__ diag_slot, double* __restrict__ tt, double* __restrict__ tn,
                                 double* __restrict__ R, int M, double* __restrict__ mcarry, int BM, double* __restrict__ vm0) {
  const int t = threadIdx.x;
  if (t == 0 && nlz_slot) {
    double s = 0.0;
    for (int r = 0; r < c.world; ++r) s += comm_record(c, c.rank, seq, r)[0];
    *nlz_slot = -s;
  }
  if (t < 2 && diag_slot) {
    unsigned long long m = 0;
    for (int r = 0; r < c.world; ++r) {
      const unsigned long long v = (unsigned long long)__double_as_longlong(comm_record(c, c.rank, seq, r)[1 + t]);
      m = v > m ? v : m;
    }
    diag_slot[t] = m;
  }
  if (t < M) {
    if (tt && c.rank > 0) {
      const double* prev = comm_record(c, c.rank, seq, c.rank - 1);
      tt[t] = prev[4 + t]; tn[t] = prev[4 + M + t]; R[t] = prev[4 + 2 * M + t];
    }
    const double* r0 = comm_record(c, c.rank, seq, 0);
    if (mcarry)
      for (int i = 0; i < BM; ++i) mcarry[t * BM + i] = r0[4 + 3 * M + t * BM + i];
    if (vm0) vm0[t] = r0[4 + 3 * M + M * BM + t];
  }
}

}  // namespace nsagp
