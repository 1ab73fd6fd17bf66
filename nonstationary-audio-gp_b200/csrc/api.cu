// C ABI of libnsagp.so (include/nsagp.h): plan construction, host orchestration
// of the EP schedule, result fetch.  No CPU compute path exists here: every
// entry point either runs the CUDA kernels or returns an error.
#include "../../include/nsagp.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <string>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "ihgp.cuh"
#include "gfep.cuh"
#include "mombatch.cuh"
#include "adfcta.cuh"
#include "ekf.cuh"
#include "ekfscan.cuh"
#include "ekfbig.cuh"
#include "ekfgrad.cuh"
#include "mcrec.cuh"
#include "comm.cuh"
#include "siteupd.cuh"
#include "fastfb.cuh"

using namespace nsagp;

namespace {

thread_local std::string g_err;
thread_local cudaStream_t g_stream = nullptr;
thread_local bool g_own_stream = false;
thread_local bool g_stream_set = false;      // a caller's stream was given (nullptr = the legacy default stream is a valid choice)
thread_local int g_device = -1;              // device the thread-local stream and memory cache belong to
std::atomic<long long> g_launches{0};        // incremented from several host threads (emulated ranks)

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define CU(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(NSAGP_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));        \
  } while (0)

#define LAUNCH_CHECK()                                                                        \
  do {                                                                                        \
    ++g_launches;                                                                             \
    cudaError_t e_ = cudaGetLastError();                                                      \
    if (e_ != cudaSuccess)                                                                    \
      return fail(NSAGP_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e_));   \
  } while (0)

int ensure_stream() {
  if (g_device < 0) CU(cudaGetDevice(&g_device));
  if (!g_stream_set && g_stream == nullptr) {
    CU(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
    g_own_stream = true;
  }
  return NSAGP_OK;
}

int round_bm(int b) {
  const int opts[] = {2, 3, 4, 6, 8};
  for (int o : opts) if (b <= o) return o;
  return -1;
}

// Smallest double x in (lo, hi] with nearest(x) > i, by bisection over bit patterns.
double bisect_threshold(const std::vector<double>& r, int i) {
  auto above = [&](double x) { return nearest_bruteforce(r.data(), (int)r.size(), x) > i; };
  uint64_t lo, hi;
  double a = r[i], b = r[i + 1];
  std::memcpy(&lo, &a, 8);
  std::memcpy(&hi, &b, 8);      // positive doubles: bit patterns are ordered
  // invariant: !above(lo), above(hi)
  while (hi - lo > 1) {
    const uint64_t mid = lo + (hi - lo) / 2;
    double x;
    std::memcpy(&x, &mid, 8);
    if (above(x)) hi = mid; else lo = mid;
  }
  double x;
  std::memcpy(&x, &hi, 8);
  return x;
}

// Largest double t > 0 with fl(1/t) >= thr: the nearest-neighbour decision of the reference is
// taken on R = 1./ttau, and fl(1/t) is non-increasing in t, so comparing ttau with this value
// is the same decision without the division.
double bisect_ttau_threshold(double thr) {
  auto ok = [&](double t) { return 1.0 / t >= thr; };
  double a = 2.3e-308, b = 1e300;            // ok(a), !ok(b) for any thr in (1e-300, 1e300)
  uint64_t lo, hi;
  std::memcpy(&lo, &a, 8);
  std::memcpy(&hi, &b, 8);
  while (hi - lo > 1) {
    const uint64_t mid = lo + (hi - lo) / 2;
    double x;
    std::memcpy(&x, &mid, 8);
    if (ok(x)) lo = mid; else hi = mid;
  }
  double x;
  std::memcpy(&x, &lo, 8);
  return x;
}

// Device memory cache.  The entry points are called thousands of times with the same shapes
// (fminunc around the nlZ mode, demo_toy_modulators_nmf.m:100-104), and cudaMalloc / cudaFree
// cost more than a short EP run, so released blocks are kept (per host thread, exact-size bins
// rounded up to 256 B) and handed out again.  nsagp_release_cache() returns them to the driver.
struct DeviceCache {
  std::multimap<size_t, void*> free_blocks;
  size_t cached_bytes = 0;
  static constexpr size_t kMaxCached = (size_t)64 << 30;
  void* take(size_t nb) {
    auto it = free_blocks.find(nb);
    if (it == free_blocks.end()) return nullptr;
    void* q = it->second;
    free_blocks.erase(it);
    cached_bytes -= nb;
    return q;
  }
  void give(void* q, size_t nb) {
    if (cached_bytes + nb > kMaxCached) { cudaFree(q); return; }
    free_blocks.emplace(nb, q);
    cached_bytes += nb;
  }
  void clear() {
    for (auto& kv : free_blocks) cudaFree(kv.second);
    free_blocks.clear();
    cached_bytes = 0;
  }
};
thread_local DeviceCache g_cache;

constexpr int kMaxPrevShards = 16;      // aggregates of other GPUs' shards that can precede a plan's scan tiles

struct DeviceArena {
  std::vector<std::pair<void*, size_t>> ptrs;
  size_t bytes = 0;
  template <class T>
  int alloc(T** p, size_t count) {
    const size_t nb = ((std::max<size_t>(count, 1) * sizeof(T)) + 255) & ~(size_t)255;
    void* q = g_cache.take(nb);
    if (!q) {
      cudaError_t e = cudaMalloc(&q, nb);
      if (e != cudaSuccess) {                       // out of memory: give the cache back and retry once
        cudaGetLastError();
        g_cache.clear();
        e = cudaMalloc(&q, nb);
      }
      if (e != cudaSuccess) return fail(NSAGP_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    }
    ptrs.emplace_back(q, nb);
    bytes += nb;
    *p = static_cast<T*>(q);
    return NSAGP_OK;
  }
  template <class T>
  int upload(T** p, const std::vector<T>& h) {
    int rc = alloc(p, h.size());
    if (rc) return rc;
    if (!h.empty()) CU(cudaMemcpy(*p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return NSAGP_OK;
  }
  void release() {
    for (auto& q : ptrs) g_cache.give(q.first, q.second);
    ptrs.clear();
  }
};

}  // namespace

struct nsagp_plan {
  int kind = 0, B = 0, mode = 0;
  long long T = 0;
  int D = 0, N = 0, M = 0, bz = 0, bg = 0, BM = 0, n = 0, DP = 0, S = 0, nr = 0, PB = 0;
  double alpha = 1.0;
  std::vector<double> damping;
  int ep_itts = 1;
  DeviceArena arena;
  std::vector<DevProblem> h_probs;
  std::vector<DevState> h_states;
  DevProblem* d_probs = nullptr;
  DevState* d_states = nullptr;
  double* d_chunk = nullptr;    // scan: one map per (chunk, block)
  double* d_tile = nullptr;     // scan: one map per (CTA tile, block)
  double* d_start = nullptr;    // scan: state entering each CTA tile
  double* d_total = nullptr;    // scan: this plan's single aggregate (time-chunked runs)
  double* d_seg = nullptr;      // two-level carry: one map per (segment of tiles, block)
  double* d_segstart = nullptr; // two-level carry: state entering each segment
  long long nseg_max = 0;
  long long t0 = 0, t1 = -1;    // shard of the frozen-site passes this plan executes (time-chunked runs)
  double* d_nlZ = nullptr;      // [B][ep_itts + 1]  (last slot: edata)
  double* d_diag = nullptr;     // [B][ep_itts][2]
  double* d_MF = nullptr;       // [B][T][n] filtered means of the last pass (predict mode)
  double* d_PF = nullptr;       // [B][T][PB] filtered covariances of the last pass (full-state, predict)
  double* d_y = nullptr;
  double* d_bounds = nullptr;   // [3][T][M] Varft / lb / ub staging for fetch
  std::vector<double> h_vminf;  // [B][M] h Pinf h'
  cudaEvent_t ev[2] = {nullptr, nullptr};
  std::vector<cudaEvent_t> phase_ev;
  double timings[5] = {0, 0, 0, 0, 0};
  int adf_form = 0;             // 0: CTA per signal, width by the size of the launch; 1: warp per signal; 2: half-width CTA, two per SM; 3: full-width CTA
  int adf_chunks = 0;           // > 1: first filter pass parallel in time, burn-in overlap (opt-in, approximate; adfcta.cuh AdfPar)
  long long adf_burn = 0;
  double* d_bstate = nullptr;   // [B][adf_chunks][n] state at the end of each chunk's burn-in
  unsigned long long* d_adfdiag = nullptr;   // [2] bit patterns: max |mismatch|, max |mean| at the chunk boundaries
  bool ran = false;
};

namespace {

int validate_shapes(int B, const nsagp_model* models, const nsagp_lik* liks, const nsagp_ep* ep, long long T) {
  if (B < 1 || !models || !liks || !ep) return fail(NSAGP_ERR_INVALID, "null argument or B < 1");
  if (T < 1) return fail(NSAGP_ERR_INVALID, "T must be >= 1");
  if (ep->ep_itts < 1 || !ep->ep_damping) return fail(NSAGP_ERR_INVALID, "ep_itts >= 1 and ep_damping[ep_itts] required");
  if (!(ep->ep_fraction > 0.0)) return fail(NSAGP_ERR_INVALID, "ep_fraction must be positive");
  const nsagp_model& m0 = models[0];
  if (m0.D < 1 || m0.N < 1 || m0.N > kNP) return fail(NSAGP_ERR_INVALID, "need D >= 1 and 1 <= N <= 4 modulators");
  if (m0.D + m0.N > 32) return fail(NSAGP_ERR_INVALID, "D + N must be <= 32 (one warp lane per latent)");
  if (m0.bz < 1 || m0.bz > kMaxBlock || m0.bg < 1 || m0.bg > kMaxBlock)
    return fail(NSAGP_ERR_INVALID, "block sizes must be in 1..8");
  for (int b = 0; b < B; ++b) {
    const nsagp_model& m = models[b];
    if (m.D != m0.D || m.N != m0.N || m.bz != m0.bz || m.bg != m0.bg)
      return fail(NSAGP_ERR_INVALID, "all problems of a batch must share D, N, bz, bg");
    if (!m.A || !m.Q || !m.Pinf || !m.h) return fail(NSAGP_ERR_INVALID, "null model array");
    const nsagp_lik& l = liks[b];
    if (l.kind != 0 && l.kind != 1) return fail(NSAGP_ERR_INVALID, "lik.kind must be 0 or 1");
    if (l.S < 1 || l.S != liks[0].S || !l.W || !l.wn || !l.xn)
      return fail(NSAGP_ERR_INVALID, "likelihood tables missing or S differs within the batch");
  }
  return NSAGP_OK;
}

int build_lik_arrays(const nsagp_lik& l, int D, int N, int DP, std::vector<double>& W, std::vector<double>& wn,
                     std::vector<double>& xn) {
  W.assign((size_t)DP * kNP, 0.0);
  for (int d = 0; d < D; ++d)
    for (int j = 0; j < N; ++j) W[(size_t)d * kNP + j] = l.W[d + (size_t)j * D];
  wn.assign(l.wn, l.wn + l.S);
  xn.assign((size_t)kNP * l.S, 0.0);
  for (int s = 0; s < l.S; ++s)
    for (int j = 0; j < N; ++j) xn[(size_t)j * l.S + s] = l.xn[j + (size_t)s * N];
  return NSAGP_OK;
}

}  // namespace

extern "C" {

const char* nsagp_version(void) { return "nsagp-b200 0.1 (sm_100a)"; }
const char* nsagp_last_error(void) { return g_err.c_str(); }

int nsagp_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int nsagp_set_device(int device) {
  if (g_device >= 0 && g_device != device) {
    // the thread's stream and cached blocks live on the previous device: neither may be used on the new one
    if (g_own_stream && g_stream) { cudaStreamSynchronize(g_stream); cudaStreamDestroy(g_stream); }
    g_stream = nullptr; g_own_stream = false; g_stream_set = false;
    g_cache.clear();
  }
  CU(cudaSetDevice(device));
  g_device = device;
  return NSAGP_OK;
}

int nsagp_set_stream(void* s) {
  if (g_own_stream && g_stream) { cudaStreamSynchronize(g_stream); cudaStreamDestroy(g_stream); }
  g_own_stream = false;
  g_stream = static_cast<cudaStream_t>(s);      // 0 = the legacy default stream (what torch.cuda.current_stream() is by default)
  g_stream_set = true;
  return NSAGP_OK;
}

int64_t nsagp_launch_count(int reset) {
  return reset ? g_launches.exchange(0) : g_launches.load();
}

// ------------------------------------------------------------------ mom batch
static int mom_batch_impl(const nsagp_lik* lik, int32_t D, int32_t N, double alpha, int64_t T, const double* y,
                          const double* mu, const double* s2, double* lZ, double* dlZ, double* d2lZ, int warp_form) {
  if (!lik || !y || !mu || !s2 || !lZ || !dlZ || !d2lZ) return fail(NSAGP_ERR_INVALID, "null argument");
  if (D < 1 || N < 1 || N > kNP || D + N > 32) return fail(NSAGP_ERR_INVALID, "need 1 <= N <= 4, D + N <= 32");
  if (lik->S < 1 || T < 0) return fail(NSAGP_ERR_INVALID, "bad S or T");
  if (T == 0) return NSAGP_OK;
  int rc = ensure_stream();
  if (rc) return rc;
  const int M = D + N, DP = (D <= 16) ? 16 : 32;
  std::vector<double> W, wn, xn;
  build_lik_arrays(*lik, D, N, DP, W, wn, xn);
  DeviceArena ar;
  MomBatchArgs a;
  double *dW, *dwn, *dxn, *dy, *dmu, *ds2, *dl, *d1, *d2;
  if ((rc = ar.upload(&dW, W)) || (rc = ar.upload(&dwn, wn)) || (rc = ar.upload(&dxn, xn))) { ar.release(); return rc; }
  if ((rc = ar.alloc(&dy, T)) || (rc = ar.alloc(&dmu, T * M)) || (rc = ar.alloc(&ds2, T * M)) ||
      (rc = ar.alloc(&dl, T)) || (rc = ar.alloc(&d1, T * M)) || (rc = ar.alloc(&d2, T * M))) { ar.release(); return rc; }
  auto body = [&]() -> int {
    CU(cudaMemcpyAsync(dy, y, T * sizeof(double), cudaMemcpyHostToDevice, g_stream));
    CU(cudaMemcpyAsync(dmu, mu, T * M * sizeof(double), cudaMemcpyHostToDevice, g_stream));
    CU(cudaMemcpyAsync(ds2, s2, T * M * sizeof(double), cudaMemcpyHostToDevice, g_stream));
    a.p.D = D; a.p.N = N; a.p.S = lik->S; a.p.kind = lik->kind; a.p.sn2 = lik->sn2; a.p.shift = lik->link_shift;
    a.p.W = dW; a.p.wn = dwn; a.p.xn = dxn;
    a.alpha = alpha; a.T = T; a.M = M; a.y = dy; a.mu = dmu; a.s2 = ds2; a.lZ = dl; a.d1 = d1; a.d2 = d2;
    const size_t lik_sm = ((size_t)DP * kNP + (size_t)lik->S * (1 + kNP)) * sizeof(double);
    if (warp_form) {
      constexpr int WPB = 4;
      const size_t sm = lik_sm + WPB * 64 * sizeof(double);
      const unsigned grid = (unsigned)((T + WPB - 1) / WPB);
      if (DP == 16) mom_batch_warp_kernel<16, WPB><<<grid, 32 * WPB, sm, g_stream>>>(a);
      else mom_batch_warp_kernel<32, WPB><<<grid, 32 * WPB, sm, g_stream>>>(a);
    } else {
      constexpr int TPB = 64;
      const size_t sm = lik_sm + (size_t)4 * M * TPB * sizeof(double);
      const unsigned grid = (unsigned)((T + TPB - 1) / TPB);
      if (DP == 16) {
        CU(cudaFuncSetAttribute(mom_batch_thread_kernel<16, TPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        mom_batch_thread_kernel<16, TPB><<<grid, TPB, sm, g_stream>>>(a);
      } else {
        CU(cudaFuncSetAttribute(mom_batch_thread_kernel<32, TPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        mom_batch_thread_kernel<32, TPB><<<grid, TPB, sm, g_stream>>>(a);
      }
    }
    LAUNCH_CHECK();
    CU(cudaMemcpyAsync(lZ, dl, T * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
    CU(cudaMemcpyAsync(dlZ, d1, T * M * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
    CU(cudaMemcpyAsync(d2lZ, d2, T * M * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
    CU(cudaStreamSynchronize(g_stream));
    return NSAGP_OK;
  };
  rc = body();
  ar.release();
  return rc;
}

int nsagp_mom_batch(const nsagp_lik* lik, int32_t D, int32_t N, double ep_fraction, int64_t T, const double* y,
                    const double* mu, const double* s2, double* lZ, double* dlZ, double* d2lZ) {
  return mom_batch_impl(lik, D, N, ep_fraction, T, y, mu, s2, lZ, dlZ, d2lZ, 0);
}

int nsagp_mom_batch_warp(const nsagp_lik* lik, int32_t D, int32_t N, double ep_fraction, int64_t T, const double* y,
                         const double* mu, const double* s2, double* lZ, double* dlZ, double* d2lZ) {
  return mom_batch_impl(lik, D, N, ep_fraction, T, y, mu, s2, lZ, dlZ, d2lZ, 1);
}

// ------------------------------------------------------------- fastmath hook
}  // extern "C"

namespace {
__global__ void fastmath_kernel(int op, long long n, const double* __restrict__ x, double* __restrict__ out) {
  __shared__ __align__(16) double s_logtab[kLogTabDoubles];
  log_tab_fill(s_logtab, threadIdx.x, blockDim.x);
  __syncthreads();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double v = x[i];
  double r;
  switch (op) {
    case 0: r = rcp_fast(v); break;
    case 1: r = rsqrt_fast(v); break;
    case 2: r = sqrt_fast(v); break;
    case 3: r = exp_fast(v); break;
    case 4: r = log_ge1_fast(v); break;
    case 5: r = softplus_fast(v); break;
    case 6: r = rcp_fast2(v); break;
    case 7: r = rsqrt_fast2(v); break;
    case 8: r = sqrt_fast2(v); break;
    case 9: r = log_ge1_tab(v, (unsigned)__cvta_generic_to_shared(s_logtab)); break;
    default: r = softplus_tab_finite(v, (unsigned)__cvta_generic_to_shared(s_logtab)); break;
  }
  out[i] = r;
}
}  // namespace

extern "C" {

int nsagp_fastmath_eval(int32_t op, int64_t n, const double* x, double* out) {
  if (op < 0 || op > 10 || n < 0 || !x || !out) return fail(NSAGP_ERR_INVALID, "bad argument");
  if (n == 0) return NSAGP_OK;
  int rc = ensure_stream();
  if (rc) return rc;
  DeviceArena ar;
  double *dx, *dout;
  if ((rc = ar.alloc(&dx, n)) || (rc = ar.alloc(&dout, n))) { ar.release(); return rc; }
  auto body = [&]() -> int {
    CU(cudaMemcpyAsync(dx, x, n * sizeof(double), cudaMemcpyHostToDevice, g_stream));
    fastmath_kernel<<<(unsigned)((n + 255) / 256), 256, 0, g_stream>>>(op, n, dx, dout);
    LAUNCH_CHECK();
    CU(cudaMemcpyAsync(out, dout, n * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
    CU(cudaStreamSynchronize(g_stream));
    return NSAGP_OK;
  };
  rc = body();
  ar.release();
  return rc;
}

// ----------------------------------------------------------------------- plan
int nsagp_plan_create(nsagp_plan** out_plan, int32_t kind, int32_t B, const nsagp_model* models,
                      const nsagp_lik* liks, const nsagp_ep* ep, const nsagp_tables* tables, const double* y,
                      int64_t T, int32_t mode) {
  if (!out_plan) return fail(NSAGP_ERR_INVALID, "null plan pointer");
  *out_plan = nullptr;
  int rc = validate_shapes(B, models, liks, ep, T);
  if (rc) return rc;
  if (kind != 0 && kind != 1) return fail(NSAGP_ERR_INVALID, "plan kind must be 0 (ihgp) or 1 (full-state EP)");
  if (mode < 0 || mode > 2) return fail(NSAGP_ERR_INVALID, "bad mode");
  if (kind == 1 && mode == NSAGP_MODE_NLZ_RUNNING) return fail(NSAGP_ERR_INVALID, "running-site mode is ihgp only");
  if (!y) return fail(NSAGP_ERR_INVALID, "null y");
  if (kind == 0) {
    if (!tables) return fail(NSAGP_ERR_INVALID, "ihgp needs tables");
    for (int b = 0; b < B; ++b) {
      if (tables[b].nr < 2 || tables[b].nr != tables[0].nr || !tables[b].r || !tables[b].PP)
        return fail(NSAGP_ERR_INVALID, "bad ihgp tables");
      if (mode == NSAGP_MODE_PREDICT && !tables[b].PG) return fail(NSAGP_ERR_INVALID, "predict mode needs smoother tables PG");
    }
  }
  if ((rc = ensure_stream())) return rc;

  nsagp_plan* pl = new nsagp_plan();
  pl->kind = kind; pl->B = B; pl->mode = mode; pl->T = T;
  const nsagp_model& m0 = models[0];
  pl->D = m0.D; pl->N = m0.N; pl->M = m0.D + m0.N; pl->bz = m0.bz; pl->bg = m0.bg;
  pl->BM = round_bm(std::max(m0.bz, m0.bg));
  pl->n = m0.D * m0.bz + m0.N * m0.bg;
  pl->DP = (m0.D <= 16) ? 16 : 32;
  pl->S = liks[0].S;
  pl->nr = (kind == 0) ? tables[0].nr : 0;
  pl->alpha = ep->ep_fraction;
  pl->ep_itts = ep->ep_itts;
  pl->damping.assign(ep->ep_damping, ep->ep_damping + ep->ep_itts);
  const int D = pl->D, N = pl->N, M = pl->M, BM = pl->BM, nr = pl->nr, DP = pl->DP;
  pl->PB = M * BM * BM;
  auto bsize = [&](int i) { return i < D ? pl->bz : pl->bg; };

  auto cleanup = [&](int code) {
    pl->arena.release();
    delete pl;
    return code;
  };

  std::vector<int> off(M + 1, 0);
  for (int i = 0; i < M; ++i) off[i + 1] = off[i] + bsize(i);
  int* d_off = nullptr;
  if ((rc = pl->arena.upload(&d_off, off))) return cleanup(rc);

  pl->h_probs.resize(B);
  pl->h_states.resize(B);
  pl->h_vminf.assign((size_t)B * M, 0.0);
  const bool predict = (mode == NSAGP_MODE_PREDICT);

  for (int b = 0; b < B; ++b) {
    const nsagp_model& m = models[b];
    const nsagp_lik& l = liks[b];
    DevProblem& P = pl->h_probs[b];
    std::memset(&P, 0, sizeof(P));
    P.D = D; P.N = N; P.M = M; P.BM = BM; P.n = pl->n; P.bz = pl->bz; P.bg = pl->bg;
    P.S = l.S; P.nr = nr; P.lik_kind = l.kind; P.sn2 = l.sn2; P.link_shift = l.link_shift;
    P.off = d_off;
    std::vector<double> A((size_t)M * BM * BM, 0.0), Q(A.size(), 0.0), Pi(A.size(), 0.0), h((size_t)M * BM, 0.0),
        hA((size_t)M * BM, 0.0);
    size_t po = 0, ho = 0;
    for (int i = 0; i < M; ++i) {
      const int bs = bsize(i);
      for (int c = 0; c < bs; ++c)
        for (int r = 0; r < bs; ++r) {
          A[(size_t)i * BM * BM + r + c * BM] = m.A[po + r + c * bs];
          Q[(size_t)i * BM * BM + r + c * BM] = m.Q[po + r + c * bs];
          Pi[(size_t)i * BM * BM + r + c * BM] = m.Pinf[po + r + c * bs];
        }
      for (int r = 0; r < bs; ++r) h[(size_t)i * BM + r] = m.h[ho + r];
      for (int c = 0; c < bs; ++c) {
        double s = 0.0;
        for (int r = 0; r < bs; ++r) s += m.h[ho + r] * m.A[po + r + c * bs];
        hA[(size_t)i * BM + c] = s;
      }
      po += (size_t)bs * bs; ho += bs;
    }
    std::vector<double> W, wn, xn;
    build_lik_arrays(l, D, N, DP, W, wn, xn);
    {
      // distinct sigma-point coordinates per modulator (common.cuh: DevProblem::ndist)
      std::vector<double> xd((size_t)kNP * kMaxDist, 0.0);
      std::vector<unsigned char> xi((size_t)kNP * l.S, 0);
      int nd = 0;
      bool ok = true;
      for (int j = 0; j < N && ok; ++j) {
        int cnt = 0;
        for (int s_ = 0; s_ < l.S && ok; ++s_) {
          const double v = xn[(size_t)j * l.S + s_];
          int q = 0;
          while (q < cnt && std::memcmp(&xd[(size_t)j * kMaxDist + q], &v, 8) != 0) ++q;
          if (q == cnt) { if (cnt == kMaxDist) { ok = false; break; } xd[(size_t)j * kMaxDist + cnt++] = v; }
          xi[(size_t)j * l.S + s_] = (unsigned char)q;
        }
        nd = std::max(nd, cnt);
      }
      P.ndist = 0;
      static const bool no_tab = std::getenv("NSAGP_NO_LINK_TABLE") != nullptr;      // (A/B measurements)
      if (ok && !no_tab && nd > 0 && nd * 3 <= l.S) {            // worth it only if it saves most link evaluations
        std::vector<double> xdp((size_t)kNP * nd, 0.0);
        for (int j = 0; j < kNP; ++j)
          for (int q = 0; q < nd; ++q) xdp[(size_t)j * nd + q] = xd[(size_t)j * kMaxDist + q];
        double* dxd; unsigned char* dxi;
        if ((rc = pl->arena.upload(&dxd, xdp)) || (rc = pl->arena.upload(&dxi, xi))) return cleanup(rc);
        P.ndist = nd; P.xdist = dxd; P.xidx = dxi;
      }
    }
    double *dA, *dQ, *dPi, *dh, *dhA, *dW, *dwn, *dxn;
    if ((rc = pl->arena.upload(&dA, A)) || (rc = pl->arena.upload(&dQ, Q)) || (rc = pl->arena.upload(&dPi, Pi)) ||
        (rc = pl->arena.upload(&dh, h)) || (rc = pl->arena.upload(&dhA, hA)) || (rc = pl->arena.upload(&dW, W)) ||
        (rc = pl->arena.upload(&dwn, wn)) || (rc = pl->arena.upload(&dxn, xn)))
      return cleanup(rc);
    P.A = dA; P.Q = dQ; P.Pinf = dPi; P.h = dh; P.hA = dhA; P.W = dW; P.wn = dwn; P.xn = dxn;

    std::vector<double> vm0(M, 0.0);
    if (kind == 0) {
      const nsagp_tables& tb = tables[b];
      std::vector<double> r(tb.r, tb.r + nr), thr(nr - 1), cthr(nr - 1);
      for (int i = 0; i + 1 < nr; ++i) {
        if (!(r[i] > 0.0) || !(r[i + 1] > r[i])) return cleanup(fail(NSAGP_ERR_INVALID, "table grid r must be positive and ascending"));
        thr[i] = bisect_threshold(r, i);
        cthr[i] = bisect_ttau_threshold(thr[i]);
      }
      std::vector<double> Wtab((size_t)M * (nr + 1) * BM, 0.0), HPH((size_t)M * (nr + 1), 0.0);
      std::vector<double> Gtab, vmtab;
      if (tb.PG) { Gtab.assign((size_t)M * nr * BM * BM, 0.0); vmtab.assign((size_t)M * nr, 0.0); }
      size_t ppo = 0, pgo = 0;
      ho = 0; po = 0;
      for (int i = 0; i < M; ++i) {
        const int bs = bsize(i);
        const double* hi = m.h + ho;
        for (int row = 0; row <= nr; ++row) {
          const double* PP = (row < nr) ? tb.PP + ppo + (size_t)row * bs * bs : m.Pinf + po;
          double hph = 0.0;
          for (int r_ = 0; r_ < bs; ++r_) {
            double w = 0.0;
            for (int c = 0; c < bs; ++c) w += PP[r_ + c * bs] * hi[c];     // W = PP*h'
            Wtab[((size_t)i * (nr + 1) + row) * BM + r_] = w;
          }
          for (int r_ = 0; r_ < bs; ++r_) hph += hi[r_] * Wtab[((size_t)i * (nr + 1) + row) * BM + r_];
          HPH[(size_t)i * (nr + 1) + row] = hph;
        }
        vm0[i] = HPH[(size_t)i * (nr + 1) + nr];
        if (tb.PG) {
          for (int row = 0; row < nr; ++row) {
            const double* Ps = tb.PG + pgo + (size_t)row * 2 * bs * bs;
            const double* G = Ps + bs * bs;
            for (int c = 0; c < bs; ++c)
              for (int r_ = 0; r_ < bs; ++r_) Gtab[((size_t)i * nr + row) * BM * BM + r_ + c * BM] = G[r_ + c * bs];
            double v = 0.0;
            for (int c = 0; c < bs; ++c) {
              double hp = 0.0;
              for (int r_ = 0; r_ < bs; ++r_) hp += hi[r_] * Ps[r_ + c * bs];   // (h*P)
              v += hp * hi[c];
            }
            vmtab[(size_t)i * nr + row] = v;
          }
          pgo += (size_t)nr * 2 * bs * bs;
        }
        ppo += (size_t)nr * bs * bs; po += (size_t)bs * bs; ho += bs;
      }
      std::vector<double> SDt((size_t)N * (nr + 1)), RS2t((size_t)N * (nr + 1));
      for (int j = 0; j < N; ++j)
        for (int row = 0; row <= nr; ++row) {
          const double v = HPH[(size_t)(D + j) * (nr + 1) + row];
          SDt[(size_t)j * (nr + 1) + row] = std::sqrt(v);
          RS2t[(size_t)j * (nr + 1) + row] = 1.0 / v;
        }
      // window look-up (adfcta.cuh: lookup_by_ttau) reads 10 thresholds around the previous row without bounds checks:
      // kCthrPad sentinels on either side (+Inf: "ttau <= c" holds; -Inf: it does not)
      std::vector<double> cthr_pad(cthr.size() + 2 * kCthrPad);
      for (int i = 0; i < kCthrPad; ++i) { cthr_pad[i] = INFINITY; cthr_pad[kCthrPad + cthr.size() + i] = -INFINITY; }
      std::copy(cthr.begin(), cthr.end(), cthr_pad.begin() + kCthrPad);
      std::vector<double> thr_pad(thr.size() + 2 * kCthrPad);      // "thr <= R": -Inf below (holds), +Inf above (does not)
      for (int i = 0; i < kCthrPad; ++i) { thr_pad[i] = -INFINITY; thr_pad[kCthrPad + thr.size() + i] = INFINITY; }
      std::copy(thr.begin(), thr.end(), thr_pad.begin() + kCthrPad);
      {
        // #{thr <= R} ~ floor(position of R on the log-spaced grid + 0.49), log2 R ~ hi32(R) / 2^20 - 1023 + 0.043
        const double scale = (double)(nr - 1) / (std::log10(r[nr - 1]) - std::log10(r[0]));
        const double l2 = 0.30102999566398120 * scale;
        P.rg_b = l2 / 1048576.0;
        P.rg_a = -std::log10(r[0]) * scale - (1023.0 - 0.043) * l2 + 0.49;
      }
      double *dr, *dthr, *dcthr, *dWt, *dHPH, *dSD, *dRS2, *dG = nullptr, *dvm = nullptr;
      if ((rc = pl->arena.upload(&dr, r)) || (rc = pl->arena.upload(&dthr, thr_pad)) || (rc = pl->arena.upload(&dWt, Wtab)) ||
          (rc = pl->arena.upload(&dHPH, HPH)) || (rc = pl->arena.upload(&dcthr, cthr_pad)) ||
          (rc = pl->arena.upload(&dSD, SDt)) || (rc = pl->arena.upload(&dRS2, RS2t)))
        return cleanup(rc);
      P.cthr = dcthr + kCthrPad; P.SDtab = dSD; P.RS2tab = dRS2;
      if (tb.PG && ((rc = pl->arena.upload(&dG, Gtab)) || (rc = pl->arena.upload(&dvm, vmtab)))) return cleanup(rc);
      P.r = dr; P.thr = dthr + kCthrPad; P.Wtab = dWt; P.HPHtab = dHPH; P.Gtab = dG; P.vmtab = dvm;
    } else {
      ho = 0; po = 0;
      for (int i = 0; i < M; ++i) {
        const int bs = bsize(i);
        double v = 0.0;
        for (int c = 0; c < bs; ++c) {
          double hp = 0.0;
          for (int r_ = 0; r_ < bs; ++r_) hp += m.h[ho + r_] * m.Pinf[po + r_ + c * bs];
          v += hp * m.h[ho + c];
        }
        vm0[i] = v;
        po += (size_t)bs * bs; ho += bs;
      }
    }
    for (int i = 0; i < M; ++i) pl->h_vminf[(size_t)b * M + i] = vm0[i];

    DevState& St = pl->h_states[b];
    std::memset(&St, 0, sizeof(St));
    double *dy, *dtt, *dtn, *dR, *dMS, *dE, *dlZ, *dmc, *dvm0, *dV = nullptr, *dPS = nullptr;
    unsigned long long *dmax, *dneg;
    int* dstat;
    if ((rc = pl->arena.alloc(&dy, T)) || (rc = pl->arena.alloc(&dtt, T * M)) || (rc = pl->arena.alloc(&dtn, T * M)) ||
        (rc = pl->arena.alloc(&dR, T * M)) || (rc = pl->arena.alloc(&dMS, T * pl->n)) ||
        (rc = pl->arena.alloc(&dE, T * M)) || (rc = pl->arena.alloc(&dlZ, T)) ||
        (rc = pl->arena.alloc(&dmc, (size_t)M * BM)) || (rc = pl->arena.upload(&dvm0, vm0)) ||
        (rc = pl->arena.alloc(&dmax, 2)) || (rc = pl->arena.alloc(&dneg, 1)) || (rc = pl->arena.alloc(&dstat, 1)))
      return cleanup(rc);
    if (kind == 1 && (predict || pl->ep_itts > 1)) {     // the smoother needs the filtered covariances
      if ((rc = pl->arena.alloc(&dV, T * M)) || (rc = pl->arena.alloc(&dPS, (size_t)T * pl->PB))) return cleanup(rc);
    }
    St.y = dy; St.ttau = dtt; St.tnu = dtn; St.R = dR; St.MS = dMS; St.E = dE; St.V = dV; St.lZ = dlZ; St.PS = dPS;
    St.mcarry = dmc; St.vm0 = dvm0; St.maxdiff = dmax; St.negcav = dneg; St.status = dstat;
    cudaError_t e = cudaMemcpy(dy, y + (size_t)b * T, T * sizeof(double), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return cleanup(fail(NSAGP_ERR_CUDA, std::string("upload y: ") + cudaGetErrorString(e)));
  }
  if ((rc = pl->arena.upload(&pl->d_probs, pl->h_probs)) || (rc = pl->arena.upload(&pl->d_states, pl->h_states)))
    return cleanup(rc);
  const long long nchunks = scan_num_chunks(T);
  const size_t map_d = (kind == 0) ? (size_t)(BM * BM + BM) : (size_t)(3 * BM * BM + 2 * BM);   // largest scan element
  const size_t state_d = (kind == 0) ? (size_t)BM : (size_t)(BM * BM + BM);
  const long long ntiles_max = scan_num_tiles(T, 1);        // CH >= 1
  // (kMaxPrevShards slots in front of the tile arrays: aggregates of other GPUs' shards, time-chunked runs)
  if ((rc = pl->arena.alloc(&pl->d_chunk, (size_t)B * nchunks * M * map_d)) ||
      (rc = pl->arena.alloc(&pl->d_tile, ((size_t)B * ntiles_max + kMaxPrevShards) * M * map_d)) ||
      (rc = pl->arena.alloc(&pl->d_start, ((size_t)B * ntiles_max + kMaxPrevShards) * M * state_d)) ||
      (rc = pl->arena.alloc(&pl->d_total, (size_t)M * map_d)) ||
      (rc = pl->arena.alloc(&pl->d_seg, (size_t)B * (pl->nseg_max = (long long)std::ceil(std::sqrt((double)(ntiles_max + kMaxPrevShards))) + 2) * M * map_d)) ||
      (rc = pl->arena.alloc(&pl->d_segstart, (size_t)B * pl->nseg_max * M * state_d)) ||
      (rc = pl->arena.alloc(&pl->d_nlZ, (size_t)B * (pl->ep_itts + 1))) ||
      (rc = pl->arena.alloc(&pl->d_diag, (size_t)B * pl->ep_itts * 2)))
    return cleanup(rc);
  pl->d_tile += (size_t)kMaxPrevShards * M * map_d;
  pl->d_start += (size_t)kMaxPrevShards * M * state_d;
  if (predict && (rc = pl->arena.alloc(&pl->d_MF, (size_t)B * T * pl->n))) return cleanup(rc);
  cudaEventCreate(&pl->ev[0]);
  cudaEventCreate(&pl->ev[1]);
  *out_plan = pl;
  return NSAGP_OK;
}

int nsagp_plan_keep_pf(nsagp_plan* pl, int keep) {
  if (!pl) return fail(NSAGP_ERR_INVALID, "null plan");
  if (keep && !pl->d_PF) {
    if (pl->kind != 1 || pl->mode != NSAGP_MODE_PREDICT)
      return fail(NSAGP_ERR_INVALID, "filtered covariances exist only in the full-state predict mode");
    return pl->arena.alloc(&pl->d_PF, (size_t)pl->B * pl->T * pl->PB);
  }
  return NSAGP_OK;
}

long long g_fam_min_steps = 0;                  // see DISPATCH_FAM below
int nsagp_scan_config(int64_t family_min_steps) {
  if (family_min_steps < 0) return fail(NSAGP_ERR_INVALID, "family_min_steps must be >= 0");
  g_fam_min_steps = family_min_steps;
  return NSAGP_OK;
}

int g_site_form = std::getenv("NSAGP_SITE_FORM") ? std::atoi(std::getenv("NSAGP_SITE_FORM")) : 0;
int nsagp_site_config(int32_t form) {
  if (form < 0 || form > 2) return fail(NSAGP_ERR_INVALID, "site-update form must be 0, 1 or 2");
  g_site_form = form;
  return NSAGP_OK;
}

int g_scan_merge = 1, g_scan_l2pf = 1, g_scan_ch_max = 16, g_scan_threads2 = 256;
int nsagp_scan_tile(int32_t max_chunks, int32_t max_threads) {
  if (max_chunks < 1 || max_chunks > 16) return fail(NSAGP_ERR_INVALID, "max_chunks must be in 1..16");
  if (max_threads != 320 && max_threads != 256) return fail(NSAGP_ERR_INVALID, "max_threads must be 320 or 256");
  g_scan_ch_max = max_chunks;
  g_scan_threads2 = max_threads;
  return NSAGP_OK;
}
int nsagp_scan_merge(int32_t on) {
  g_scan_merge = on ? 1 : 0;
  return NSAGP_OK;
}
int nsagp_scan_prefetch(int32_t on) {
  g_scan_l2pf = on ? 1 : 0;
  return NSAGP_OK;
}

int nsagp_release_cache(void) {
  g_cache.clear();
  return NSAGP_OK;
}

int nsagp_plan_set_adf_form(nsagp_plan* pl, int form) {
  if (!pl) return fail(NSAGP_ERR_INVALID, "null plan");
  if (form < 0 || form > 3) return fail(NSAGP_ERR_INVALID, "adf form must be 0 (automatic), 1 (warp per signal), 2 (half-width CTA, two per SM) or 3 (full-width CTA)");
  pl->adf_form = form;
  return NSAGP_OK;
}

int nsagp_plan_set_adf_parallel(nsagp_plan* pl, int32_t chunks, int64_t burnin) {
  if (!pl) return fail(NSAGP_ERR_INVALID, "null plan");
  if (chunks < 0 || burnin < 0) return fail(NSAGP_ERR_INVALID, "chunks and burnin must be >= 0");
  if (chunks > 1 && burnin < 1) return fail(NSAGP_ERR_INVALID, "a parallel first pass needs a burn-in of at least one step");
  if (chunks > 4096) return fail(NSAGP_ERR_INVALID, "at most 4096 chunks");
  pl->adf_chunks = chunks > 1 ? chunks : 0;
  pl->adf_burn = burnin;
  if (pl->adf_chunks && !pl->d_adfdiag) {
    int rc = pl->arena.alloc(&pl->d_adfdiag, 2);
    if (rc) return rc;
    CU(cudaMemset(pl->d_adfdiag, 0, 16));
  }
  if (pl->adf_chunks) {       // (re)size the burn-in state record
    int rc = pl->arena.alloc(&pl->d_bstate, (size_t)pl->B * pl->adf_chunks * pl->n);
    if (rc) return rc;
  }
  return NSAGP_OK;
}

int nsagp_plan_adf_mismatch(nsagp_plan* pl, double* out2) {
  if (!pl || !out2) return fail(NSAGP_ERR_INVALID, "null argument");
  out2[0] = out2[1] = 0.0;
  if (!pl->d_adfdiag) return NSAGP_OK;
  CU(cudaMemcpy(out2, pl->d_adfdiag, 16, cudaMemcpyDeviceToHost));     // bit patterns of non-negative doubles
  return NSAGP_OK;
}

int nsagp_plan_destroy(nsagp_plan* pl) {
  if (!pl) return NSAGP_OK;
  pl->arena.release();
  if (pl->ev[0]) cudaEventDestroy(pl->ev[0]);
  if (pl->ev[1]) cudaEventDestroy(pl->ev[1]);
  for (cudaEvent_t e : pl->phase_ev) cudaEventDestroy(e);
  delete pl;
  return NSAGP_OK;
}

}  // extern "C"

// ------------------------------------------------------------ run: dispatch
namespace {

// lb/ub = Eft -/+ 1.96 sqrt(Varft) (gf_ep_modulator_nmf.m:346-347).  IHGP: Varft is one M-vector
// (marginal variance of the last look-up) replicated over time with abs() (ihgp_ep_modulator_nmf.m:492-496).
__global__ void bounds_kernel(const double* __restrict__ E, const double* __restrict__ V, const double* __restrict__ vm0,
                              int M, long long total, double* __restrict__ Vout, double* __restrict__ lb,
                              double* __restrict__ ub) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const double v = vm0 ? fabs(vm0[i % M]) : V[i];
  const double sd = 1.96 * sqrt(v);
  if (Vout) Vout[i] = v;
  if (lb) lb[i] = E[i] - sd;
  if (ub) ub[i] = E[i] + sd;
}

// (variadic: the bodies contain commas, e.g. in <<<...>>>)
#define DISPATCH_BM(BMV, ...)                                             \
  switch (BMV) {                                                          \
    case 2: { constexpr int BM_ = 2; __VA_ARGS__; } break;                \
    case 3: { constexpr int BM_ = 3; __VA_ARGS__; } break;                \
    case 4: { constexpr int BM_ = 4; __VA_ARGS__; } break;                \
    case 6: { constexpr int BM_ = 6; __VA_ARGS__; } break;                \
    default: { constexpr int BM_ = 8; __VA_ARGS__; } break;               \
  }

#define DISPATCH_DP(DPV, ...)                                             \
  if ((DPV) == 16) { constexpr int DP_ = 16; __VA_ARGS__; } else { constexpr int DP_ = 32; __VA_ARGS__; }

struct PhaseTimer {
  nsagp_plan* pl;
  // pairs of events with a phase id
  std::vector<int> ids;
  size_t used = 0;
  int begin(int id) {
    if (used + 2 > pl->phase_ev.size()) {
      cudaEvent_t a, b;
      CU(cudaEventCreate(&a));
      CU(cudaEventCreate(&b));
      pl->phase_ev.push_back(a);
      pl->phase_ev.push_back(b);
    }
    ids.push_back(id);
    CU(cudaEventRecord(pl->phase_ev[used], g_stream));
    return NSAGP_OK;
  }
  int end() {
    CU(cudaEventRecord(pl->phase_ev[used + 1], g_stream));
    used += 2;
    return NSAGP_OK;
  }
  int collect() {
    for (int i = 1; i < 5; ++i) pl->timings[i] = 0.0;
    for (size_t i = 0; i < ids.size(); ++i) {
      float ms = 0.f;
      CU(cudaEventElapsedTime(&ms, pl->phase_ev[2 * i], pl->phase_ev[2 * i + 1]));
      pl->timings[ids[i]] += ms;
    }
    return NSAGP_OK;
  }
};

size_t lik_smem_bytes(const nsagp_plan* pl) {
  return ((size_t)pl->DP * kNP + (size_t)pl->S * (1 + kNP)) * sizeof(double);
}

// all problems of a plan share S; the link table is used only if every problem has one of the same size
size_t site_tab_bytes(const nsagp_plan* pl, int TPB) {
  int nd = pl->h_probs.empty() ? 0 : pl->h_probs[0].ndist;
  for (const DevProblem& P : pl->h_probs) nd = std::max(nd, P.ndist);
  return (size_t)site_tab_doubles(nd, pl->S, TPB) * sizeof(double);
}

int launch_sum(nsagp_plan* pl, int slot, int neg, long long k0 = 0, long long k1 = -1) {
  if (k1 < 0) k1 = pl->T;
  sum_kernel<<<pl->B, 1024, 0, g_stream>>>(pl->d_states, k0, k1 - k0, pl->d_nlZ, pl->ep_itts + 1, slot, neg);
  LAUNCH_CHECK();
  return NSAGP_OK;
}

int reset_diag(nsagp_plan* pl) {
  for (int b = 0; b < pl->B; ++b) CU(cudaMemsetAsync(pl->h_states[b].maxdiff, 0, 16, g_stream));
  return NSAGP_OK;
}

int save_diag(nsagp_plan* pl, int itt) {
  for (int b = 0; b < pl->B; ++b)
    CU(cudaMemcpyAsync(pl->d_diag + ((size_t)b * pl->ep_itts + itt) * 2, pl->h_states[b].maxdiff, 16,
                       cudaMemcpyDeviceToDevice, g_stream));
  return NSAGP_OK;
}

// Launch geometry of the CTA-cooperative sequential passes (adfcta.cuh).
struct AdfGeom {
  int threads, dpt;
  bool single;
  size_t smem;
  bool tab_smem;
};

static int sm_count() {
  static thread_local int dev_cached = -1, n_cached = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != dev_cached) {
    cudaDeviceGetAttribute(&n_cached, cudaDevAttrMultiProcessorCount, dev);
    dev_cached = dev;
  }
  return n_cached > 0 ? n_cached : 148;
}

// nblocks: CTAs of the launch (problems x chunks of the parallel first pass)
AdfGeom adf_geom(const nsagp_plan* pl, bool want_tables, bool fullstate, long long nblocks) {
  AdfGeom g;
  const int want = ((4 * pl->S + 31) / 32) * 32;
  int nmt = std::min(want, kAdfMaxMomThreads);            // moment threads
  // form 2 (batches of more problems than SMs): half the moment threads (each takes two sigma points per step) and the
  // steady-state tables left in HBM / L1, so that TWO CTAs fit an SM (2 x 192 threads x 164 registers) -- one problem's
  // Kalman section runs while the other's moment warps integrate
  // (C5, 256 problems on 148 SMs: 122 -> 88 ms, profiles/r3a_c5.jsonl; with at most one problem per SM the full-width CTA is
  // 1.3x faster, so form 0 picks by the size of the launch)
  const bool half = pl->adf_form == 2 || (pl->adf_form == 0 && nblocks > sm_count());
  if (half) { nmt = std::max(32, ((nmt / 2 + 31) / 32) * 32); want_tables = false; }
  g.threads = 32 + nmt;                                   // + the Kalman warp
  g.single = 4 * pl->S <= nmt;
  g.dpt = (pl->D <= 16) ? 4 : 8;
  const int NVP = (2 * g.dpt + 3 + 3) & ~3;
  const AdfSmem with(nmt / 32, NVP, pl->S, pl->M, pl->N, pl->nr, pl->BM, true, fullstate);
  const AdfSmem without(nmt / 32, NVP, pl->S, pl->M, pl->N, pl->nr, pl->BM, false, fullstate);
  g.tab_smem = want_tables && (size_t)with.total * 8 <= 200 * 1024;
  g.smem = (size_t)(g.tab_smem ? with.total : without.total) * 8;
  return g;
}

#define DISPATCH_DPT(DPTV, ...)                                           \
  if ((DPTV) == 4) { constexpr int DPT_ = 4; __VA_ARGS__; } else { constexpr int DPT_ = 8; __VA_ARGS__; }
#define DISPATCH_SINGLE(SV, ...)                                          \
  if (SV) { constexpr bool SINGLE_ = true; __VA_ARGS__; } else { constexpr bool SINGLE_ = false; __VA_ARGS__; }

// Chunk geometry of the parallel-in-time first pass over the window [w0, w1).
AdfPar adf_par(const nsagp_plan* pl, long long w0, long long w1, int& nch) {
  AdfPar par;
  nch = (int)std::max<long long>(1, std::min<long long>(pl->adf_chunks, (w1 - w0 + 63) / 64));    // >= 64 steps per chunk
  par.w0 = w0; par.w1 = w1; par.chunk_len = (w1 - w0 + nch - 1) / nch; par.burn = pl->adf_burn; par.bstate = pl->d_bstate;
  nch = (int)((w1 - w0 + par.chunk_len - 1) / par.chunk_len);
  return par;
}

int adf_mismatch(nsagp_plan* pl, const AdfPar& par, int nch) {
  if (nch < 2) return NSAGP_OK;
  adf_mismatch_kernel<<<pl->B, 128, 0, g_stream>>>(pl->d_probs, pl->d_states, par, nch, pl->d_adfdiag);
  LAUNCH_CHECK();
  return NSAGP_OK;
}

int ihgp_adf(nsagp_plan* pl, long long k0, long long k1, int mom_all, double damp, int running, bool parallel = false) {
  AdfPar par{0, 0, 0, 0, nullptr};
  int nch = 1;
  if (parallel && pl->adf_chunks > 1 && pl->adf_form != 1 && mom_all && !running) par = adf_par(pl, k0, k1, nch);
  if (pl->adf_form == 1) {
    const size_t sm = 64 * sizeof(double) + lik_smem_bytes(pl);
    DISPATCH_DP(pl->DP, DISPATCH_BM(pl->BM, {
      auto kern = ihgp_adf_kernel<DP_, BM_>;
      if (sm > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      kern<<<pl->B, 32, sm, g_stream>>>(pl->d_probs, pl->d_states, pl->T, k0, k1, mom_all, damp, running);
    }));
    LAUNCH_CHECK();
    return NSAGP_OK;
  }
  const AdfGeom g = adf_geom(pl, k1 - k0 > 64, false, (long long)pl->B * nch);      // the table copy only pays off on a long pass
  DISPATCH_DPT(g.dpt, DISPATCH_BM(pl->BM, DISPATCH_SINGLE(g.single, {
    if (g.tab_smem) {
      auto kern = ihgp_adf_cta_kernel<DPT_, BM_, SINGLE_, true>;
      if (g.smem > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
      kern<<<dim3(pl->B, nch), g.threads, g.smem, g_stream>>>(pl->d_probs, pl->d_states, pl->T, k0, k1, mom_all, damp, running, par);
    } else {
      auto kern = ihgp_adf_cta_kernel<DPT_, BM_, SINGLE_, false>;
      if (g.smem > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
      kern<<<dim3(pl->B, nch), g.threads, g.smem, g_stream>>>(pl->d_probs, pl->d_states, pl->T, k0, k1, mom_all, damp, running, par);
    }
  })));
  LAUNCH_CHECK();
  return adf_mismatch(pl, par, nch);
}

// Geometry of the three-phase scan (scan.cuh): CH chunks of kScanSteps steps per CTA tile, limited
// by the shared memory the tile needs for its chunk aggregates.
int scan_ch(const nsagp_plan* pl, int map_doubles, int max_threads, int lat) {
  const size_t per_chunk = (size_t)pl->M * map_doubles * sizeof(double);
  int ch = std::min(g_scan_ch_max, std::max(1, max_threads / lat));     // (latent, chunk) threads of a CTA tile
  while (ch > 1 && per_chunk * ch > 80 * 1024) --ch;         // phase 3 stages maps + states: < 2x this
  return ch;
}

// The two latent families of a scan (scan.cuh): subbands compute with bz, modulators with bg; storage stays padded to BM.
template <class EZ_, class EG_> struct ElemPair { using EZ = EZ_; using EG = EG_; };

// Specialised pairs for the block-size combinations of the reference's kernels that matter (exp / matern32 / matern52
// subbands with matern52 modulators); everything else computes both families at the padded size.
// (nsagp_scan_config: signals below g_fam_min_steps steps run both families padded.  With one launch per family the
// second launch and the small modulator-family CTAs cost more than the subband family saved on short signals -- C2 at
// T = 1e5 was 2 ms slower per EP run, hence a threshold of 400 000 steps; with both families in one CTA tile the
// specialised form wins at every length, profiles/r2l_ab.jsonl, and the default threshold is 0.)
#define DISPATCH_FAM(PL, ELEMT, ...)                                                                           \
  do {                                                                                                         \
    const bool fam_ = (PL)->T >= g_fam_min_steps;                                                              \
    const int bm_ = (PL)->BM, bz_ = fam_ ? (PL)->bz : -1, bg_ = (PL)->bg;                                      \
    if (bm_ == 3 && bz_ == 2 && bg_ == 3) { using Pair_ = ElemPair<ELEMT<3, 2>, ELEMT<3, 3>>; __VA_ARGS__; }    \
    else if (bm_ == 4 && bz_ == 4 && bg_ == 3) { using Pair_ = ElemPair<ELEMT<4, 4>, ELEMT<4, 3>>; __VA_ARGS__; } \
    else if (bm_ == 6 && bz_ == 6 && bg_ == 3) { using Pair_ = ElemPair<ELEMT<6, 6>, ELEMT<6, 3>>; __VA_ARGS__; } \
    else if (bm_ == 2) { using Pair_ = ElemPair<ELEMT<2, 2>, ELEMT<2, 2>>; __VA_ARGS__; }                       \
    else if (bm_ == 3) { using Pair_ = ElemPair<ELEMT<3, 3>, ELEMT<3, 3>>; __VA_ARGS__; }                       \
    else if (bm_ == 4) { using Pair_ = ElemPair<ELEMT<4, 4>, ELEMT<4, 4>>; __VA_ARGS__; }                       \
    else if (bm_ == 6) { using Pair_ = ElemPair<ELEMT<6, 6>, ELEMT<6, 6>>; __VA_ARGS__; }                       \
    else { using Pair_ = ElemPair<ELEMT<8, 8>, ELEMT<8, 8>>; __VA_ARGS__; }                                     \
  } while (0)

// Distinct element types per family: one CTA tile runs both (scan_reduce2 / scan_apply2_kernel, whole warps per family);
// nsagp_scan_merge(0) launches each family on its own instead (the form measured in profiles/r2i_*, kept for comparison).
static bool scan_merge() { return g_scan_merge != 0; }
// thread bound of the merged tile (scan.cuh: TH)
template <class Pair> static int scan2_bound() {
  return (ScanBounds<typename Pair::EZ>::kThreads == kScanThreads2 && g_scan_threads2 == 256) ? 256 : ScanBounds<typename Pair::EZ>::kThreads;
}
template <class Pair> static auto scan2_reduce_fn() {
  return scan2_bound<Pair>() == 256 ? scan_reduce2_kernel<typename Pair::EZ, typename Pair::EG, 256> : scan_reduce2_kernel<typename Pair::EZ, typename Pair::EG>;
}
template <class Pair> static auto scan2_apply_fn() {
  return scan2_bound<Pair>() == 256 ? scan_apply2_kernel<typename Pair::EZ, typename Pair::EG, 256> : scan_apply2_kernel<typename Pair::EZ, typename Pair::EG>;
}

template <class Pair>
int scan_setup(nsagp_plan* pl, ScanArgs& a, long long kfirst, long long nsteps, int dir, int init, long long kinit, int flags) {
  a.kfirst = kfirst; a.nsteps = nsteps; a.dir = dir; a.init = init; a.kinit = kinit; a.nprev = 0;
  a.flags = flags | (g_scan_l2pf ? kScanFlagL2Prefetch : 0);
  constexpr bool same = std::is_same<typename Pair::EZ, typename Pair::EG>::value;
  const bool merge = !same && scan_merge();
  const int maxthr = merge ? scan2_bound<Pair>() : ScanBounds<typename Pair::EZ>::kThreads;
  a.CH = scan_ch(pl, Pair::EZ::kMapDoubles, maxthr, same || merge ? pl->M : std::max(pl->D, pl->N));
  while (merge && a.CH > 1 && scan2_threads(pl->D, pl->N, a.CH) > maxthr) --a.CH;
  // a CTA tile of (family latents)*CH threads must fit an SM (registers, shared memory): ask the occupancy calculator
  // for both families' kernels -- they share the tile geometry, because they share the buffers
  auto fits = [&](int ch) {
    auto ok = [&](auto red, auto app, int lat, int threads) {
      const size_t sm1 = (size_t)ch * lat * Pair::EZ::kMapDoubles * sizeof(double);
      const size_t sm3 = (size_t)ch * lat * (Pair::EZ::kStateDoubles + Pair::EZ::kMapDoubles) * sizeof(double);
      if (sm1 > 48 * 1024) cudaFuncSetAttribute(red, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1);
      if (sm3 > 48 * 1024) cudaFuncSetAttribute(app, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm3);
      int n1 = 0, n3 = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n1, red, threads, sm1) != cudaSuccess) n1 = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n3, app, threads, sm3) != cudaSuccess) n3 = 0;
      cudaGetLastError();
      return n1 >= 1 && n3 >= 1;
    };
    if constexpr (same) {
      return ok(scan_reduce_kernel<typename Pair::EZ>, scan_apply_kernel<typename Pair::EZ>, pl->M, pl->M * ch);
    } else {
      if (merge)
        return ok(scan2_reduce_fn<Pair>(), scan2_apply_fn<Pair>(), pl->M, scan2_threads(pl->D, pl->N, ch));
      return ok(scan_reduce_kernel<typename Pair::EZ>, scan_apply_kernel<typename Pair::EZ>, pl->D, pl->D * ch) &&
             ok(scan_reduce_kernel<typename Pair::EG>, scan_apply_kernel<typename Pair::EG>, pl->N, pl->N * ch);
    }
  };
  while (a.CH > 1 && !fits(a.CH)) --a.CH;
  return NSAGP_OK;
}

// phase 1 (+ the shard aggregate on request: agg_host [M][W])
template <class Pair>
int scan_reduce(nsagp_plan* pl, const ScanArgs& a, double* agg_host) {
  const long long ntiles = scan_num_tiles(a.nsteps, a.CH);
  const dim3 grid((unsigned)ntiles, pl->B);
  auto launch = [&](auto kern, int n0, int cnt) -> int {
    if (cnt <= 0) return NSAGP_OK;
    const size_t sm1 = (size_t)a.CH * cnt * Pair::EZ::kMapDoubles * sizeof(double);
    if (sm1 > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1));
    kern<<<grid, cnt * a.CH, sm1, g_stream>>>(pl->d_probs, pl->d_states, a, pl->d_chunk, pl->d_tile, n0, cnt);
    LAUNCH_CHECK();
    return NSAGP_OK;
  };
  int rcl;
  if constexpr (std::is_same<typename Pair::EZ, typename Pair::EG>::value) {
    if ((rcl = launch(scan_reduce_kernel<typename Pair::EZ>, 0, pl->M))) return rcl;
  } else if (scan_merge()) {
    auto kern = scan2_reduce_fn<Pair>();
    const size_t sm1 = (size_t)a.CH * pl->M * Pair::EZ::kMapDoubles * sizeof(double);
    if (sm1 > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1));
    kern<<<grid, scan2_threads(pl->D, pl->N, a.CH), sm1, g_stream>>>(pl->d_probs, pl->d_states, a, pl->d_chunk, pl->d_tile);
    LAUNCH_CHECK();
  } else {
    if ((rcl = launch(scan_reduce_kernel<typename Pair::EZ>, 0, pl->D)) || (rcl = launch(scan_reduce_kernel<typename Pair::EG>, pl->D, pl->N))) return rcl;
  }
  if (agg_host) {
    scan_total_kernel<typename Pair::EZ, typename Pair::EG><<<1, 32, 0, g_stream>>>(pl->d_probs, pl->d_tile, ntiles, pl->d_total);
    LAUNCH_CHECK();
    CU(cudaMemcpyAsync(agg_host, pl->d_total, (size_t)pl->M * Pair::EZ::kMapDoubles * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
    CU(cudaStreamSynchronize(g_stream));
  }
  return NSAGP_OK;
}

// phases 2 and 3; prev_host: nprev aggregates [nprev][M][W] of the shards processed before this one
template <class Pair>
int scan_finish(nsagp_plan* pl, ScanArgs a, const double* prev_host, int nprev, bool prev_on_device = false) {
  const long long ntiles = scan_num_tiles(a.nsteps, a.CH);
  const dim3 grid((unsigned)ntiles, pl->B);
  if (nprev > 0) {
    if (pl->B != 1 || nprev > kMaxPrevShards) return fail(NSAGP_ERR_INVALID, "time-chunked scans need B = 1 and at most 16 shards");
    const size_t w = (size_t)pl->M * Pair::EZ::kMapDoubles;
    if (!prev_on_device)
      CU(cudaMemcpyAsync(pl->d_tile - (size_t)nprev * w, prev_host, (size_t)nprev * w * sizeof(double), cudaMemcpyHostToDevice, g_stream));
  }
  a.nprev = nprev;
  {
    const size_t per_tile = (size_t)pl->M * Pair::EZ::kMapDoubles * sizeof(double);
    const long long nlist = ntiles + nprev;                   // maps the walk passes through (per problem)
    auto carry = [&](const double* maps, double* starts, long long nmaps, long long count) -> int {
      int batch = (int)std::max<size_t>(1, std::min<size_t>((size_t)count, (64 * 1024) / per_tile));
      const size_t sm2 = per_tile * batch;
      if (sm2 > 48 * 1024) CU(cudaFuncSetAttribute(scan_carry_kernel<typename Pair::EZ, typename Pair::EG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
      scan_carry_kernel<typename Pair::EZ, typename Pair::EG><<<pl->B, kCarryThreads, sm2, g_stream>>>(pl->d_probs, pl->d_states, a, maps, starts, batch, nmaps);
      LAUNCH_CHECK();
      return NSAGP_OK;
    };
    if (nlist <= 96) {
      int rc = carry(pl->d_tile, pl->d_start, 0, nlist);
      if (rc) return rc;
    } else {
      // two-level carry (scan.cuh): segments of ~sqrt(nlist) tiles
      int seg_len = (int)std::ceil(std::sqrt((double)nlist));
      const long long nseg = (nlist + seg_len - 1) / seg_len;
      if (nseg > pl->nseg_max) return fail(NSAGP_ERR_INVALID, "internal: segment scratch too small");
      const double* maps = pl->d_tile - (size_t)nprev * pl->M * Pair::EZ::kMapDoubles;
      double* starts = pl->d_start - (size_t)nprev * pl->M * Pair::EZ::kStateDoubles;
      const int gy = std::max(1, 128 / pl->M);
      const dim3 sblock(pl->M, gy), sgrid((unsigned)((nseg + gy - 1) / gy), pl->B);
      carry_seg_reduce_kernel<typename Pair::EZ, typename Pair::EG><<<sgrid, sblock, 0, g_stream>>>(pl->d_probs, maps, pl->d_seg, nlist, seg_len);
      LAUNCH_CHECK();
      int rc = carry(pl->d_seg, pl->d_segstart, nseg, nseg);
      if (rc) return rc;
      carry_seg_apply_kernel<typename Pair::EZ, typename Pair::EG><<<sgrid, sblock, 0, g_stream>>>(pl->d_probs, maps, pl->d_segstart, starts, nlist, seg_len);
      LAUNCH_CHECK();
    }
  }
  {
    auto launch = [&](auto kern, int n0, int cnt) -> int {
      if (cnt <= 0) return NSAGP_OK;
      const size_t sm3 = (size_t)a.CH * cnt * (Pair::EZ::kStateDoubles + Pair::EZ::kMapDoubles) * sizeof(double);
      if (sm3 > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm3));
      kern<<<grid, cnt * a.CH, sm3, g_stream>>>(pl->d_probs, pl->d_states, a, pl->d_chunk, pl->d_start, n0, cnt);
      LAUNCH_CHECK();
      return NSAGP_OK;
    };
    int rcl;
    if constexpr (std::is_same<typename Pair::EZ, typename Pair::EG>::value) {
      if ((rcl = launch(scan_apply_kernel<typename Pair::EZ>, 0, pl->M))) return rcl;
    } else if (scan_merge()) {
      auto kern = scan2_apply_fn<Pair>();
      const size_t sm3 = (size_t)a.CH * pl->M * (Pair::EZ::kStateDoubles + Pair::EZ::kMapDoubles) * sizeof(double);
      if (sm3 > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm3));
      kern<<<grid, scan2_threads(pl->D, pl->N, a.CH), sm3, g_stream>>>(pl->d_probs, pl->d_states, a, pl->d_chunk, pl->d_start);
      LAUNCH_CHECK();
    } else {
      if ((rcl = launch(scan_apply_kernel<typename Pair::EZ>, 0, pl->D)) || (rcl = launch(scan_apply_kernel<typename Pair::EG>, pl->D, pl->N))) return rcl;
    }
  }
  return NSAGP_OK;
}

template <class Pair>
int run_scan(nsagp_plan* pl, long long kfirst, long long nsteps, int dir, int init, long long kinit, int flags = 0) {
  if (nsteps <= 0) return NSAGP_OK;
  ScanArgs a;
  int rc = scan_setup<Pair>(pl, a, kfirst, nsteps, dir, init, kinit, flags);
  if (rc || (rc = scan_reduce<Pair>(pl, a, nullptr))) return rc;
  return scan_finish<Pair>(pl, a, nullptr, 0);
}

template <template <int, int> class ElemT>
int affine_scan(nsagp_plan* pl, long long kfirst, long long nsteps, int dir, int init, long long kinit, int flags = 0) {
  int rc = NSAGP_OK;
  DISPATCH_FAM(pl, ElemT, { rc = run_scan<Pair_>(pl, kfirst, nsteps, dir, init, kinit, flags); });
  return rc;
}

// Smoother-side site update over steps [k0, k1): four lanes per step (siteupd.cuh), sigma points two at a time when every
// problem of the plan has the link table of distinct coordinates; NSAGP_SITE_FORM=2 forces the one-point-at-a-time form,
// NSAGP_SITE_FORM=1 selects the first-generation one-thread-per-step kernel (both kept as independent cross-checks).
template <bool FULL>
int site_update_launch(nsagp_plan* pl, double damp, int write_lZ, int clamp_R, long long k0, long long k1) {
  const int form = g_site_form;
  if (form == 1) {
    constexpr int TPB = 64;
    const size_t sm = (size_t)4 * pl->M * TPB * sizeof(double) + lik_smem_bytes(pl) + site_tab_bytes(pl, TPB);
    const dim3 grid((unsigned)((k1 - k0 + TPB - 1) / TPB), pl->B);
    DISPATCH_DP(pl->DP, {
      auto kern = site_update_kernel<DP_, TPB, FULL>;
      CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      kern<<<grid, TPB, sm, g_stream>>>(pl->d_probs, pl->d_states, k0, k1, pl->alpha, damp, write_lZ, clamp_R);
    });
    LAUNCH_CHECK();
    return NSAGP_OK;
  }
  int nd = 0, nd_min = 1 << 30;
  for (const DevProblem& P : pl->h_probs) { nd = std::max(nd, P.ndist); nd_min = std::min(nd_min, P.ndist); }
  const bool pair = form != 2 && nd_min > 0;
  const size_t sm = (size_t)site4_smem_doubles(pl->M, pl->S, nd, FULL) * sizeof(double);
  const dim3 grid((unsigned)((k1 - k0 + kSiteSteps - 1) / kSiteSteps), pl->B);
  DISPATCH_DPT((pl->D <= 16) ? 4 : 8, {
    auto kern = pair ? site_update4_kernel<DPT_, FULL, true> : site_update4_kernel<DPT_, FULL, false>;
    if (sm > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    kern<<<grid, kSiteThreads, sm, g_stream>>>(pl->d_probs, pl->d_states, k0, k1, pl->alpha, damp, write_lZ, clamp_R);
  });
  LAUNCH_CHECK();
  return NSAGP_OK;
}

int ihgp_site_update(nsagp_plan* pl, double damp, int write_lZ, long long k0 = 0, long long k1 = -1) {
  if (k1 < 0) k1 = pl->T - 1;
  if (k1 <= k0) return NSAGP_OK;
  return site_update_launch<false>(pl, damp, write_lZ, 0, k0, k1);
}

int run_ihgp(nsagp_plan* pl, PhaseTimer& tm) {
  const long long T = pl->T;
  int rc;
  if (pl->mode != NSAGP_MODE_PREDICT) {
    // nlZ mode: a single ADF sweep (ihgp_ep_modulator_nmf.m:533-624)
    if ((rc = tm.begin(1))) return rc;
    if ((rc = ihgp_adf(pl, 0, T, 1, pl->damping[0], pl->mode == NSAGP_MODE_NLZ_RUNNING, true))) return rc;
    if ((rc = tm.end())) return rc;
    return launch_sum(pl, pl->ep_itts, 1);
  }
  double damp = pl->damping[0];
  for (int itt = 1; itt <= pl->ep_itts; ++itt) {
    if ((rc = reset_diag(pl))) return rc;
    if (itt == 1) {
      if ((rc = tm.begin(1)) || (rc = ihgp_adf(pl, 0, T, 1, damp, 0, true)) || (rc = tm.end())) return rc;
      if ((rc = launch_sum(pl, 0, 1))) return rc;
    } else {
      if ((rc = tm.begin(2))) return rc;
      if ((rc = affine_scan<FilterElem>(pl, 0, T - 1, +1, 0, 0))) return rc;
      if ((rc = ihgp_adf(pl, T - 1, T, 0, damp, 0))) return rc;     // last step: moments + update (:253)
      if ((rc = tm.end())) return rc;
    }
    if (itt == pl->ep_itts && pl->d_MF) {
      for (int b = 0; b < pl->B; ++b)
        CU(cudaMemcpyAsync(pl->d_MF + (size_t)b * T * pl->n, pl->h_states[b].MS, (size_t)T * pl->n * sizeof(double),
                           cudaMemcpyDeviceToDevice, g_stream));
    }
    if (itt < pl->ep_itts) damp = pl->damping[itt];                 // ep_damping(itt+1) (:369-371)
    if ((rc = tm.begin(3))) return rc;
    if ((rc = affine_scan<SmootherElem>(pl, T - 2, T - 1, -1, 1, T - 1))) return rc;
    if ((rc = tm.end())) return rc;
    if (itt < pl->ep_itts) {
      if ((rc = tm.begin(4)) || (rc = ihgp_site_update(pl, damp, itt > 1)) || (rc = tm.end())) return rc;
      if ((rc = launch_sum(pl, itt, 1))) return rc;
      DISPATCH_BM(pl->BM, { carry_mean_kernel<BM_><<<pl->B, 32, 0, g_stream>>>(pl->d_probs, pl->d_states); });
      LAUNCH_CHECK();
    }
    if ((rc = save_diag(pl, itt - 1))) return rc;
  }
  return NSAGP_OK;
}

// (re)initialise the mutable state so a plan can be run repeatedly
int plan_reset(nsagp_plan* pl) {
  const long long T = pl->T;
  for (int b = 0; b < pl->B; ++b) {
    DevState& St = pl->h_states[b];
    CU(cudaMemsetAsync(St.ttau, 0, (size_t)T * pl->M * 8, g_stream));
    CU(cudaMemsetAsync(St.tnu, 0, (size_t)T * pl->M * 8, g_stream));
    CU(cudaMemsetAsync(St.R, 0, (size_t)T * pl->M * 8, g_stream));
    CU(cudaMemsetAsync(St.MS, 0, (size_t)T * pl->n * 8, g_stream));
    CU(cudaMemsetAsync(St.E, 0, (size_t)T * pl->M * 8, g_stream));
    CU(cudaMemsetAsync(St.lZ, 0, (size_t)T * 8, g_stream));
    CU(cudaMemsetAsync(St.mcarry, 0, (size_t)pl->M * pl->BM * 8, g_stream));
    CU(cudaMemsetAsync(St.maxdiff, 0, 16, g_stream));
    CU(cudaMemsetAsync(St.negcav, 0, 8, g_stream));
    CU(cudaMemsetAsync(St.status, 0, 4, g_stream));
    if (St.V) CU(cudaMemsetAsync(St.V, 0, (size_t)T * pl->M * 8, g_stream));
    if (pl->kind == 0 && T == 1) {
      // the reference resets P to zeros before the smoother loop (ihgp_ep_modulator_nmf.m:358), and with a
      // single time step that loop never runs: Varft = diag(H*0*H') = 0
      CU(cudaMemsetAsync(St.vm0, 0, pl->M * 8, g_stream));
    } else {
      CU(cudaMemcpyAsync(St.vm0, pl->h_vminf.data() + (size_t)b * pl->M, pl->M * 8, cudaMemcpyHostToDevice, g_stream));
    }
  }
  CU(cudaMemsetAsync(pl->d_nlZ, 0, (size_t)pl->B * (pl->ep_itts + 1) * 8, g_stream));
  CU(cudaMemsetAsync(pl->d_diag, 0, (size_t)pl->B * pl->ep_itts * 16, g_stream));
  if (pl->d_adfdiag) CU(cudaMemsetAsync(pl->d_adfdiag, 0, 16, g_stream));
  return NSAGP_OK;
}

}  // namespace

#include "api_full.inc"
#include "api_ekf.inc"
#include "api_chunk.inc"
#include "api_comm.inc"
#include "api_fb.inc"
#include "api_tables.inc"
#include "api_mc.inc"

extern "C" {

int nsagp_plan_run(nsagp_plan* pl) {
  if (!pl) return fail(NSAGP_ERR_INVALID, "null plan");
  int rc = ensure_stream();
  if (rc) return rc;
  if ((rc = plan_reset(pl))) return rc;
  PhaseTimer tm{pl};
  CU(cudaEventRecord(pl->ev[0], g_stream));
  rc = (pl->kind == 0) ? run_ihgp(pl, tm) : run_full(pl, tm);
  if (rc) return rc;
  CU(cudaEventRecord(pl->ev[1], g_stream));
  CU(cudaStreamSynchronize(g_stream));
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, pl->ev[0], pl->ev[1]));
  pl->timings[0] = ms;
  if ((rc = tm.collect())) return rc;
  pl->ran = true;
  // sticky device-side status (full-state path: Cholesky failure / non-positive variance)
  for (int b = 0; b < pl->B; ++b) {
    int st = 0;
    CU(cudaMemcpy(&st, pl->h_states[b].status, 4, cudaMemcpyDeviceToHost));
    if (st == 1) return fail(NSAGP_ERR_NOT_PD, "smoother Cholesky failed (predicted covariance not positive definite)");
    if (st == 2) return fail(NSAGP_ERR_NONPOS_VAR, "non-positive predictive variance in nlZ mode");
  }
  return NSAGP_OK;
}

int nsagp_plan_timings(nsagp_plan* pl, double* ms, int32_t n) {
  if (!pl || !ms) return 0;
  const int k = std::min<int>(n, 5);
  for (int i = 0; i < k; ++i) ms[i] = pl->timings[i];
  return k;
}

int nsagp_plan_fetch(nsagp_plan* pl, int32_t b, nsagp_outputs* o) {
  if (!pl || !o) return fail(NSAGP_ERR_INVALID, "null argument");
  if (!pl->ran) return fail(NSAGP_ERR_INVALID, "plan has not been run");
  if (b < 0 || b >= pl->B) return fail(NSAGP_ERR_INVALID, "problem index out of range");
  const DevState& St = pl->h_states[b];
  const long long T = pl->T;
  const int M = pl->M;
  const size_t TM = (size_t)T * M * 8;
  auto get = [&](void* dst, const void* src, size_t nb) -> int {
    if (!dst) return NSAGP_OK;
    if (!src) return fail(NSAGP_ERR_INVALID, "requested output is not produced in this mode");
    CU(cudaMemcpyAsync(dst, src, nb, cudaMemcpyDeviceToHost, g_stream));
    return NSAGP_OK;
  };
  int rc;
  const bool predict = pl->mode == NSAGP_MODE_PREDICT;
  if ((rc = get(o->ttau, St.ttau, TM)) || (rc = get(o->tnu, St.tnu, TM)) || (rc = get(o->R, St.R, TM)) ||
      (rc = get(o->lZ, St.lZ, (size_t)T * 8)))
    return rc;
  if (predict) {
    if ((rc = get(o->Eft, St.E, TM)) || (rc = get(o->MS, St.MS, (size_t)T * pl->n * 8)) ||
        (rc = get(o->MF, pl->d_MF ? pl->d_MF + (size_t)b * T * pl->n : nullptr, (size_t)T * pl->n * 8)))
      return rc;
    if (pl->kind == 1) {
      if ((rc = get(o->Varft, St.V, TM))) return rc;
      if ((rc = fetch_full_cov(pl, b, o))) return rc;
    }
    if ((rc = get(o->nlZ, pl->d_nlZ + (size_t)b * (pl->ep_itts + 1), (size_t)pl->ep_itts * 8))) return rc;
  } else {
    if (o->Eft || o->Varft || o->lb || o->ub || o->MS || o->MF || o->PF || o->PS)
      return fail(NSAGP_ERR_INVALID, "posterior outputs are only available in predict mode");
  }
  if ((rc = get(o->edata, pl->d_nlZ + (size_t)b * (pl->ep_itts + 1) + pl->ep_itts, 8))) return rc;
  std::vector<double> diag;
  if (o->maxDiffM || o->maxDiffP) {
    diag.resize((size_t)pl->ep_itts * 2);
    CU(cudaMemcpyAsync(diag.data(), pl->d_diag + (size_t)b * pl->ep_itts * 2, diag.size() * 8, cudaMemcpyDeviceToHost, g_stream));
  }
  unsigned long long neg = 0;
  if (o->n_negcav) CU(cudaMemcpyAsync(&neg, St.negcav, 8, cudaMemcpyDeviceToHost, g_stream));
  if (predict && ((pl->kind == 0 && o->Varft) || o->lb || o->ub)) {
    const long long total = (long long)T * M;
    if (!pl->d_bounds) { int rc2 = pl->arena.alloc(&pl->d_bounds, (size_t)3 * total); if (rc2) return rc2; }
    double* dV = pl->d_bounds; double* dlb = dV + total; double* dub = dlb + total;
    bounds_kernel<<<(unsigned)((total + 255) / 256), 256, 0, g_stream>>>(St.E, St.V, pl->kind == 0 ? St.vm0 : nullptr, M, total,
                                                                        pl->kind == 0 ? dV : nullptr, o->lb ? dlb : nullptr,
                                                                        o->ub ? dub : nullptr);
    LAUNCH_CHECK();
    if (pl->kind == 0 && o->Varft) CU(cudaMemcpyAsync(o->Varft, dV, TM, cudaMemcpyDeviceToHost, g_stream));
    if (o->lb) CU(cudaMemcpyAsync(o->lb, dlb, TM, cudaMemcpyDeviceToHost, g_stream));
    if (o->ub) CU(cudaMemcpyAsync(o->ub, dub, TM, cudaMemcpyDeviceToHost, g_stream));
  }
  CU(cudaStreamSynchronize(g_stream));
  if (o->n_negcav) *o->n_negcav = (int64_t)neg;
  for (int i = 0; i < pl->ep_itts && !diag.empty(); ++i) {
    if (o->maxDiffM) o->maxDiffM[i] = diag[2 * i];
    if (o->maxDiffP) o->maxDiffP[i] = diag[2 * i + 1];
  }
  return NSAGP_OK;
}

static int run_hostbuf(int kind, int32_t B, const nsagp_model* models, const nsagp_lik* liks, const nsagp_ep* ep,
                       const nsagp_tables* tables, const double* y, int64_t T, int32_t mode, nsagp_outputs* outs) {
  if (!outs) return fail(NSAGP_ERR_INVALID, "null outputs");
  nsagp_plan* pl = nullptr;
  int rc = nsagp_plan_create(&pl, kind, B, models, liks, ep, tables, y, T, mode);
  if (rc) return rc;
  for (int b = 0; b < B && !rc; ++b)
    if (outs[b].PF) rc = nsagp_plan_keep_pf(pl, 1);
  if (!rc) rc = nsagp_plan_run(pl);
  for (int b = 0; b < B && !rc; ++b) rc = nsagp_plan_fetch(pl, b, &outs[b]);
  nsagp_plan_destroy(pl);
  return rc;
}

int nsagp_giekf(const nsagp_model* model, const double* W, double sigma2, int32_t g_iter, int32_t l_iter,
                const double* y, int64_t T, int32_t mode, nsagp_outputs* out) {
  return giekf_impl(model, W, sigma2, g_iter, l_iter, 0, y, T, mode, out);
}

int nsagp_giekf_carry(const nsagp_model* model, const double* W, double sigma2, int32_t g_iter, int32_t l_iter,
                      const double* y, int64_t T, int32_t mode, nsagp_outputs* out) {
  return giekf_impl(model, W, sigma2, g_iter, l_iter, 1, y, T, mode, out);
}

int nsagp_giekf_grad(const nsagp_model* model, const double* W, double sigma2, int32_t nparam, const int32_t* latent,
                     const double* dA, const double* dQ, const double* dPinf, const double* dR, const double* y, int64_t T,
                     double* edata, double* gdata) {
  return giekf_grad_impl(model, W, sigma2, nparam, latent, dA, dQ, dPinf, dR, y, T, edata, gdata);
}

int nsagp_mc_reconstruct(int32_t D, int32_t N, int64_t T, int32_t s, const double* Eft, const double* Varft, const double* W,
                         double link_shift, int32_t sqrt_model, const double* Z, uint64_t seed, double* Esig, double* Vsig,
                         double* Eft_mod, double* Varft_mod) {
  return mc_reconstruct_impl(D, N, T, s, Eft, Varft, W, link_shift, sqrt_model, Z, seed, Esig, Vsig, Eft_mod, Varft_mod);
}

int nsagp_giekf_config(int32_t smoother_form, int32_t chunk_len, int32_t chunks_per_segment) {
  if (smoother_form < 0 || smoother_form > 2 || chunk_len < 0 || chunks_per_segment < 0)
    return fail(NSAGP_ERR_INVALID, "smoother_form must be 0, 1 or 2; chunk_len, chunks_per_segment >= 0");
  g_giekf_cfg.form = smoother_form;
  g_giekf_cfg.chunk_len = chunk_len;
  g_giekf_cfg.seg_chunks = chunks_per_segment;
  return NSAGP_OK;
}

int nsagp_giekf_timings(double* ms, int32_t n) {
  if (!ms || n < 2) return fail(NSAGP_ERR_INVALID, "need room for 2 values");
  ms[0] = g_giekf_ms[0]; ms[1] = g_giekf_ms[1];
  return NSAGP_OK;
}

int nsagp_ep_ihgp(const nsagp_model* model, const nsagp_lik* lik, const nsagp_ep* ep, const nsagp_tables* tables,
                  const double* y, int64_t T, int32_t mode, nsagp_outputs* out) {
  return run_hostbuf(0, 1, model, lik, ep, tables, y, T, mode, out);
}

int nsagp_ep_ihgp_batch(int32_t B, const nsagp_model* models, const nsagp_lik* liks, const nsagp_ep* ep,
                        const nsagp_tables* tables, const double* y, int64_t T, int32_t mode, nsagp_outputs* outs) {
  return run_hostbuf(0, B, models, liks, ep, tables, y, T, mode, outs);
}

int nsagp_ep_full(const nsagp_model* model, const nsagp_lik* lik, const nsagp_ep* ep, const double* y, int64_t T,
                  int32_t mode, nsagp_outputs* out) {
  return run_hostbuf(1, 1, model, lik, ep, nullptr, y, T, mode, out);
}

int nsagp_ep_full_batch(int32_t B, const nsagp_model* models, const nsagp_lik* liks, const nsagp_ep* ep,
                        const double* y, int64_t T, int32_t mode, nsagp_outputs* outs) {
  return run_hostbuf(1, B, models, liks, ep, nullptr, y, T, mode, outs);
}

}  // extern "C"
