/*
 * nsagp.h -- C ABI of libnsagp.so: the B200 (sm_100a) implementation of the
 * EP-in-Kalman inference hot path of AaltoML/nonstationary-audio-gp.
 *
 * This is the drop-in boundary.  The reference has no FFI (it is 100 % MATLAB);
 * these entry points are what a MEX gateway for its entry functions binds
 * (INTEGRATION.md shows the gateway and the .m wrappers).  Each function cites
 * the reference interface it replaces.  Conventions:
 *   - all floating point data is IEEE FP64, matrices are column-major
 *     (MATLAB layout), so M-by-T site arrays are contiguous per time step;
 *   - every buffer is caller-owned HOST memory unless the name says "dev";
 *     the library copies to/from the GPU itself and allocates nothing the
 *     caller can see;
 *   - NaN in y marks a missing / test-only sample (gf_ep_modulator_nmf.m:59);
 *   - functions return NSAGP_OK or a negative nsagp_status; nsagp_last_error()
 *     gives the message.  There is no CPU fallback: without a CUDA device every
 *     compute entry point returns NSAGP_ERR_CUDA.
 */
#ifndef NSAGP_H
#define NSAGP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum nsagp_status {
  NSAGP_OK = 0,
  NSAGP_ERR_INVALID = -1,     /* bad argument / unsupported shape */
  NSAGP_ERR_CUDA = -2,        /* CUDA runtime failure (message has the detail) */
  NSAGP_ERR_NOT_PD = -3,      /* smoother Cholesky failed; the reference would add
                                 rand() jitter here (gf_ep_modulator_nmf.m:216-223) */
  NSAGP_ERR_NONPOS_VAR = -4,  /* predictive variance <= 0; the reference drops into
                                 `keyboard` (gf_ep_modulator_nmf.m:408-410) */
  NSAGP_ERR_NAN = -5          /* NaN energy (gf_giekf_modulator_nmf_constraints.m:423-426) */
} nsagp_status;

/* Discrete-time model in block form.  Replaces the dense (A,Q,H,Pinf) the
 * reference builds with ss(...) + lti_disc(...) (gf_ep_modulator_nmf.m:78,108;
 * ihgp_ep_modulator_nmf.m:79-97): those matrices are block diagonal with one
 * block and one observation row per latent (ss_modulators_nmf.m:128-132).
 * D subband blocks of size bz come first, then N modulator blocks of size bg.
 * A, Q, Pinf: packed b-by-b blocks, column-major, D*bz*bz + N*bg*bg doubles.
 * h: packed observation rows, D*bz + N*bg doubles. */
typedef struct nsagp_model {
  int32_t D, N;
  int32_t bz, bg;             /* 1..8 each */
  const double* A;
  const double* Q;
  const double* Pinf;
  const double* h;
} nsagp_model;

/* Likelihood / moment-matching closure.  Replaces the `mom` function handle
 * (demo_toy_modulators_nmf.m:81, experiments/train_model.m:190), resolved to
 *   kind 0: likModulatorNMFPower.m:28-87     a(g) = W*link(g),       pEP_const = 1
 *   kind 1: experiments/likModulatorPreCalcwn.m:28-86
 *                                            a(g) = sqrt(W*link(g)), pEP_const = (2 pi sn2)^((1-alpha)/2) alpha^(-1/2)
 * with link(g) = log(1+exp(g - link_shift)).  (wn, xn) are the unit sigma points
 * of utp_ws(p,N) or mvhermgauss: wn[S], xn N-by-S column-major. */
typedef struct nsagp_lik {
  int32_t kind;
  double sn2;                 /* exp(lik_param) */
  double link_shift;
  const double* W;            /* D-by-N, column-major */
  int32_t S;
  const double* wn;
  const double* xn;
} nsagp_lik;

/* EP schedule: ep_fraction (Power-EP alpha), ep_damping[ep_itts], ep_itts
 * (arguments 12-14 of gf_ep_modulator_nmf.m:1). */
typedef struct nsagp_ep {
  double ep_fraction;
  const double* ep_damping;
  int32_t ep_itts;
} nsagp_ep;

/* Steady-state tables of the infinite-horizon path
 * (ihgp_ep_modulator_nmf.m:131-133 PPlist, :183-189 PGlist), built on the host.
 * r[nr] ascending grid of equivalent noise variances.  PP: per block nr rows of
 * b*b doubles (column-major b-by-b), blocks concatenated.  PG: per block nr rows
 * of 2*b*b doubles [smoothed covariance, smoother gain]; may be NULL in nlZ mode. */
typedef struct nsagp_tables {
  int32_t nr;
  const double* r;
  const double* PP;
  const double* PG;
} nsagp_tables;

typedef enum nsagp_mode {
  NSAGP_MODE_PREDICT = 0,     /* xt non-empty: posterior marginals            */
  NSAGP_MODE_NLZ = 1,         /* xt empty: negative log marginal likelihood   */
  NSAGP_MODE_NLZ_RUNNING = 2  /* ihgp ..._constraints nlZ mode: site vectors carried
                                 from step to step (ihgp_ep_modulator_nmf_constraints.m:568-615) */
} nsagp_mode;

/* Output buffers (any pointer may be NULL = not wanted).  Shapes as in the
 * reference's outputs: M = D+N sites, n = state dimension, T steps.
 *   Eft, Varft, lb, ub : M-by-T     (gf_ep_modulator_nmf.m:321-348)
 *   ttau, tnu, R       : M-by-T     (out.ttau, out.tnu, out.R)
 *   lZ                 : T          (out.lZ; ihgp: per-step terms of the last pass)
 *   MF, MS             : n-by-T     (out.MF filtered, out.MS smoothed means)
 *   PF, PS             : packed blocks, (D*bz*bz+N*bg*bg)-by-T (block form of out.PF/out.PS)
 *   nlZ                : ep_itts    (nlZ(itt) trace printed by the reference)
 *   maxDiffM, maxDiffP : ep_itts    (convergence diagnostics, :271-272)
 *   edata              : 1          (nlZ mode: -sum(lZ), :525)
 *   n_negcav           : 1          (count of cavity variances <= 0; the reference
 *                                    silently goes complex there) */
typedef struct nsagp_outputs {
  double* Eft; double* Varft; double* lb; double* ub;
  double* ttau; double* tnu; double* R; double* lZ;
  double* MF; double* MS; double* PF; double* PS;
  double* nlZ; double* maxDiffM; double* maxDiffP;
  double* edata;
  int64_t* n_negcav;
} nsagp_outputs;

const char* nsagp_version(void);
const char* nsagp_last_error(void);

/* Select the CUDA device used by the calling thread's subsequent calls and the
 * stream the kernels are launched on (0 = the library's own stream). */
int nsagp_set_device(int device);
int nsagp_set_stream(void* cuda_stream);
int nsagp_device_count(void);

/* Device buffers of finished calls are kept for the next call of the same shapes (the entry
 * points are called thousands of times by fminunc); this returns them to the driver. */
int nsagp_release_cache(void);

/* Number of kernel launches issued by this library since the last reset
 * (bench.py reports it as gpu_launches). */
int64_t nsagp_launch_count(int reset);

/* likModulatorNMFPower / likModulatorPreCalcwn for a batch of independent
 * calls: for k < T, [lZ(k), dlZ(:,k), d2lZ(:,k)] = mom(y(k), mu(:,k), s2(:,k), ep_fraction)
 * (likModulatorNMFPower.m:28-87).  mu, s2, dlZ, d2lZ are M-by-T. */
int nsagp_mom_batch(const nsagp_lik* lik, int32_t D, int32_t N, double ep_fraction,
                    int64_t T, const double* y, const double* mu, const double* s2,
                    double* lZ, double* dlZ, double* d2lZ);
/* Same contract, computed by the warp-cooperative form of the moment kernel (the
 * one the sequential ADF pass uses); exposed so parity tests reach it directly. */
int nsagp_mom_batch_warp(const nsagp_lik* lik, int32_t D, int32_t N, double ep_fraction,
                         int64_t T, const double* y, const double* mu, const double* s2,
                         double* lZ, double* dlZ, double* d2lZ);

/* Test hook: the straight-line FP64 routines used on the critical path of the
 * sequential passes (csrc/fastmath.cuh), evaluated for n host values.
 * op: 0 1/x, 1 1/sqrt(x), 2 sqrt(x), 3 exp(x), 4 log(x) for x >= 1, 5 log(1+exp(x)), 6-8 the variants of 0-2 with
 * one Newton step fewer, 9 the table-driven log(x) for x >= 1, 10 log(1+exp(x)) built on it (finite x). */
int nsagp_fastmath_eval(int32_t op, int64_t n, const double* x, double* out);

/* Steady-state tables of the infinite-horizon path built natively on the host
 * (ihgp_ep_modulator_nmf.m:106-134 PPlist, :150-191 PGlist): per latent block, n_coarse Riccati solutions on
 * ro = logspace(log10_lo, log10_hi, n_coarse) -- the reference calls dare(A',H',Q,ro) and dare(G',0,QQ); here a
 * doubling algorithm in long double -- interpolated piecewise linearly in r onto the n_fine-point grid
 * (apxGrid's non-equispaced branch).  The reference uses (32, 200, -2, 4).  Outputs in the nsagp_tables layout:
 * r[n_fine], PP (sum over blocks n_fine*b*b), PG (sum over blocks n_fine*2*b*b; may be NULL if !want_smoother).
 * model.Q must already be symmetrised (:97).  No GPU is needed for this call. */
int nsagp_ihgp_tables(const nsagp_model* model, int32_t want_smoother, int32_t n_coarse, int32_t n_fine,
                      double log10_lo, double log10_hi, double* r, double* PP, double* PG);

/* ihgp_ep_modulator_nmf / ihgp_ep_modulator_nmf_constraints
 * (ihgp_ep_modulator_nmf.m:195-526 predict, :533-624 nlZ).  y[T] is `yall` after
 * the merge/sort of train and test inputs (:58-67). */
int nsagp_ep_ihgp(const nsagp_model* model, const nsagp_lik* lik, const nsagp_ep* ep,
                  const nsagp_tables* tables, const double* y, int64_t T,
                  int32_t mode, nsagp_outputs* out);

/* gf_ep_modulator_nmf / gf_ep_modulator_nmf_constraints
 * (gf_ep_modulator_nmf.m:92-352 predict, :357-533 nlZ). */
int nsagp_ep_full(const nsagp_model* model, const nsagp_lik* lik, const nsagp_ep* ep,
                  const double* y, int64_t T, int32_t mode, nsagp_outputs* out);

/* gf_giekf_modulator_nmf_constraints (gf_giekf_modulator_nmf_constraints.m:144-327 predict mode,
 * :332-484 energy of the nlZ mode as every caller runs it, GradObj = 'off'): globally iterated
 * extended Kalman filter + RTS smoother with the hard-wired measurement y = (H_z x)' W softplus(H_g x)
 * (:490-513, iekf_update1.m:110-117).  The covariance is dense here.  W is D-by-N column-major,
 * sigma2 = exp(lik_param).  Predict mode: model.A/Q from lti_disc; energy mode (NSAGP_MODE_NLZ):
 * model.A = expm(F), model.Q = Pinf - A Pinf A' (:376-380), result in out->edata (NaN and
 * NSAGP_ERR_NAN when the innovation variance is not positive, :417-427).  Outputs used: Eft, Varft,
 * lb, ub, MF, MS, maxDiffP[g_iter] (over the marginal variances), PF/PS as DENSE n-by-n-by-T. */
int nsagp_giekf(const nsagp_model* model, const double* W, double sigma2, int32_t g_iter, int32_t l_iter,
                const double* y, int64_t T, int32_t mode, nsagp_outputs* out);

/* gf_giekf_modulator_nmf (the file without constraints; callers: experiments/missing_data_music.m:128,
 * noise_reduction_speech.m:97, synthetic_data_experiment.m:176): as nsagp_giekf, but the state (m, P) is
 * initialised on the first global iteration only (gf_giekf_modulator_nmf.m:127-131), so the smoothed mean AND
 * covariance of the first time step start the next filter pass. */
int nsagp_giekf_carry(const nsagp_model* model, const double* W, double sigma2, int32_t g_iter, int32_t l_iter,
                      const double* y, int64_t T, int32_t mode, nsagp_outputs* out);

/* gf_giekf_modulator_nmf with GradObj = 'on' and no test inputs (gf_giekf_modulator_nmf.m:296-437): the EKF energy
 * and its gradient by the sensitivity equations, one (dm_j, dP_j) recursion per hyper-parameter j.
 * model: A = expm(F), Q = Pinf - A Pinf A' (:355-357).  Every parameter lives in ONE latent (ss_modulators_nmf.m:25-129)
 * or in none (the noise variance, j = 0 in the reference): latent[j] in -1 .. D+N-1, and for that latent's block
 *   dA[j]    = lower-left block of expm([F 0; dF_j F])       (:328-338)
 *   dQ[j]    = dPinf_j - dA Pinf A' - A dPinf_j A' - (dA Pinf A')'   (:362-364)
 *   dPinf[j] = the initial dP_j                               (:314)
 * each a bmax-by-bmax column-major block, bmax = max(bz, bg), zero padded; dR[j] = d sigma2 / d theta_j (:96).
 * Out: edata[1], gdata[nparam] BEFORE the log-scale factor exp(w) of :432-433.  NaN + NSAGP_ERR_NAN as nsagp_giekf.
 * One CTA per parameter with P and dP_j in shared memory (2 n^2 doubles <= ~220 KB: n <= ~115); for larger n (<= 160;
 * C4's matern32 shape, n = 137) dP_j is kept in an HBM scratch instead -- same arithmetic, slower. */
int nsagp_giekf_grad(const nsagp_model* model, const double* W, double sigma2, int32_t nparam, const int32_t* latent,
                     const double* dA, const double* dQ, const double* dPinf, const double* dR, const double* y, int64_t T,
                     double* edata, double* gdata);

/* Form of the dense RTS pass of nsagp_giekf (gf_giekf_modulator_nmf_constraints.m:221-253):
 * smoother_form 0 = automatic (the parallel scan over time on the FP64 tensor cores when n <= 80, csrc/ekfscan.cuh),
 * 1 = the first-generation kernels (filter and a smoother that is sequential in time; kept as a cross-check),
 * 2 = scan (error if n > 80).  chunk_len (0 = automatic, 16..256 depending on T): steps composed per CTA;
 * chunks_per_segment (0 = 2 x SM count): the scan works through the signal in segments of that many chunks,
 * which bounds its scratch memory (n^2 doubles per step of a segment).  Process-wide setting. */
int nsagp_giekf_config(int32_t smoother_form, int32_t chunk_len, int32_t chunks_per_segment);

/* Device time of the last nsagp_giekf call on this thread: ms[0] filter passes, ms[1] smoother passes (CUDA events). */
int nsagp_giekf_timings(double* ms, int32_t n);

/* Monte-Carlo reconstruction of the signal and of the NMF components from the posterior marginals -- the consumer of
 * Eft / Varft in every demo (demo_toy_modulators_nmf.m:119-165; experiments/missing_data_music.m:138-176 with
 * sqrt_model = 1: sig = sum_d sqrt(W link(g))_d z_d).  Eft, Varft: M-by-T column-major (M = D + N); W: D-by-N
 * column-major; link = log(1 + exp(g - link_shift)); s samples per step.
 * Z: standard-normal draws, T-by-s-by-M column-major (page i belongs to latent i; a wrapper that fills the pages
 * with randn(T,s) in the order i = 1, D+1, 2, D+2, ... reproduces the reference's random stream), or NULL: the
 * draws are generated on the device (Philox4x32-10 keyed by `seed`, Box-Muller), nothing but the marginals is read.
 * Outputs: Esig, Vsig [T] (mean, var with the N-1 normalisation over the s samples); Eft_mod, Varft_mod N-by-T
 * (may be NULL). */
int nsagp_mc_reconstruct(int32_t D, int32_t N, int64_t T, int32_t s, const double* Eft, const double* Varft, const double* W,
                         double link_shift, int32_t sqrt_model, const double* Z, uint64_t seed, double* Esig, double* Vsig,
                         double* Eft_mod, double* Varft_mod);

/* Batched forms: B independent problems of equal shapes (clips x hyper-parameter
 * grid x finite-difference perturbations; what fminunc does around the nlZ mode,
 * demo_toy_modulators_nmf.m:100-104).  models/liks/tables/outs are arrays of B
 * structs, y is T-by-B column-major (one clip per column). */
int nsagp_ep_ihgp_batch(int32_t B, const nsagp_model* models, const nsagp_lik* liks,
                        const nsagp_ep* ep, const nsagp_tables* tables, const double* y,
                        int64_t T, int32_t mode, nsagp_outputs* outs);
int nsagp_ep_full_batch(int32_t B, const nsagp_model* models, const nsagp_lik* liks,
                        const nsagp_ep* ep, const double* y, int64_t T, int32_t mode,
                        nsagp_outputs* outs);

/* Device-resident plan API (what the host-buffer calls above are built from; the
 * benchmark uses it to time the sweep with inputs already in HBM).
 *   create : uploads model, likelihood, tables and y for B problems
 *   run    : executes the whole EP schedule on the device (no host copies)
 *   fetch  : copies the requested outputs of problem b back to the host
 * kind: 0 = ihgp, 1 = full-state EP. */
typedef struct nsagp_plan nsagp_plan;
int nsagp_plan_create(nsagp_plan** plan, int32_t kind, int32_t B, const nsagp_model* models,
                      const nsagp_lik* liks, const nsagp_ep* ep, const nsagp_tables* tables,
                      const double* y, int64_t T, int32_t mode);
/* Full-state predict mode only: also keep the filtered covariances of the last
 * pass (out.PF, gf_ep_modulator_nmf.m:197) -- doubles the covariance storage. */
int nsagp_plan_keep_pf(nsagp_plan* plan, int keep);
/* Form of the sequential (ADF) filter pass, the only part of the schedule that is a
 * nonlinear recurrence in time (ihgp_ep_modulator_nmf.m:253-271, gf_ep_modulator_nmf.m:141-156):
 * 0 = one CTA per problem, its width chosen by the size of the launch (default): full width (one CTA per SM, lowest
 *     latency per step) while there are no more problems than SMs, half width (form 2) beyond,
 * 1 = one warp per problem (first-generation kernel with library arithmetic; kept as an independent implementation),
 * 2 = one CTA per problem with half the moment threads and the steady-state tables left in HBM / L1, so that two CTAs
 *     share an SM: one problem's Kalman section overlaps the other's cubature (256 problems on 148 SMs: 1.39x),
 * 3 = full-width CTAs whatever the batch size. */
int nsagp_plan_set_adf_form(nsagp_plan* plan, int form);
int nsagp_plan_run(nsagp_plan* plan);
int nsagp_plan_fetch(nsagp_plan* plan, int32_t b, nsagp_outputs* out);
int nsagp_plan_destroy(nsagp_plan* plan);
/* Time-chunked execution of ONE signal over several GPUs (infinite-horizon predict mode, B = 1).
 * Every rank creates the same plan over the whole signal, declares the range [t0, t1) of time steps
 * whose frozen-site passes it executes, and the host drives the EP schedule stage by stage, moving
 * the O(state^2) scan aggregates, the one-step site halo and the lZ / max-diff scalars between ranks
 * (nonstationary-audio-gp_b200/chunked.py does it with torch.distributed over NCCL).  The ADF pass
 * (ihgp_ep_modulator_nmf.m:253-271) is a nonlinear recurrence and runs replicated on every rank.
 * Stage numbers and their buffers: csrc/api_chunk.inc (enum ChunkStage). */
int nsagp_plan_set_range(nsagp_plan* plan, int64_t t0, int64_t t1);
int nsagp_plan_stage(nsagp_plan* plan, int32_t stage, double x, int64_t k, const double* in, int64_t n_in,
                     double* out, int64_t n_out);
/* The same with a DEVICE-SIDE carry exchange (csrc/comm.cuh): every rank owns a mailbox in its HBM that its peers
 * write through NVLink peer access; one small kernel per pass posts the rank's record (scan aggregate, one-step site
 * halo, end-point mean, lZ partial sum) into every peer's mailbox and waits for theirs.  nsagp_plan_run_chunked
 * enqueues the rank's whole EP schedule (ihgp_ep_modulator_nmf.m:223-454 / gf_ep_modulator_nmf.m:113-283) without
 * a host synchronisation.  Set-up: nsagp_comm_create on every rank, exchange the 64-byte IPC handles
 * (nsagp_comm_export) out of band -- chunked.py uses torch.distributed once -- then nsagp_comm_connect.  Ranks that
 * are threads of one process pass device addresses instead of handles. */
typedef struct nsagp_comm nsagp_comm;
int nsagp_comm_create(nsagp_comm** out, int32_t rank, int32_t world, int64_t slot_doubles);
int nsagp_comm_export(nsagp_comm* comm, void* handle64, uint64_t* local_ptr);
int nsagp_comm_connect(nsagp_comm* comm, const void* handles, const uint64_t* ptrs);
int nsagp_comm_destroy(nsagp_comm* comm);
int64_t nsagp_plan_comm_slot_doubles(nsagp_plan* plan);
int nsagp_plan_run_chunked(nsagp_plan* plan, nsagp_comm* comm);
/* Opt-in, approximate, error-reported: the first filter pass (ADF; ihgp_ep_modulator_nmf.m:233-310,
 * gf_ep_modulator_nmf.m:126-184) is a nonlinear recurrence in time.  With chunks > 1 it is run as `chunks` time
 * chunks in parallel (one CTA each), every chunk starting `burnin` steps early from the stationary prior and
 * discarding the burn-in; with nsagp_plan_run_chunked every rank does so over its own range only.  The filter
 * forgets its start at the rate of the slowest latent, so the deviation from the exact pass decays with `burnin`;
 * nsagp_plan_adf_mismatch returns the MEASURED disagreement at the chunk boundaries of the last run:
 * out2[0] = max |mean a chunk holds after its burn-in - mean the preceding chunk stored for that step|,
 * out2[1] = max |stored mean| there.  chunks <= 1: the exact sequential pass (default). */
int nsagp_plan_set_adf_parallel(nsagp_plan* plan, int32_t chunks, int64_t burnin);
int nsagp_plan_adf_mismatch(nsagp_plan* plan, double* out2);
/* Stationary (infinite-horizon) Kalman filter / RTS smoother of the probabilistic filter bank, the step that
 * initialises the subbands before the EP path: the two time loops of
 * matlab/unifying_prob_tf/kernel_ss_kalmanFastFB.m:86-147.  The caller keeps the reference's `dare` calls (:50, :131)
 * and passes the constant matrices, all n-by-n column-major: A, AKHA = A - K H A (:63), the gain Kg = K (:60, n),
 * HA = H A (:77, n), the innovation variance S (:53), and -- for the smoother, NULL = filter only (KF = 1) -- the
 * smoother gain G (:126).  y[T] with NaN = missing (:103-106).  Out: MS n-by-T (filtered or smoothed means, :109/:146)
 * and lik_quad = sum over observed steps of v^2 / (2 S) (:99); the caller adds the constant part (:80). */
int nsagp_fastfb(int32_t n, const double* A, const double* AKHA, const double* Kg, const double* HA, double S,
                 const double* G, const double* y, int64_t T, double* MS, double* lik_quad);
/* Tuning knobs of the frozen-site scans (csrc/scan.cuh): signals of at least `family_min_steps` steps run the subband
 * and the modulator latents each at its own block size instead of the padded one (default 0 = always).  nsagp_scan_merge: 1 (default) = the two families share one CTA tile, whole warps per family, one launch
 * per phase; 0 = one launch per family and phase (the earlier form, kept for comparison).  Results agree to rounding
 * in every combination. */
int nsagp_scan_config(int64_t family_min_steps);
int nsagp_scan_merge(int32_t on);
/* 1 (default) = every CTA tile of a scan starts by bulk-prefetching its input rows into L2 (cp.async.bulk.prefetch.L2);
 * 0 = off.  No effect on results. */
int nsagp_scan_prefetch(int32_t on);
/* Geometry of a CTA tile of the scans: at most max_chunks 32-step chunks (1..16, default 16) and at most max_threads
 * threads (256, the default: 128 registers per thread at two tiles per SM, hardly any spills; or 320: 96 registers,
 * more warps).  No effect on results beyond the association of the tile aggregates (rounding). */
int nsagp_scan_tile(int32_t max_chunks, int32_t max_threads);
/* Form of the smoother-side site update (csrc/siteupd.cuh): 0 (default) = four lanes per time step, sigma points two at
 * a time when the rule has few distinct coordinates (every utp_ws rule), else one at a time; 2 = always one at a time;
 * 1 = the first-generation kernel, one thread per step (csrc/ihgp.cuh).  The environment variable NSAGP_SITE_FORM sets
 * the initial value.  Kept as independent cross-checks; results agree to rounding. */
int nsagp_site_config(int32_t form);
/* Device time (ms, CUDA events on the launch stream) of the phases of the last
 * nsagp_plan_run: [0] total, [1] ADF filter pass, [2] fixed-site filter passes,
 * [3] smoother passes, [4] site-update passes.  Returns the number written. */
int nsagp_plan_timings(nsagp_plan* plan, double* ms, int32_t n);

#ifdef __cplusplus
}
#endif
#endif /* NSAGP_H */
