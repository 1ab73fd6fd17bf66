// nsagp_mex.cpp -- thin MEX gateway from MATLAB to the C ABI of libnsagp.so (include/nsagp.h).
//
// Build (on a machine with MATLAB and the CUDA library built by __graft_entry__.build()):
//   mex -R2018a CXXFLAGS='$CXXFLAGS -std=c++17' -I<repo>/include nsagp_mex.cpp ...
//       -L<repo>/nonstationary-audio-gp_b200/csrc -lnsagp
// It cannot be linked in the build container (no MATLAB); `make -C integration/matlab check`
// compiles it against integration/matlab/stub/mex.h to keep it syntactically honest.
//
// Call forms (used by the .m wrappers in this directory, which keep the reference's
// entry-point names and argument lists):
//   out = nsagp_mex('ep_ihgp', model, lik, ep, tables, yall, mode)
//   out = nsagp_mex('ep_full', model, lik, ep, [],     yall, mode)
//   out = nsagp_mex('giekf', model, W, sigma2, g_iter, l_iter, yall, mode)        (P reset every global iteration)
//   out = nsagp_mex('giekf_carry', model, W, sigma2, g_iter, l_iter, yall, mode)  (gf_giekf_modulator_nmf.m: m, P carried)
//   [edata, gdata] = nsagp_mex('giekf_grad', model, W, sigma2, latent, dA, dQ, dPinf, dR, yall)   (:296-437, GradObj 'on')
//   [Esig, Vsig, Eft_mod, Varft_mod] = nsagp_mex('mc_reconstruct', Eft, Varft, W, link_shift, sqrt_model, s, Z_or_seed)
//   [lZ, dlZ, d2lZ] = nsagp_mex('mom', lik, D, N, ep_fraction, y, mu, s2)
//   [MS, lik_quad] = nsagp_mex('fastfb', A, AKHA, K, HA, S, G_or_empty, y)
// model  : struct with fields D, N, bz, bg, A, Q, Pinf, h   (packed per-latent blocks, see nsagp_model)
// lik    : struct with fields kind, sn2, link_shift, W (D-by-N), wn (1-by-S), xn (N-by-S)
// ep     : struct with fields ep_fraction, ep_damping (vector), ep_itts
// tables : struct with fields r (1-by-nr), PP, PG (packed, see nsagp_tables); PG may be []
// mode   : 0 predict, 1 nlZ, 2 nlZ with running sites (ihgp ..._constraints)
// out    : struct; predict mode: Eft, Varft, lb, ub, ttau, tnu, R, lZ, MF, MS, nlZ, maxDiffM,
//          maxDiffP, n_negcav (and PS for 'ep_full'); nlZ modes: edata, ttau, tnu, R, lZ.
// The gateway only reads mxGetDoubles pointers and writes into mxCreateDoubleMatrix outputs:
// MATLAB owns every array, the library owns every device buffer.
#include <cstring>
#include <string>
#include <vector>

#include "mex.h"
#include "nsagp.h"

namespace {

const mxArray* field(const mxArray* s, const char* name, bool required = true) {
  const mxArray* f = mxIsStruct(s) ? mxGetField(s, 0, name) : nullptr;
  if (!f && required) mexErrMsgIdAndTxt("nsagp:arg", "missing struct field '%s'", name);
  return f;
}

const double* dbl(const mxArray* a, const char* what) {
  if (!a || mxIsEmpty(a)) return nullptr;
  if (!mxIsDouble(a) || mxIsComplex(a)) mexErrMsgIdAndTxt("nsagp:arg", "%s must be a real double array", what);
  return mxGetDoubles(a);
}

double scalar(const mxArray* s, const char* name) { return mxGetScalar(field(s, name)); }

// Every array handed to the C ABI is read by element count there: a mis-sized argument must be an error here, not an
// out-of-bounds read inside MATLAB's address space.
void need(const mxArray* a, size_t count, const char* what) {
  const size_t have = a ? mxGetNumberOfElements(a) : 0;
  if (have != count) mexErrMsgIdAndTxt("nsagp:arg", "%s has %zu elements, expected %zu", what, have, count);
}

void fill_model(const mxArray* m, nsagp_model* out) {
  out->D = (int32_t)scalar(m, "D"); out->N = (int32_t)scalar(m, "N");
  out->bz = (int32_t)scalar(m, "bz"); out->bg = (int32_t)scalar(m, "bg");
  if (out->D < 1 || out->N < 1 || out->bz < 1 || out->bz > 8 || out->bg < 1 || out->bg > 8)
    mexErrMsgIdAndTxt("nsagp:arg", "model: need D, N >= 1 and block sizes in 1..8");
  const size_t nb = (size_t)out->D * out->bz * out->bz + (size_t)out->N * out->bg * out->bg;
  const size_t nh = (size_t)out->D * out->bz + (size_t)out->N * out->bg;
  need(field(m, "A"), nb, "model.A"); need(field(m, "Q"), nb, "model.Q"); need(field(m, "Pinf"), nb, "model.Pinf");
  need(field(m, "h"), nh, "model.h");
  out->A = dbl(field(m, "A"), "model.A"); out->Q = dbl(field(m, "Q"), "model.Q");
  out->Pinf = dbl(field(m, "Pinf"), "model.Pinf"); out->h = dbl(field(m, "h"), "model.h");
}

void fill_lik(const mxArray* l, nsagp_lik* out, int32_t D, int32_t N) {
  out->kind = (int32_t)scalar(l, "kind"); out->sn2 = scalar(l, "sn2"); out->link_shift = scalar(l, "link_shift");
  out->S = (int32_t)mxGetNumberOfElements(field(l, "wn"));
  need(field(l, "W"), (size_t)D * N, "lik.W");
  need(field(l, "xn"), (size_t)N * out->S, "lik.xn");
  out->W = dbl(field(l, "W"), "lik.W");
  out->wn = dbl(field(l, "wn"), "lik.wn"); out->xn = dbl(field(l, "xn"), "lik.xn");
}

void check(int status) {
  if (status == NSAGP_OK) return;
  const char* id = status == NSAGP_ERR_NOT_PD ? "nsagp:notPD" : status == NSAGP_ERR_NONPOS_VAR ? "nsagp:nonposVar"
                 : status == NSAGP_ERR_NAN ? "nsagp:nan" : status == NSAGP_ERR_CUDA ? "nsagp:cuda" : "nsagp:invalid";
  mexErrMsgIdAndTxt(id, "%s", nsagp_last_error());
}

mxArray* put(mxArray* s, const char* name, mwSize r, mwSize c, double** ptr) {
  mxArray* a = mxCreateDoubleMatrix(r, c, mxREAL);
  mxAddField(s, name);
  mxSetField(s, 0, name, a);
  *ptr = mxGetDoubles(a);
  return a;
}

void ep_call(bool ihgp, int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs != 7) mexErrMsgIdAndTxt("nsagp:arg", "usage: out = nsagp_mex(cmd, model, lik, ep, tables, yall, mode)");
  (void)nlhs;
  nsagp_model model; nsagp_lik lik; nsagp_ep ep; nsagp_tables tab;
  fill_model(prhs[1], &model);
  fill_lik(prhs[2], &lik, model.D, model.N);
  ep.ep_fraction = scalar(prhs[3], "ep_fraction");
  ep.ep_itts = (int32_t)scalar(prhs[3], "ep_itts");
  ep.ep_damping = dbl(field(prhs[3], "ep_damping"), "ep.ep_damping");
  if ((int)mxGetNumberOfElements(field(prhs[3], "ep_damping")) < ep.ep_itts)
    mexErrMsgIdAndTxt("nsagp:arg", "ep_damping needs ep_itts entries (it is indexed by itt+1)");
  if (ihgp) {
    tab.nr = (int32_t)mxGetNumberOfElements(field(prhs[4], "r"));
    tab.r = dbl(field(prhs[4], "r"), "tables.r");
    tab.PP = dbl(field(prhs[4], "PP"), "tables.PP");
    tab.PG = dbl(field(prhs[4], "PG", false), "tables.PG");
    const size_t nb2 = (size_t)model.D * model.bz * model.bz + (size_t)model.N * model.bg * model.bg;
    need(field(prhs[4], "PP"), (size_t)tab.nr * nb2, "tables.PP");
    if (tab.PG) need(field(prhs[4], "PG", false), (size_t)tab.nr * 2 * nb2, "tables.PG");
  }
  const double* y = dbl(prhs[5], "yall");
  const int64_t T = (int64_t)mxGetNumberOfElements(prhs[5]);
  const int32_t mode = (int32_t)mxGetScalar(prhs[6]);
  const mwSize M = model.D + model.N, n = model.D * model.bz + model.N * model.bg;
  const mwSize nb = model.D * model.bz * model.bz + model.N * model.bg * model.bg;

  mxArray* s = mxCreateStructMatrix(1, 1, 0, nullptr);
  nsagp_outputs o;
  std::memset(&o, 0, sizeof(o));
  put(s, "ttau", M, T, &o.ttau); put(s, "tnu", M, T, &o.tnu); put(s, "R", M, T, &o.R); put(s, "lZ", 1, T, &o.lZ);
  int64_t negcav = 0;
  if (mode == NSAGP_MODE_PREDICT) {
    put(s, "Eft", M, T, &o.Eft); put(s, "Varft", M, T, &o.Varft); put(s, "lb", M, T, &o.lb); put(s, "ub", M, T, &o.ub);
    put(s, "MF", n, T, &o.MF); put(s, "MS", n, T, &o.MS);
    put(s, "nlZ", 1, ep.ep_itts, &o.nlZ); put(s, "maxDiffM", 1, ep.ep_itts, &o.maxDiffM);
    put(s, "maxDiffP", 1, ep.ep_itts, &o.maxDiffP);
    if (!ihgp) put(s, "PS", nb, T, &o.PS);
    o.n_negcav = &negcav;
  } else {
    put(s, "edata", 1, 1, &o.edata);
  }
  check(ihgp ? nsagp_ep_ihgp(&model, &lik, &ep, &tab, y, T, mode, &o) : nsagp_ep_full(&model, &lik, &ep, y, T, mode, &o));
  if (mode == NSAGP_MODE_PREDICT) {
    double* p;
    put(s, "n_negcav", 1, 1, &p);
    *p = (double)negcav;
  }
  plhs[0] = s;
}

void giekf_call(bool carry, int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs != 8) mexErrMsgIdAndTxt("nsagp:arg", "usage: out = nsagp_mex('giekf', model, W, sigma2, g_iter, l_iter, yall, mode)");
  (void)nlhs;
  nsagp_model model;
  fill_model(prhs[1], &model);
  need(prhs[2], (size_t)model.D * model.N, "W");
  const double* W = dbl(prhs[2], "W");
  const double sigma2 = mxGetScalar(prhs[3]);
  const int32_t g_iter = (int32_t)mxGetScalar(prhs[4]), l_iter = (int32_t)mxGetScalar(prhs[5]);
  const double* y = dbl(prhs[6], "yall");
  const int64_t T = (int64_t)mxGetNumberOfElements(prhs[6]);
  const int32_t mode = (int32_t)mxGetScalar(prhs[7]);
  const mwSize M = model.D + model.N, n = model.D * model.bz + model.N * model.bg;
  mxArray* s = mxCreateStructMatrix(1, 1, 0, nullptr);
  nsagp_outputs o;
  std::memset(&o, 0, sizeof(o));
  if (mode == NSAGP_MODE_PREDICT) {
    put(s, "Eft", M, T, &o.Eft); put(s, "Varft", M, T, &o.Varft); put(s, "lb", M, T, &o.lb); put(s, "ub", M, T, &o.ub);
    put(s, "MF", n, T, &o.MF); put(s, "MS", n, T, &o.MS); put(s, "maxDiffP", 1, g_iter, &o.maxDiffP);
    check((carry ? nsagp_giekf_carry : nsagp_giekf)(&model, W, sigma2, g_iter, l_iter, y, T, mode, &o));
  } else {
    put(s, "edata", 1, 1, &o.edata);
    const int st = (carry ? nsagp_giekf_carry : nsagp_giekf)(&model, W, sigma2, g_iter, l_iter, y, T, mode, &o);
    if (st != NSAGP_ERR_NAN) check(st);          // a NaN energy is a value the reference returns to the optimiser
  }
  plhs[0] = s;
}

// [edata, gdata] = nsagp_mex('giekf_grad', model, W, sigma2, latent, dA, dQ, dPinf, dR, yall): latent nparam-by-1
// (0-based, -1 = none), dA / dQ / dPinf bmax-by-bmax-by-nparam (include/nsagp.h nsagp_giekf_grad).
void giekf_grad_call(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs != 10) mexErrMsgIdAndTxt("nsagp:arg", "usage: [edata, gdata] = nsagp_mex('giekf_grad', model, W, sigma2, latent, dA, dQ, dPinf, dR, yall)");
  nsagp_model model;
  fill_model(prhs[1], &model);
  need(prhs[2], (size_t)model.D * model.N, "W");
  const double* W = dbl(prhs[2], "W");
  const double sigma2 = mxGetScalar(prhs[3]);
  const size_t nparam = mxGetNumberOfElements(prhs[4]);
  const size_t bmax = (size_t)(model.bz > model.bg ? model.bz : model.bg);
  const double* latd = dbl(prhs[4], "latent");
  std::vector<int32_t> latent(nparam);
  for (size_t j = 0; j < nparam; ++j) latent[j] = (int32_t)latd[j];
  need(prhs[5], nparam * bmax * bmax, "dA"); need(prhs[6], nparam * bmax * bmax, "dQ"); need(prhs[7], nparam * bmax * bmax, "dPinf");
  need(prhs[8], nparam, "dR");
  const double* y = dbl(prhs[9], "yall");
  const int64_t T = (int64_t)mxGetNumberOfElements(prhs[9]);
  plhs[0] = mxCreateDoubleMatrix(1, 1, mxREAL);
  mxArray* g = mxCreateDoubleMatrix(1, (mwSize)nparam, mxREAL);
  const int st = nsagp_giekf_grad(&model, W, sigma2, (int32_t)nparam, latent.data(), dbl(prhs[5], "dA"), dbl(prhs[6], "dQ"),
                                  dbl(prhs[7], "dPinf"), dbl(prhs[8], "dR"), y, T, mxGetDoubles(plhs[0]), mxGetDoubles(g));
  if (st != NSAGP_ERR_NAN) check(st);            // NaN energy and gradient are values the reference returns (:391-394)
  if (nlhs > 1) plhs[1] = g; else mxDestroyArray(g);
}

// [Esig, Vsig, Eft_mod, Varft_mod] = nsagp_mex('mc_reconstruct', Eft, Varft, W, link_shift, sqrt_model, s, Z_or_seed)
// Z_or_seed: a T-by-s-by-M array of standard-normal draws (page i = latent i), or a scalar seed.
void mc_call(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs != 8) mexErrMsgIdAndTxt("nsagp:arg", "usage: [Esig,Vsig,Eft_mod,Varft_mod] = nsagp_mex('mc_reconstruct', Eft, Varft, W, link_shift, sqrt_model, s, Z_or_seed)");
  const int32_t D = (int32_t)mxGetM(prhs[3]), N = (int32_t)mxGetN(prhs[3]);
  const int64_t T = (int64_t)mxGetN(prhs[1]);
  const int32_t s = (int32_t)mxGetScalar(prhs[6]);
  if ((int64_t)mxGetM(prhs[1]) != D + N) mexErrMsgIdAndTxt("nsagp:arg", "Eft must be (D+N)-by-T");
  need(prhs[2], (size_t)(D + N) * (size_t)T, "Varft");
  const bool seeded = mxGetNumberOfElements(prhs[7]) == 1;
  if (!seeded && (int64_t)mxGetNumberOfElements(prhs[7]) != T * s * (D + N)) mexErrMsgIdAndTxt("nsagp:arg", "Z must be T-by-s-by-(D+N)");
  mxArray* Es = mxCreateDoubleMatrix(T, 1, mxREAL);
  mxArray* Vs = mxCreateDoubleMatrix(T, 1, mxREAL);
  mxArray* Em = mxCreateDoubleMatrix(N, T, mxREAL);
  mxArray* Vm = mxCreateDoubleMatrix(N, T, mxREAL);
  check(nsagp_mc_reconstruct(D, N, T, s, dbl(prhs[1], "Eft"), dbl(prhs[2], "Varft"), dbl(prhs[3], "W"), mxGetScalar(prhs[4]),
                             (int32_t)mxGetScalar(prhs[5]), seeded ? nullptr : dbl(prhs[7], "Z"),
                             seeded ? (uint64_t)mxGetScalar(prhs[7]) : 0, mxGetDoubles(Es), mxGetDoubles(Vs), mxGetDoubles(Em),
                             mxGetDoubles(Vm)));
  plhs[0] = Es;
  if (nlhs > 1) plhs[1] = Vs; else mxDestroyArray(Vs);
  if (nlhs > 2) plhs[2] = Em; else mxDestroyArray(Em);
  if (nlhs > 3) plhs[3] = Vm; else mxDestroyArray(Vm);
}

void mom_call(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs != 8) mexErrMsgIdAndTxt("nsagp:arg", "usage: [lZ,dlZ,d2lZ] = nsagp_mex('mom', lik, D, N, ep_fraction, y, mu, s2)");
  nsagp_lik lik;
  const int32_t D = (int32_t)mxGetScalar(prhs[2]), N = (int32_t)mxGetScalar(prhs[3]);
  fill_lik(prhs[1], &lik, D, N);
  const int64_t T = (int64_t)mxGetNumberOfElements(prhs[5]);
  need(prhs[6], (size_t)(D + N) * (size_t)T, "mu"); need(prhs[7], (size_t)(D + N) * (size_t)T, "s2");
  mxArray* lZ = mxCreateDoubleMatrix(1, T, mxREAL);
  mxArray* d1 = mxCreateDoubleMatrix(D + N, T, mxREAL);
  mxArray* d2 = mxCreateDoubleMatrix(D + N, T, mxREAL);
  check(nsagp_mom_batch(&lik, D, N, mxGetScalar(prhs[4]), T, dbl(prhs[5], "y"), dbl(prhs[6], "mu"), dbl(prhs[7], "s2"),
                        mxGetDoubles(lZ), mxGetDoubles(d1), mxGetDoubles(d2)));
  plhs[0] = lZ;
  if (nlhs > 1) plhs[1] = d1; else mxDestroyArray(d1);
  if (nlhs > 2) plhs[2] = d2; else mxDestroyArray(d2);
}

// [MS, lik_quad] = nsagp_mex('fastfb', A, AKHA, K, HA, S, G_or_empty, y)   (kernel_ss_kalmanFastFB.m:86-147)
void fastfb_call(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs != 8) mexErrMsgIdAndTxt("nsagp:arg", "usage: [MS, lik_quad] = nsagp_mex('fastfb', A, AKHA, K, HA, S, G, y)");
  const int32_t n = (int32_t)mxGetM(prhs[1]);
  const size_t nn = (size_t)n * n;
  need(prhs[1], nn, "A"); need(prhs[2], nn, "AKHA"); need(prhs[3], (size_t)n, "K"); need(prhs[4], (size_t)n, "HA");
  const bool smooth = !mxIsEmpty(prhs[6]);
  if (smooth) need(prhs[6], nn, "G");
  const int64_t T = (int64_t)mxGetNumberOfElements(prhs[7]);
  mxArray* MS = mxCreateDoubleMatrix(n, T, mxREAL);
  double quad = 0.0;
  check(nsagp_fastfb(n, dbl(prhs[1], "A"), dbl(prhs[2], "AKHA"), dbl(prhs[3], "K"), dbl(prhs[4], "HA"), mxGetScalar(prhs[5]),
                     smooth ? dbl(prhs[6], "G") : nullptr, dbl(prhs[7], "y"), T, mxGetDoubles(MS), &quad));
  plhs[0] = MS;
  if (nlhs > 1) plhs[1] = mxCreateDoubleScalar(quad);
}

}  // namespace

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs < 1 || !mxIsChar(prhs[0])) mexErrMsgIdAndTxt("nsagp:arg", "first argument must be a command string");
  char cmd[32];
  mxGetString(prhs[0], cmd, sizeof(cmd));
  const std::string c(cmd);
  if (c == "ep_ihgp") ep_call(true, nlhs, plhs, nrhs, prhs);
  else if (c == "ep_full") ep_call(false, nlhs, plhs, nrhs, prhs);
  else if (c == "giekf") giekf_call(false, nlhs, plhs, nrhs, prhs);
  else if (c == "giekf_carry") giekf_call(true, nlhs, plhs, nrhs, prhs);
  else if (c == "giekf_grad") giekf_grad_call(nlhs, plhs, nrhs, prhs);
  else if (c == "mc_reconstruct") mc_call(nlhs, plhs, nrhs, prhs);
  else if (c == "mom") mom_call(nlhs, plhs, nrhs, prhs);
  else if (c == "fastfb") fastfb_call(nlhs, plhs, nrhs, prhs);
  else if (c == "version") plhs[0] = mxCreateString(nsagp_version());
  else mexErrMsgIdAndTxt("nsagp:arg", "unknown command '%s'", cmd);
}
