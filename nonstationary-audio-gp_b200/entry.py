"""Entry points with the reference's own names, argument order and return values.

    gf_ep_modulator_nmf                  matlab/gf_ep_modulator_nmf.m:1
    gf_ep_modulator_nmf_constraints      matlab/gf_ep_modulator_nmf_constraints.m:1
    ihgp_ep_modulator_nmf                matlab/ihgp_ep_modulator_nmf.m:1
    ihgp_ep_modulator_nmf_constraints    matlab/ihgp_ep_modulator_nmf_constraints.m:1

Same parameter vectors in; ``(nlZ, grad)`` (xt empty) or
``(Eft, Varft, Covft, lb, ub, out)`` (xt given) out.  Host work is limited to
what the reference also does once per call with MATLAB built-ins (unpacking,
``ss``, ``balance``, ``lti_disc``, DARE tables); the time loops run in the CUDA
library through the C ABI (include/nsagp.h).  Indices are 0-based.
"""
import ctypes as C

import numpy as np

from . import _lib, ssmodel, tables as tables_mod
from .lik import Moments


# --------------------------------------------------------------------- helpers
def merge_inputs(x, y, xt):
    """Combine observations and test points, sort, de-duplicate keeping the first
    occurrence; test-only points carry NaN (gf_ep_modulator_nmf.m:58-66)."""
    x = np.asarray(x, float).ravel()
    y = np.asarray(y, float).ravel()
    if x.size != y.size:
        raise ValueError("x and y must have the same number of elements")
    xt = np.zeros(0) if xt is None else np.asarray(xt, float).ravel()
    xall = np.concatenate([x, xt])
    yall = np.concatenate([y, np.full(xt.size, np.nan)])
    _, first, inverse = np.unique(xall, return_index=True, return_inverse=True)
    return yall[first], inverse[xall.size - xt.size:]


def sigmoid(x, sig_range=(0.0, 20.0), c=0.0, a=1.0):
    """matlab/sigmoid.m:17-19."""
    lo, up = float(sig_range[0]), float(sig_range[-1])
    return (up - lo) / (1.0 + np.exp(-a * (np.asarray(x, float) - c))) + lo


def inv_sigmoid(y, sig_range=(0.0, 20.0), c=0.0, a=1.0):
    """matlab/inv_sigmoid.m:17-23."""
    lo, up = float(sig_range[0]), float(sig_range[-1])
    y = np.asarray(y, float)
    ratio = (up - y) / (y - lo)
    if np.any(~(ratio > 0)):
        raise ValueError("Error with inverse sigmoid transformation: parameter outside of user specified range")
    return c - np.log(ratio) / a


def lambda_map(lin, kernel):
    """matlab/lambda_map.m."""
    c = {"exp": 1.0, "matern32": 3.0 ** 0.5, "matern52": 5.0 ** 0.5, "matern72": 7.0 ** 0.5}[kernel]
    return c / np.asarray(lin, float)


def _unpack_log(w, nlik, D, N):
    w = np.asarray(w, float).ravel()
    need = nlik + 3 * D + 2 * N + D * N
    if w.size != need:
        raise ValueError("w has %d entries, expected %d" % (w.size, need))
    e = np.exp(w[nlik:])
    return w[:nlik], e[:3 * D], e[3 * D:3 * D + 2 * N], e[3 * D + 2 * N:].reshape((D, N), order="F")


def _unpack_constrained(w, nlik, D, N, constraints, w_fixed, tune_hypers):
    """Split tuned / fixed parameters and squash them into their boxes
    (gf_ep_modulator_nmf_constraints.m:75-110)."""
    src = {True: [np.asarray(w, float).ravel(), 0], False: [np.asarray(w_fixed, float).ravel(), 0]}
    cons = np.asarray(constraints, float)
    tune = [bool(t) for t in np.asarray(tune_hypers).ravel()]

    def take(flag, count):
        arr, pos = src[flag]
        src[flag][1] = pos + count
        return arr[pos:pos + count]

    lik_param = take(tune[0], nlik)
    groups = [sigmoid(take(tune[i], D if i <= 3 else N), cons[i - 1]) for i in range(1, 6)]
    arr, pos = src[tune[6]]
    Wnmf = sigmoid(arr[pos:], cons[5]).reshape((D, N), order="F")
    return lik_param, np.concatenate(groups[:3]), np.concatenate(groups[3:]), Wnmf


def _discrete_model(ss, x, param1, param2, kernel1, kernel2, D, N, balance, symmetrise_Q):
    F, L, Qc, H, Pinf = ss(x, param1, param2, kernel1, kernel2)[:5]
    if balance:
        F, L, H, Pinf = ssmodel.balance(F, L, H, Pinf)
    A, Q = ssmodel.lti_disc(F, L, Qc, 1.0)              # dt = 1 is hard-coded in the reference
    if symmetrise_Q:
        Q = (Q + Q.T) / 2                               # ihgp_ep_modulator_nmf.m:97
    return ssmodel.to_block_model(A, Q, H, Pinf, D, N), H


def _require_moments(mom):
    if not isinstance(mom, Moments):
        raise TypeError("`mom` must be built with likModulatorNMFPower(...) or likModulatorPreCalcwn(...): "
                        "an arbitrary function handle cannot run on the GPU")
    return mom


# ------------------------------------------------------------------------ plan
class Plan:
    """Device-resident EP problem(s): thin wrapper over nsagp_plan_* (include/nsagp.h).

    models: list of ssmodel.BlockModel; liks: list of (Moments, lik_param, W);
    y: (B, T) array (NaN = missing); tables: list of tables.IhgpTables for kind 0."""

    def __init__(self, kind, models, liks, ep_fraction, ep_damping, ep_itts, y, mode, tables=None):
        L = _lib.lib()
        self.kind, self.mode = kind, mode
        self.B = len(models)
        y = _lib.as_f64(np.atleast_2d(y))
        if y.shape[0] != self.B:
            raise ValueError("y must have one row per problem")
        self.T = y.shape[1]
        self.models = models
        self.ep_itts = int(ep_itts)
        damping = _lib.as_f64(np.atleast_1d(ep_damping).ravel())
        if damping.size < self.ep_itts:
            raise ValueError("ep_damping must have at least ep_itts entries (it is indexed by itt+1)")
        keep = [y, damping]
        cm = (_lib.Model * self.B)()
        cl = (_lib.Lik * self.B)()
        ct = (_lib.Tables * self.B)() if kind == _lib.KIND_IHGP else None
        for b, (mdl, (mom, lik_param, W)) in enumerate(zip(models, liks)):
            arrs = [_lib.as_f64(a) for a in (mdl.A, mdl.Q, mdl.Pinf, mdl.h)]
            keep.extend(arrs)
            cm[b].D, cm[b].N, cm[b].bz, cm[b].bg = mdl.D, mdl.N, mdl.bz, mdl.bg
            cm[b].A, cm[b].Q, cm[b].Pinf, cm[b].h = [_lib.dptr(a) for a in arrs]
            cl[b] = _require_moments(mom).c_lik(lik_param, W, keep)
            if ct is not None:
                tb = tables[b]
                pp, pg = tb.packed()
                r = _lib.as_f64(tb.r); pp = _lib.as_f64(pp)
                keep.extend([r, pp])
                ct[b].nr = r.size; ct[b].r = _lib.dptr(r); ct[b].PP = _lib.dptr(pp)
                if pg is not None:
                    pg = _lib.as_f64(pg); keep.append(pg)
                    ct[b].PG = _lib.dptr(pg)
        ep = _lib.Ep(float(ep_fraction), _lib.dptr(damping), self.ep_itts)
        self._h = C.c_void_p()
        _lib.check(L.nsagp_plan_create(C.byref(self._h), kind, self.B, cm, cl, C.byref(ep), ct,
                                       _lib.dptr(y), self.T, mode))
        self._keep = keep

    def keep_pf(self, keep=True):
        """Also keep the filtered covariances of the last pass (out.PF); full-state predict mode."""
        _lib.check(_lib.lib().nsagp_plan_keep_pf(self._h, int(keep)))
        return self

    def set_adf_form(self, form):
        """0: one CTA per problem, full width up to one problem per SM, half width (two CTAs per SM) beyond (default);
        1: one warp per problem (first-generation kernel); 2 / 3: force half / full width."""
        _lib.check(_lib.lib().nsagp_plan_set_adf_form(self._h, int(form)))
        return self

    def set_adf_parallel(self, chunks, burnin):
        """Opt-in, approximate: first filter pass as `chunks` parallel time chunks with `burnin` steps of burn-in
        overlap each (include/nsagp.h).  chunks <= 1 restores the exact sequential pass."""
        _lib.check(_lib.lib().nsagp_plan_set_adf_parallel(self._h, int(chunks), int(burnin)))
        return self

    def adf_mismatch(self):
        """(max |boundary mismatch of the means|, max |mean| there) of the last run with a parallel first pass."""
        out = np.zeros(2)
        _lib.check(_lib.lib().nsagp_plan_adf_mismatch(self._h, _lib.dptr(out)))
        return float(out[0]), float(out[1])

    def run(self):
        _lib.check(_lib.lib().nsagp_plan_run(self._h))
        return self

    def timings(self):
        ms = np.zeros(5)
        _lib.lib().nsagp_plan_timings(self._h, _lib.dptr(ms), 5)
        return dict(total=ms[0], adf=ms[1], fixed_filter=ms[2], smoother=ms[3], site_update=ms[4])

    def fetch(self, b=0, names=("Eft", "Varft", "lb", "ub")):
        """Copy the named outputs of problem b to the host.  M-by-T / n-by-T arrays
        come back in MATLAB orientation (sites or states along axis 0)."""
        mdl = self.models[b]
        M, n, T = mdl.M, mdl.n, self.T
        nb = mdl.D * mdl.bz ** 2 + mdl.N * mdl.bg ** 2
        shapes = dict(Eft=(T, M), Varft=(T, M), lb=(T, M), ub=(T, M), ttau=(T, M), tnu=(T, M), R=(T, M), lZ=(T,),
                      MF=(T, n), MS=(T, n), PF=(T, nb), PS=(T, nb), nlZ=(self.ep_itts,),
                      maxDiffM=(self.ep_itts,), maxDiffP=(self.ep_itts,), edata=(1,))
        o = _lib.Outputs()
        bufs = {}
        for nm in names:
            if nm == "n_negcav":
                bufs[nm] = np.zeros(1, np.int64)
                o.n_negcav = bufs[nm].ctypes.data_as(C.POINTER(C.c_int64))
            else:
                bufs[nm] = np.empty(shapes[nm])
                setattr(o, nm, _lib.dptr(bufs[nm]))
        _lib.check(_lib.lib().nsagp_plan_fetch(self._h, b, C.byref(o)))
        out = {}
        for nm, a in bufs.items():
            if nm == "n_negcav":
                out[nm] = int(a[0])
            elif nm == "edata":
                out[nm] = float(a[0])
            else:
                out[nm] = a.T if a.ndim == 2 else a
        return out

    def close(self):
        if self._h:
            _lib.lib().nsagp_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_DEBUG = ("tnu", "ttau", "lZ", "R", "MF", "MS", "nlZ", "maxDiffM", "maxDiffP", "n_negcav")


def _predict_outputs(plan, mdl, return_ind, nargout, debug_cov):
    names = ["Eft"]
    if nargout > 1:
        names.append("Varft")
    if nargout > 3:
        names += ["lb", "ub"]
    if nargout > 5:
        names += list(_DEBUG)
        if debug_cov and plan.kind == _lib.KIND_FULL:
            names += ["PF", "PS"]
    res = plan.fetch(0, names)
    sel = lambda a: a[:, return_ind]
    Eft = sel(res["Eft"])
    if nargout <= 1:
        return (Eft,)
    Varft = sel(res["Varft"])
    if nargout <= 3:
        return Eft, Varft
    lb, ub = sel(res["lb"]), sel(res["ub"])
    if nargout <= 5:
        return Eft, Varft, None, lb, ub
    out = {k: res[k] for k in _DEBUG}
    for k in ("PF", "PS"):
        if k in res:
            out[k] = res[k]              # packed per-block covariances, (sum b^2)-by-T
    return Eft, Varft, None, lb, ub, out


def _run(kind, w, x, y, ss, mom, xt, kernel1, kernel2, num_lik_params, D, N, ep_fraction, ep_damping, ep_itts,
         constrained, balance, nargout, debug_cov, nlz_mode=_lib.MODE_NLZ, adf_form=0):
    mom = _require_moments(mom)
    yall, return_ind = merge_inputs(x, y, xt)
    if constrained is None:
        lik_param, param1, param2, Wnmf = _unpack_log(w, num_lik_params, D, N)
    else:
        lik_param, param1, param2, Wnmf = _unpack_constrained(w, num_lik_params, D, N, *constrained)
    mdl, _ = _discrete_model(ss, x, param1, param2, kernel1, kernel2, D, N, balance,
                             symmetrise_Q=(kind == _lib.KIND_IHGP))
    predict = xt is not None and np.size(xt) > 0
    tabs = None
    if kind == _lib.KIND_IHGP:
        tabs = [tables_mod.build_tables(mdl, want_smoother=predict)]
    mode = _lib.MODE_PREDICT if predict else nlz_mode
    with Plan(kind, [mdl], [(mom, lik_param, Wnmf)], ep_fraction, ep_damping, ep_itts, yall[None, :], mode,
              tables=tabs) as plan:
        if predict and debug_cov and kind == _lib.KIND_FULL and nargout > 5:
            plan.keep_pf()
        plan.set_adf_form(adf_form)
        plan.run()
        if predict:
            res = _predict_outputs(plan, mdl, return_ind, nargout, debug_cov)
            if nargout > 5 and tabs is not None:
                res[5].update(r=tabs[0].r, ro=tabs[0].ro, PPlist=tabs[0].PP)     # ihgp_ep_modulator_nmf.m:137-141
            return res
        edata = plan.fetch(0, ("edata",))["edata"]
    # the reference returns an all-zero gradient (gf_ep_modulator_nmf.m:363,531)
    return edata, np.zeros(np.size(w))


# ---------------------------------------------------------------- entry points
def gf_ep_modulator_nmf(w, x, y, ss, mom, xt, kernel1, kernel2, num_lik_params, D, N,
                        ep_fraction, ep_damping, ep_itts, nargout=6, debug_cov=False, adf_form=0):
    """Solve the time-frequency-NMF GP model by Power EP (full-state Kalman filter /
    RTS smoother).  Drop-in for matlab/gf_ep_modulator_nmf.m."""
    return _run(_lib.KIND_FULL, w, x, y, ss, mom, xt, kernel1, kernel2, num_lik_params, D, N, ep_fraction,
                ep_damping, ep_itts, None, False, nargout, debug_cov, adf_form=adf_form)


def gf_ep_modulator(w, x, y, ss, mom, xt, kernel1, kernel2, num_lik_params, ep_fraction, ep_damping, ep_itts,
                    nargout=6, debug_cov=False, adf_form=0):
    """Drop-in for matlab/gf_ep_modulator.m, the model WITHOUT NMF weights (demo_toy_modulators.m:99,109): D carrier x
    modulator pairs, y = sum_d z_d link(g_d).  It is gf_ep_modulator_nmf.m line by line with W = I and N = D (the
    files differ in the ``mom`` signature, :138,:227), run on the balanced model (:75-81): ``ss(x, param, k1, k2)``
    takes the five parameter groups in one vector (:69-72), ``mom`` comes from ``likModulatorPower``.  The sigma
    points live in D dimensions, so D <= 4 pairs (the demo has 2)."""
    w = np.asarray(w, float).ravel()
    nlik = int(num_lik_params)
    if (w.size - nlik) % 5:
        raise ValueError("w must hold num_lik_params + 5 D entries")
    pairs = (w.size - nlik) // 5
    # the NMF entry point's parameter vector: [lik; var1, len1, omega; var2, len2; log W(:)] with W = I
    with np.errstate(divide="ignore"):
        w_nmf = np.concatenate([w, np.log(np.eye(pairs).reshape(-1, order="F"))])
    ss_nmf = lambda x_, p1, p2, k1, k2: ss(x_, np.concatenate([p1, p2]), k1, k2)
    res = _run(_lib.KIND_FULL, w_nmf, x, y, ss_nmf, mom, xt, kernel1, kernel2, nlik, pairs, pairs, ep_fraction,
               ep_damping, ep_itts, None, True, nargout, debug_cov, adf_form=adf_form)
    if not (xt is not None and np.size(xt) > 0):
        return res[0], np.zeros(w.size)
    return res


def gf_ep_modulator_nmf_constraints(w, x, y, ss, mom, xt, kernel1, kernel2, num_lik_params, D, N,
                                    ep_fraction, ep_damping, ep_itts, constraints, w_fixed, tune_hypers,
                                    nargout=6, debug_cov=False, adf_form=0):
    """matlab/gf_ep_modulator_nmf_constraints.m: box-constrained parameters, balanced model."""
    return _run(_lib.KIND_FULL, w, x, y, ss, mom, xt, kernel1, kernel2, num_lik_params, D, N, ep_fraction,
                ep_damping, ep_itts, (constraints, w_fixed, tune_hypers), True, nargout, debug_cov, adf_form=adf_form)


def ihgp_ep_modulator_nmf(w, x, y, ss, mom, xt, kernel1, kernel2, num_lik_params, D, N,
                          ep_fraction, ep_damping, ep_itts, nargout=6, adf_form=0):
    """Power EP with the infinite-horizon (steady-state gain) approximation.
    Drop-in for matlab/ihgp_ep_modulator_nmf.m."""
    return _run(_lib.KIND_IHGP, w, x, y, ss, mom, xt, kernel1, kernel2, num_lik_params, D, N, ep_fraction,
                ep_damping, ep_itts, None, True, nargout, False, adf_form=adf_form)


def ihgp_ep_modulator_nmf_constraints(w, x, y, ss, mom, xt, kernel1, kernel2, num_lik_params, D, N,
                                      ep_fraction, ep_damping, ep_itts, constraints, w_fixed, tune_hypers,
                                      nargout=6, adf_form=0):
    """matlab/ihgp_ep_modulator_nmf_constraints.m (its nlZ mode carries the site
    vectors from step to step, :568-615)."""
    return _run(_lib.KIND_IHGP, w, x, y, ss, mom, xt, kernel1, kernel2, num_lik_params, D, N, ep_fraction,
                ep_damping, ep_itts, (constraints, w_fixed, tune_hypers), True, nargout, False,
                nlz_mode=_lib.MODE_NLZ_RUNNING, adf_form=adf_form)


def gf_giekf_modulator_nmf(w, x, y, ss, mom, xt, kernel1, kernel2, num_lik_params, D, N, g_iter, l_iter,
                           GradObj="off", nargout=6, debug_cov=False, balance_derivatives=False):
    """Drop-in for matlab/gf_giekf_modulator_nmf.m: log-scale parameter vector
    (:70-73), balanced model (:78-85), and the state (m, P) initialised on the first global iteration
    only (:127-131) -- the smoothed mean and covariance of step 1 start the next filter pass.

    ``GradObj='on'`` with ``xt`` empty returns ``(energy, gradient)`` with the analytic gradient of :296-437 over
    ``w(1:end-D*N)`` (the sensitivity equations; csrc/ekfgrad.cuh).  The reference balances F, L, H, Pinf but leaves
    dF and dPinf in the unbalanced coordinates (its balancing loop is commented out, :82-84), so its gradient is the
    derivative of the energy only where balancing is the identity; ``balance_derivatives=True`` runs that loop."""
    if GradObj != "off" and not (xt is not None and np.size(xt) > 0):
        return _giekf_grad(w, x, y, ss, kernel1, kernel2, num_lik_params, D, N, balance_derivatives)
    return _giekf(w, x, y, ss, xt, kernel1, kernel2, num_lik_params, D, N, g_iter, l_iter, None, nargout, debug_cov)


def gf_giekf_modulator_nmf_constraints(w, x, y, ss, mom, xt, kernel1, kernel2, num_lik_params, D, N,
                                       g_iter, l_iter, constraints, w_fixed, tune_hypers, nargout=6, debug_cov=False):
    """Globally iterated EKF + RTS smoother, the comparison variant of the EP entry points.
    Drop-in for matlab/gf_giekf_modulator_nmf_constraints.m with GradObj = 'off' (how every caller
    runs it, experiments/train_model.m:226,239-240): ``(energy, zeros)`` when ``xt`` is empty,
    otherwise ``(Eft, Varft, Covft, lb, ub, out)``.  ``mom`` is ignored, as in the reference
    (its measurement model is hard-wired, :133-138)."""
    return _giekf(w, x, y, ss, xt, kernel1, kernel2, num_lik_params, D, N, g_iter, l_iter,
                  (constraints, w_fixed, tune_hypers), nargout, debug_cov)


def _deriv_blocks(stack, starts, sizes):
    """[(latent, block)] of a derivative stack: a ssmodel.DerivStack as it is, a dense n x n x P array (what a
    reference-style ``ss`` closure returns) by locating the one diagonal block each slice occupies."""
    if isinstance(stack, ssmodel.DerivStack):
        return stack.items
    if stack is None:
        raise ValueError("GradObj='on' needs the derivative stacks dF, dQc, dPinf from the ss closure (outputs 6-8)")
    stack = np.asarray(stack, float)
    items = []
    for p in range(stack.shape[2]):
        X = stack[:, :, p]
        rows = np.flatnonzero(np.any(X != 0, axis=1) | np.any(X != 0, axis=0))
        lat = int(np.searchsorted(starts, rows[0], side="right") - 1) if rows.size else 0
        o, b = starts[lat], sizes[lat]
        if rows.size and (rows[0] < o or rows[-1] >= o + b):
            raise ValueError("derivative slice %d is not confined to one latent's block" % p)
        items.append((lat, X[o:o + b, o:o + b].copy()))
    return items


def _giekf_grad(w, x, y, ss, kernel1, kernel2, num_lik_params, D, N, balance_derivatives):
    """Host side of gf_giekf_modulator_nmf.m:296-437: per-parameter blocks of dA, dQ, dPinf (:328-338, :362-364),
    then nsagp_giekf_grad."""
    import scipy.linalg as sla
    yall, _ = merge_inputs(x, y, None)
    w = np.asarray(w, float).ravel()
    lik_param, param1, param2, Wnmf = _unpack_log(w, num_lik_params, D, N)
    res = ss(x, param1, param2, kernel1, kernel2)
    F, L, Qc, H, Pinf = res[:5]
    Fb, Lb, Hb, Pb, Tm = ssmodel.balance(F, L, H, Pinf, return_T=True)              # :78-81
    sigma2 = float(np.exp(np.asarray(lik_param, float).ravel()[0]))
    A = sla.expm(Fb)                                                                 # :355-357
    Q = Pb - A @ Pb @ A.T
    mdl = ssmodel.to_block_model(A, Q, Hb, Pb, D, N)
    st, sizes = mdl.starts(), mdl.block_sizes()
    dFi = _deriv_blocks(res[5] if len(res) > 5 else None, st, sizes)
    dPi = _deriv_blocks(res[7] if len(res) > 7 else None, st, sizes)
    nparam = 1 + len(dFi)                                                            # the noise variance first (:93-96)
    if nparam != w.size - D * N:
        raise ValueError("ss returned %d derivative slices, w(1:end-D*N) has %d entries" % (len(dFi), w.size - D * N))
    bmax = max(mdl.bz, mdl.bg)
    latent = np.full(nparam, -1, np.int32)
    dA = np.zeros((nparam, bmax, bmax)); dQ = np.zeros_like(dA); dP0 = np.zeros_like(dA)
    dR = np.zeros(nparam); dR[0] = 1.0
    for j, ((lat, dFj), (lat2, dPj)) in enumerate(zip(dFi, dPi), start=1):
        if lat != lat2 and np.any(dPj != 0) and np.any(dFj != 0):
            raise ValueError("dF and dPinf of parameter %d live in different latents" % j)
        if not np.any(dFj != 0) and np.any(dPj != 0):
            lat = lat2
        o, b = st[lat], sizes[lat]
        sl = slice(o, o + b)
        if not np.any(dFj != 0):
            dFj = np.zeros((b, b))                                                   # e.g. a variance: F does not depend on it
        if not np.any(dPj != 0):
            dPj = np.zeros((b, b))                                                   # e.g. a frequency: Pinf does not depend on it
        if balance_derivatives:                                                      # the commented-out loop, :82-84
            Tl = Tm[sl, sl]
            dFj = sla.solve(Tl, dFj @ Tl)
            dPj = sla.solve(Tl, sla.solve(Tl, dPj).T).T
        Fl, Pl = Fb[sl, sl], Pb[sl, sl]
        AA = sla.expm(np.block([[Fl, np.zeros((b, b))], [dFj, Fl]]))                 # :328-338, one latent's block
        Al, dAl = AA[:b, :b], AA[b:, :b]
        dAPAt = dAl @ Pl @ Al.T
        latent[j] = lat
        dA[j, :b, :b] = dAl
        dQ[j, :b, :b] = dPj - dAPAt - Al @ dPj @ Al.T - dAPAt.T                      # :362-364
        dP0[j, :b, :b] = dPj
    arrs = [_lib.as_f64(a) for a in (mdl.A, mdl.Q, mdl.Pinf, mdl.h)]
    cm = _lib.Model()
    cm.D, cm.N, cm.bz, cm.bg = mdl.D, mdl.N, mdl.bz, mdl.bg
    cm.A, cm.Q, cm.Pinf, cm.h = [_lib.dptr(a) for a in arrs]
    Wf = np.asfortranarray(np.asarray(Wnmf, float))
    # blocks column-major with leading dimension bmax
    pack = lambda X: _lib.as_f64(np.ascontiguousarray(np.transpose(X, (0, 2, 1))).ravel())
    bA, bQ, bP, vR, yb = pack(dA), pack(dQ), pack(dP0), _lib.as_f64(dR), _lib.as_f64(yall)
    edata = np.zeros(1); gdata = np.zeros(nparam)
    status = _lib.lib().nsagp_giekf_grad(C.byref(cm), Wf.ctypes.data_as(_lib.c_double_p), sigma2, nparam,
                                         latent.ctypes.data_as(C.POINTER(C.c_int32)), _lib.dptr(bA), _lib.dptr(bQ), _lib.dptr(bP),
                                         _lib.dptr(vR), _lib.dptr(yb), yall.size, _lib.dptr(edata), _lib.dptr(gdata))
    if status == -5:                                                    # NSAGP_ERR_NAN: the reference returns NaNs (:391-394)
        return float("nan"), np.full(nparam, np.nan)
    _lib.check(status)
    return float(edata[0]), gdata * np.exp(w[:nparam])                  # :431-433


def _giekf(w, x, y, ss, xt, kernel1, kernel2, num_lik_params, D, N, g_iter, l_iter, constrained, nargout, debug_cov):
    import scipy.linalg as sla
    yall, return_ind = merge_inputs(x, y, xt)
    if constrained is None:
        lik_param, param1, param2, Wnmf = _unpack_log(w, num_lik_params, D, N)
    else:
        lik_param, param1, param2, Wnmf = _unpack_constrained(w, num_lik_params, D, N, *constrained)
    F, L, Qc, H, Pinf = ss(x, param1, param2, kernel1, kernel2)[:5]
    F, L, H, Pinf = ssmodel.balance(F, L, H, Pinf)                      # :113-120
    sigma2 = float(np.exp(np.asarray(lik_param, float).ravel()[0]))
    predict = xt is not None and np.size(xt) > 0
    if predict:
        A, Q = ssmodel.lti_disc(F, L, Qc, 1.0)                          # :161
    else:
        A = sla.expm(F)                                                 # :376-380
        Q = Pinf - A @ Pinf @ A.T
    mdl = ssmodel.to_block_model(A, Q, H, Pinf, D, N)
    T, M, n = yall.size, mdl.M, mdl.n
    arrs = [_lib.as_f64(a) for a in (mdl.A, mdl.Q, mdl.Pinf, mdl.h)]
    cm = _lib.Model()
    cm.D, cm.N, cm.bz, cm.bg = mdl.D, mdl.N, mdl.bz, mdl.bg
    cm.A, cm.Q, cm.Pinf, cm.h = [_lib.dptr(a) for a in arrs]
    Wf = np.asfortranarray(np.asarray(Wnmf, float))
    o = _lib.Outputs()
    bufs = {}
    if predict:
        names = ["Eft", "Varft", "lb", "ub", "MS", "MF", "maxDiffP"] + (["PS", "PF"] if debug_cov else [])
        shapes = dict(Eft=(T, M), Varft=(T, M), lb=(T, M), ub=(T, M), MS=(T, n), MF=(T, n), maxDiffP=(int(g_iter),),
                      PS=(T, n, n), PF=(T, n, n))
    else:
        names, shapes = ["edata"], dict(edata=(1,))
    for nm in names:
        bufs[nm] = np.zeros(shapes[nm])
        setattr(o, nm, _lib.dptr(bufs[nm]))
    yb = _lib.as_f64(yall)
    call = _lib.lib().nsagp_giekf if constrained is not None else _lib.lib().nsagp_giekf_carry
    status = call(C.byref(cm), Wf.ctypes.data_as(_lib.c_double_p), sigma2, int(g_iter), int(l_iter),
                  _lib.dptr(yb), T, _lib.MODE_PREDICT if predict else _lib.MODE_NLZ, C.byref(o))
    if not predict:
        if status == -5:                                                # NSAGP_ERR_NAN: the reference returns NaN (:417-427)
            return float("nan"), np.zeros(np.size(w))
        _lib.check(status)
        return float(bufs["edata"][0]), np.zeros(np.size(w))
    _lib.check(status)
    sel = lambda a: a.T[:, return_ind]
    out = dict(MS=bufs["MS"].T, MF=bufs["MF"].T, maxDiffP=bufs["maxDiffP"], R=np.zeros((M, T)))
    if debug_cov:
        out["PS"] = np.transpose(bufs["PS"], (2, 1, 0))                 # (n, n, T), column-major blocks on the device
        out["PF"] = np.transpose(bufs["PF"], (2, 1, 0))
    res = (sel(bufs["Eft"]), sel(bufs["Varft"]), None, sel(bufs["lb"]), sel(bufs["ub"]), out)
    return res[:max(nargout, 1)] if nargout < 6 else res
