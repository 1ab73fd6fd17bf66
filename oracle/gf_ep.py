"""Oracle: full-state Power-EP Kalman filter / RTS smoother (test infrastructure only).

Restates matlab/gf_ep_modulator_nmf.m (predict mode :92-352, nlZ mode :357-533)
and matlab/gf_ep_modulator_nmf_constraints.m (identical recursion; parameter
unpacking :75-110 and ``balance`` ON :115).  Dense n-by-n arithmetic in the
reference's operation order.  Indices are 0-based: MATLAB ``k == numel(yall)``
is ``k == T-1`` here, ``ep_damping(itt+1)`` is ``ep_damping[itt]`` with itt 1-based.
"""
import numpy as np

from . import ssmodel


class OracleNotPD(RuntimeError):
    """The reference would draw ``rand`` jitter here (gf_ep_modulator_nmf.m:219-223),
    which is not reproducible; the oracle raises instead."""


def merge_inputs(x, y, xt):
    """gf_ep_modulator_nmf.m:58-66 ([~,sort_ind,return_ind] = unique(xall,'first'))."""
    x = np.asarray(x, float).ravel()
    y = np.asarray(y, float).ravel()
    xt = np.zeros(0) if xt is None else np.asarray(xt, float).ravel()
    xall = np.concatenate([x, xt])
    yall = np.concatenate([y, np.full(xt.size, np.nan)])
    _, sort_ind, return_ind = np.unique(xall, return_index=True, return_inverse=True)
    yall = yall[sort_ind]
    return_ind = return_ind[xall.size - xt.size:]
    return yall, return_ind


def _chol_lower(S):
    try:
        return np.linalg.cholesky(S)
    except np.linalg.LinAlgError as e:
        raise OracleNotPD(str(e))


def _filter_update_predict(m, P, H, fmu, W, HPH, ttau_k, tnu_k):
    """gf_ep_modulator_nmf.m:158-176 (per-subset two-branch update)."""
    ii = (ttau_k == 0)
    if np.any(ii):
        z = ttau_k[ii] * HPH[ii] + 1
        K = W[:, ii] * (ttau_k[ii] / z)[None, :]
        v = ttau_k[ii] * fmu[ii] - tnu_k[ii]
        m = m - W[:, ii] @ (v / z)
        P = P - K @ W[:, ii].T
    if np.any(~ii):
        K = W[:, ~ii] / (HPH[~ii] + 1.0 / ttau_k[~ii])[None, :]
        v = tnu_k[~ii] / ttau_k[~ii] - fmu[~ii]
        m = m + K @ v
        P = P - K @ H[~ii, :] @ P
    return m, P


def _smoother_step(A, Q, MSk, PSk, m, P):
    """gf_ep_modulator_nmf.m:209-230."""
    PSkp = A @ PSk @ A.T + Q
    L = _chol_lower(PSkp)
    # G = PSk*A'/L'/L
    G = np.linalg.solve(L.T, np.linalg.solve(L, (PSk @ A.T).T)).T
    m = MSk + G @ (m - A @ MSk)
    P = PSk + G @ (P - PSkp) @ G.T
    return m, P


def _ep_site_update(mom, lik_param, Wnmf, H, m, P, ttau_k, tnu_k, ep_fraction, ep_damp, yall, k):
    """gf_ep_modulator_nmf.m:241-259: cavity, moments, damped Power-EP update on v_cav>0."""
    m_marginal = H @ m
    v_marginal = np.diag(H @ P @ H.T).copy()
    with np.errstate(all="ignore"):
        v_cav = 1.0 / (1.0 / v_marginal - ep_fraction * ttau_k)
        m_cav = v_cav * (m_marginal / v_marginal - ep_fraction * tnu_k)
        upd = v_cav > 0
        lZk, dlZ, d2lZ = mom(lik_param, m_cav, v_cav, Wnmf, ep_fraction, yall, k)
        ttau_new = ttau_k.copy()
        tnu_new = tnu_k.copy()
        ttau_new[upd] = (1 - ep_damp * ep_fraction) * ttau_k[upd] + \
            ep_damp * (-d2lZ[upd] / (1 + d2lZ[upd] * v_cav[upd]))
        tnu_new[upd] = (1 - ep_damp * ep_fraction) * tnu_k[upd] + \
            ep_damp * ((dlZ[upd] - m_cav[upd] * d2lZ[upd]) / (1 + d2lZ[upd] * v_cav[upd]))
    return lZk, ttau_new, tnu_new, v_cav


def _max0(a):
    """MATLAB max(a,0): NaN -> 0."""
    return np.where(a > 0, a, 0.0)


def model_from_params(lik_param, param1, param2, ss, x, kernel1, kernel2, do_balance):
    F, L, Qc, H, Pinf = ss(x, param1, param2, kernel1, kernel2)[:5]
    if do_balance:
        F, L, H, Pinf, _ = ssmodel.balance_ss(F, L, H, Pinf)
    A, Q = ssmodel.lti_disc(F, L, Qc, 1.0)
    return A, Q, H, Pinf


def gf_ep_core(A, Q, H, Pinf, lik_param, Wnmf, yall, mom, ep_fraction, ep_damping, ep_itts,
               predict, return_ind=None, want_cov=False, predict_first=False):
    """The two loops of gf_ep_modulator_nmf.m given the discrete model.

    predict=True  -> (Eft, Varft, lb, ub, out)      (:113-352)
    predict=False -> (edata, out)                    (:384-531)
    """
    n = A.shape[0]
    T = yall.size
    M = H.shape[0]
    ep_damping = np.atleast_1d(np.asarray(ep_damping, float))
    MS = np.zeros((n, T)); PS = np.zeros((n, n, T))
    ttau = np.zeros((M, T)); tnu = np.zeros((M, T))
    lZ = np.zeros(T); R = np.zeros((M, T))
    nlZ = np.zeros(ep_itts)
    out = {}
    n_negcav = 0
    ep_damp = ep_damping[0]
    maxDiffM_hist, maxDiffP_hist = [], []
    for itt in range(1, ep_itts + 1):
        m = np.zeros(n); P = Pinf.copy()
        maxDiffP = 0.0; maxDiffM = 0.0
        if predict:
            PSP = PS.copy(); MSP = MS.copy()
        run_filter = predict or itt == 1 or itt < ep_itts           # :396
        if run_filter:
            for k in range(T):
                if k > 0 or (predict and predict_first):          # gf_ep_modulator.m:131-133 predicts at k = 1 too
                    m = A @ m
                    P = A @ P @ A.T + Q
                if not np.isnan(yall[k]):
                    fmu = H @ m; W = P @ H.T; HPH = np.diag(H @ P @ H.T).copy()
                    if (not predict) and HPH.min() <= 0:
                        raise RuntimeError("fs2<=0: the reference drops into `keyboard` here (:408-410)")
                    if itt == 1 or k == T - 1:
                        lZ[k], dlZ, d2lZ = mom(lik_param, fmu, HPH, Wnmf, 1, yall, k)
                        with np.errstate(all="ignore"):
                            ttau[:, k] = (1 - ep_damp) * ttau[:, k] + ep_damp * (-d2lZ / (1 + d2lZ * HPH))
                            tnu[:, k] = (1 - ep_damp) * tnu[:, k] + ep_damp * ((dlZ - fmu * d2lZ) / (1 + d2lZ * HPH))
                        if predict:
                            ttau[:, k] = _max0(ttau[:, k])                      # :151
                            with np.errstate(divide="ignore"):
                                R[:, k] = 1.0 / ttau[:, k]                       # :154
                    if predict:
                        with np.errstate(all="ignore"):
                            m, P = _filter_update_predict(m, P, H, fmu, W, HPH, ttau[:, k], tnu[:, k])
                    else:
                        ttau[:, k] = _max0(ttau[:, k])                          # :425
                        with np.errstate(all="ignore"):
                            if ttau[:, k].min() == 0:                            # :428-433
                                z = ttau[:, k] * HPH + 1
                                K = W * (ttau[:, k] / z)[None, :]
                                v = ttau[:, k] * fmu - tnu[:, k]
                                m = m - W @ (v / z)
                                P = P - K @ W.T
                            else:                                                # :435-438
                                K = W / (HPH + 1.0 / ttau[:, k])[None, :]
                                v = tnu[:, k] / ttau[:, k] - fmu
                                m = m + K @ v
                                P = P - K @ H @ P
                if predict or itt < ep_itts:
                    MS[:, k] = m; PS[:, :, k] = P
        if itt == 1 and predict:
            nlZ[0] = -np.sum(lZ)
        if predict:
            out.update(tnu=tnu.copy(), ttau=ttau.copy(), lZ=lZ.copy(), R=R.copy(), MF=MS.copy())
            if want_cov:
                out["PF"] = PS.copy()
        run_smoother = predict or itt < ep_itts                      # :451
        if itt < ep_itts:
            ep_damp = ep_damping[itt]                                 # ep_damping(itt+1)
        if run_smoother:
            for k in range(T - 2, -1, -1):
                m, P = _smoother_step(A, Q, MS[:, k], PS[:, :, k], m, P)
                MS[:, k] = m; PS[:, :, k] = P
                if itt < ep_itts and not np.isnan(yall[k]):
                    lZ[k], ttau[:, k], tnu[:, k], v_cav = _ep_site_update(
                        mom, lik_param, Wnmf, H, m, P, ttau[:, k], tnu[:, k], ep_fraction, ep_damp, yall, k)
                    n_negcav += int(np.sum(~(v_cav > 0)))
                    if predict:
                        ttau[:, k] = _max0(ttau[:, k])                # :262
                        with np.errstate(divide="ignore"):
                            R[:, k] = 1.0 / ttau[:, k]                # :265
                if predict:
                    maxDiffM = max(maxDiffM, np.max(np.abs(H @ MSP[:, k] - H @ m)))
                    maxDiffP = max(maxDiffP, np.max(np.abs(H @ PSP[:, :, k] @ H.T - H @ P @ H.T)))
        if predict and itt < ep_itts:
            nlZ[itt] = -np.sum(lZ)
        maxDiffM_hist.append(maxDiffM); maxDiffP_hist.append(maxDiffP)
    out.update(tnu=tnu, ttau=ttau, lZ=lZ, R=R, nlZ=nlZ, n_negcav=n_negcav,
               maxDiffM=np.array(maxDiffM_hist), maxDiffP=np.array(maxDiffP_hist))
    if not predict:
        return -np.sum(lZ), out                                       # :525
    out["MS"] = MS
    if want_cov:
        out["PS"] = PS
    MSr = MS[:, return_ind]
    Eft = H @ MSr
    Varft = np.stack([np.diag(H @ PS[:, :, k] @ H.T) for k in return_ind], axis=1)
    with np.errstate(invalid="ignore"):
        lb = Eft - 1.96 * np.sqrt(Varft)
        ub = Eft + 1.96 * np.sqrt(Varft)
    return Eft, Varft, lb, ub, out


def gf_ep_modulator_nmf(w, x, y, ss, mom, xt, kernel1, kernel2, num_lik_params, D, N,
                        ep_fraction, ep_damping, ep_itts, want_cov=False):
    """gf_ep_modulator_nmf.m:1.  xt None/empty -> (nlZ, grad zeros); else
    (Eft, Varft, Covft=None, lb, ub, out)."""
    yall, return_ind = merge_inputs(x, y, xt)
    lik_param, param1, param2, Wnmf = ssmodel.unpack_log(w, num_lik_params, D, N)
    A, Q, H, Pinf = model_from_params(lik_param, param1, param2, ss, x, kernel1, kernel2, do_balance=False)  # :80
    predict = xt is not None and np.size(xt) > 0
    if predict:
        Eft, Varft, lb, ub, out = gf_ep_core(A, Q, H, Pinf, lik_param, Wnmf, yall, mom, ep_fraction,
                                             ep_damping, ep_itts, True, return_ind, want_cov)
        return Eft, Varft, None, lb, ub, out
    edata, out = gf_ep_core(A, Q, H, Pinf, lik_param, Wnmf, yall, mom, ep_fraction, ep_damping, ep_itts, False)
    return edata, np.zeros(np.size(w))                                # :363,531


def gf_ep_modulator_nmf_constraints(w, x, y, ss, mom, xt, kernel1, kernel2, num_lik_params, D, N,
                                    ep_fraction, ep_damping, ep_itts, constraints, w_fixed, tune_hypers,
                                    want_cov=False):
    """gf_ep_modulator_nmf_constraints.m:1 (sigmoid-constrained parameters, balance ON)."""
    yall, return_ind = merge_inputs(x, y, xt)
    lik_param, param1, param2, Wnmf = ssmodel.unpack_constraints(w, num_lik_params, D, N, constraints,
                                                                 w_fixed, tune_hypers)
    A, Q, H, Pinf = model_from_params(lik_param, param1, param2, ss, x, kernel1, kernel2, do_balance=True)   # :115
    predict = xt is not None and np.size(xt) > 0
    if predict:
        Eft, Varft, lb, ub, out = gf_ep_core(A, Q, H, Pinf, lik_param, Wnmf, yall, mom, ep_fraction,
                                             ep_damping, ep_itts, True, return_ind, want_cov)
        return Eft, Varft, None, lb, ub, out
    edata, out = gf_ep_core(A, Q, H, Pinf, lik_param, Wnmf, yall, mom, ep_fraction, ep_damping, ep_itts, False)
    return edata, np.zeros(np.size(w))


def gf_ep_modulator(w, x, y, ss, mom, xt, kernel1, kernel2, num_lik_params, ep_fraction, ep_damping, ep_itts,
                    want_cov=False):
    """gf_ep_modulator.m:1 -- the model WITHOUT the NMF weights: D carrier x modulator pairs, y = sum_d z_d link(g_d)
    (demo_toy_modulators.m).  Same two loops as gf_ep_modulator_nmf.m with W = I (the line-by-line difference is the
    ``mom`` signature, :138,:227), on the BALANCED model (:75-81), the prediction also at k = 1 in predict mode
    (:131-133; a no-op on the stationary initial state), ``ss(x, param, kernel1, kernel2)`` with the five parameter
    groups in one vector (:69-72) and ``mom(hyp, mu, s2, ep_frac, yall, k)`` without W."""
    yall, return_ind = merge_inputs(x, y, xt)
    w = np.asarray(w, float).ravel()
    lik_param = w[:num_lik_params]
    param = np.exp(w[num_lik_params:])
    F, L, Qc, H, Pinf = ss(x, param, kernel1, kernel2)[:5]
    F, L, H, Pinf, _ = ssmodel.balance_ss(F, L, H, Pinf)
    A, Q = ssmodel.lti_disc(F, L, Qc, 1.0)
    pairs = H.shape[0] // 2
    Wid = np.eye(pairs)
    mom_w = lambda hyp, mu, s2, nmfW, ep_frac, yall_, k: mom(hyp, mu, s2, ep_frac, yall_, k)
    predict = xt is not None and np.size(xt) > 0
    if predict:
        Eft, Varft, lb, ub, out = gf_ep_core(A, Q, H, Pinf, lik_param, Wid, yall, mom_w, ep_fraction, ep_damping, ep_itts,
                                             True, return_ind, want_cov, predict_first=True)
        return Eft, Varft, None, lb, ub, out
    edata, out = gf_ep_core(A, Q, H, Pinf, lik_param, Wid, yall, mom_w, ep_fraction, ep_damping, ep_itts, False)
    return edata, np.zeros(np.size(w))
