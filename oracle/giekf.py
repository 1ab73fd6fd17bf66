"""Oracle: globally-iterated EKF comparison variant (test infrastructure only).

Restates matlab/gf_giekf_modulator_nmf_constraints.m (predict mode :144-327;
energy of the nlZ mode :332-484 with GradObj='off', which is how every caller
runs it: experiments/train_model.m:226,239-240) and matlab/iekf_update1.m:110-117.
Measurement model (:490-513): h(x) = (H_z x)' W softplus(H_g x).
"""
import math

import numpy as np
import scipy.linalg as sla

from . import ssmodel
from .gf_ep import merge_inputs, _smoother_step


def _linkf(x):
    return np.log(1 + np.exp(x))


def _dlinkf(x):
    return np.exp(x) / (np.exp(x) + 1)


def funh(x, H, D, N, W):
    """gf_giekf_modulator_nmf_constraints.m:490-494."""
    return float((H[:D] @ x) @ W @ _linkf(H[D:D + N] @ x))


def funhd(x, H, D, N, W):
    """gf_giekf_modulator_nmf_constraints.m:497-503 (row Jacobian, 1 x n)."""
    g = H[D:D + N] @ x
    partials = np.concatenate([W @ _linkf(g), ((H[:D] @ x) @ W) * _dlinkf(g)])
    return partials @ H


def iekf_update1(M, P, y, D, N, H, Wnmf, R, iters):
    """iekf_update1.m:110-117 (V = I): relinearise at the *current* mean, no
    re-centering on the prior mean; covariance from the last linearisation."""
    for _ in range(iters):
        H_ = funhd(M, H, D, N, Wnmf)
        MU = funh(M, H, D, N, Wnmf)
        S = R + H_ @ P @ H_
        K = P @ H_ / S
        M = M + K * (y - MU)
    P = P - np.outer(K * S, K)
    return M, P


def giekf_core(A, Q, H, Pinf, sigma2, Wnmf, yall, D, N, g_iter, l_iter, return_ind, want_cov=False, carry_cov=False):
    """Predict mode, gf_giekf_modulator_nmf_constraints.m:144-327.  ``carry_cov``: the file without
    constraints initialises (m, P) on the first global iteration only (gf_giekf_modulator_nmf.m:127-131)."""
    n = A.shape[0]; T = yall.size
    MS = np.zeros((n, T)); PS = np.zeros((n, n, T))
    out = {}
    maxDiffP_hist = []
    m = np.zeros(n)
    P = Pinf.copy()
    for itt in range(1, g_iter + 1):
        if not carry_cov:
            P = Pinf.copy()                  # m carries over, P is reset (:165-168)
        maxDiffP = 0.0
        PSP = PS.copy()
        for k in range(T):
            if k > 0:
                m = A @ m
                P = A @ P @ A.T + Q
            if not np.isnan(yall[k]):
                m, P = iekf_update1(m, P, yall[k], D, N, H, Wnmf, sigma2, l_iter)
            MS[:, k] = m; PS[:, :, k] = P
        out.update(MF=MS.copy())
        if want_cov:
            out["PF"] = PS.copy()
        for k in range(T - 2, -1, -1):
            m, P = _smoother_step(A, Q, MS[:, k], PS[:, :, k], m, P)
            MS[:, k] = m; PS[:, :, k] = P
            maxDiffP = max(maxDiffP, np.max(np.abs(H @ PSP[:, :, k] @ H.T - H @ P @ H.T)))
        maxDiffP_hist.append(maxDiffP)
    out.update(MS=MS, maxDiffP=np.array(maxDiffP_hist))
    if want_cov:
        out["PS"] = PS
    Eft = H @ MS[:, return_ind]
    Varft = np.stack([np.diag(H @ PS[:, :, k] @ H.T) for k in return_ind], axis=1)
    with np.errstate(invalid="ignore"):
        lb = Eft - 1.96 * np.sqrt(Varft); ub = Eft + 1.96 * np.sqrt(Varft)
    return Eft, Varft, lb, ub, out


def giekf_energy(F, H, Pinf, sigma2, Wnmf, yall, D, N):
    """nlZ mode energy, gf_giekf_modulator_nmf_constraints.m:376-468 with nparam = 0."""
    A = sla.expm(F)
    Q = Pinf - A @ Pinf @ A.T
    m = np.zeros(F.shape[0]); P = Pinf.copy()
    edata = 0.0
    for k in range(yall.size):
        m = A @ m
        P = A @ P @ A.T + Q
        mu = funh(m, H, D, N, Wnmf)
        JH = funhd(m, H, D, N, Wnmf)
        S = JH @ P @ JH + sigma2
        if not S > 0:
            return math.nan
        LS = math.sqrt(S)
        K = P @ JH / S
        v = yall[k] - mu
        edata += 0.5 * math.log(2 * math.pi) + math.log(LS) + 0.5 * v * v / S
        m = m + K * v
        P = P - np.outer(K * S, K)
    return edata


def gf_giekf_modulator_nmf_constraints(w, x, y, ss, mom, xt, kernel1, kernel2, num_lik_params, D, N,
                                       g_iter, l_iter, constraints, w_fixed, tune_hypers, want_cov=False):
    """gf_giekf_modulator_nmf_constraints.m:1 (``mom`` is ignored by the reference too)."""
    yall, return_ind = merge_inputs(x, y, xt)
    lik_param, param1, param2, Wnmf = ssmodel.unpack_constraints(w, num_lik_params, D, N, constraints,
                                                                 w_fixed, tune_hypers)
    F, L, Qc, H, Pinf = ss(x, param1, param2, kernel1, kernel2)[:5]
    F, L, H, Pinf, _ = ssmodel.balance_ss(F, L, H, Pinf)             # :113-120
    sigma2 = math.exp(float(np.asarray(lik_param).ravel()[0]))
    if xt is not None and np.size(xt) > 0:
        A, Q = ssmodel.lti_disc(F, L, Qc, 1.0)
        Eft, Varft, lb, ub, out = giekf_core(A, Q, H, Pinf, sigma2, Wnmf, yall, D, N, g_iter, l_iter,
                                             return_ind, want_cov)
        return Eft, Varft, None, lb, ub, out
    return giekf_energy(F, H, Pinf, sigma2, Wnmf, yall, D, N), np.zeros(np.size(w))


def gf_giekf_modulator_nmf(w, x, y, ss, mom, xt, kernel1, kernel2, num_lik_params, D, N, g_iter, l_iter, want_cov=False):
    """gf_giekf_modulator_nmf.m:1 (GradObj = 'off'): log-scale parameters (:70-73), balanced model (:78-85),
    (m, P) carried across the global iterations (:127-131)."""
    yall, return_ind = merge_inputs(x, y, xt)
    lik_param, param1, param2, Wnmf = ssmodel.unpack_log(w, num_lik_params, D, N)
    F, L, Qc, H, Pinf = ss(x, param1, param2, kernel1, kernel2)[:5]
    F, L, H, Pinf, _ = ssmodel.balance_ss(F, L, H, Pinf)
    sigma2 = math.exp(float(np.asarray(lik_param).ravel()[0]))
    if xt is not None and np.size(xt) > 0:
        A, Q = ssmodel.lti_disc(F, L, Qc, 1.0)
        Eft, Varft, lb, ub, out = giekf_core(A, Q, H, Pinf, sigma2, Wnmf, yall, D, N, g_iter, l_iter,
                                             return_ind, want_cov, carry_cov=True)
        return Eft, Varft, None, lb, ub, out
    return giekf_energy(F, H, Pinf, sigma2, Wnmf, yall, D, N), np.zeros(np.size(w))
