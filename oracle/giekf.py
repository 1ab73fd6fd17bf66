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


def funhd2(x, H, D, N, W):
    """gf_giekf_modulator_nmf.m:459-472 (n x n second derivative of h).  As in funhd above, the derivative with
    respect to the latent values is chained through H instead of being scattered to the columns where
    ``sum(H,1)==1`` (:452, :470): the two agree when the observed components of H are 1, and the scatter is
    ill-defined once balancing has rescaled them."""
    z = H[:D] @ x
    g = H[D:D + N] @ x
    dl = _dlinkf(g)
    Wd = W * dl[None, :]
    foo2 = np.block([[np.zeros((D, D)), Wd], [Wd.T, np.diag((z @ W) * dl * (1 - dl))]])
    return H.T @ foo2 @ H


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


def giekf_energy_grad(F, H, Pinf, dF, dPinf, sigma2, Wnmf, yall, D, N):
    """Energy and its sensitivity-equation gradient, gf_giekf_modulator_nmf.m:296-437 with GradObj = 'on'.
    dF / dPinf: n x n x nparam with the zero slice of the noise parameter first (:93-96).  Returns
    (edata, gdata) BEFORE the log-scale factor of :432-433."""
    d = F.shape[0]
    nparam = dF.shape[2]
    Z = np.zeros((d, d))
    m = np.zeros(d); P = Pinf.copy()
    dm = np.zeros((d, nparam)); dP = dPinf.copy()
    dR = np.zeros(nparam); dR[0] = 1.0
    AA = [sla.expm(np.block([[F, Z], [dF[:, :, j], F]])) for j in range(nparam)]      # :328-338
    edata = 0.0
    gdata = np.zeros(nparam)
    A = AA[0][:d, :d]
    Q = Pinf - A @ Pinf @ A.T
    for k in range(yall.size):
        for j in range(nparam):                                        # :344-369
            foo = AA[j] @ np.concatenate([m, dm[:, j]])
            mm = foo[:d]
            dm[:, j] = foo[d:]
            if j == 0:
                PP = A @ P @ A.T + Q
            dA = AA[j][d:, :d]
            dAPinfAt = dA @ Pinf @ A.T
            dQ = dPinf[:, :, j] - dAPinfAt - A @ dPinf[:, :, j] @ A.T - dAPinfAt.T
            dAPAt = dA @ P @ A.T
            dP[:, :, j] = dAPAt + A @ dP[:, :, j] @ A.T + dAPAt.T + dQ
        m = mm; P = PP                                                 # :372-373
        mu = funh(m, H, D, N, Wnmf)
        JH = funhd(m, H, D, N, Wnmf)
        dJH = funhd2(m, H, D, N, Wnmf)
        S = JH @ P @ JH + sigma2
        if not S > 0:                                                  # :384-395
            return math.nan, np.full(nparam, math.nan)
        HtiS = JH / S
        K = P @ HtiS
        v = yall[k] - mu
        vtiS = v / S
        for j in range(nparam):                                        # :405-423
            dmdJH = dm[:, j] @ dJH
            dS = dmdJH @ P @ JH + JH @ dP[:, :, j] @ JH + JH @ P @ dmdJH + dR[j]
            gdata[j] += 0.5 * dS / S - 0.5 * (JH @ dm[:, j]) * vtiS - 0.5 * vtiS * dS * vtiS - 0.5 * vtiS * (JH @ dm[:, j])
            dK = dP[:, :, j] @ HtiS + P @ dmdJH / S - P @ HtiS * dS / S
            dm[:, j] = dm[:, j] + dK * v - K * (JH @ dm[:, j])
            dKSKt = np.outer(dK * S, K)
            dP[:, :, j] = dP[:, :, j] - dKSKt - np.outer(K * dS, K) - dKSKt.T
        edata += 0.5 * math.log(2 * math.pi) + math.log(math.sqrt(S)) + 0.5 * vtiS * v
        m = m + K * v
        P = P - np.outer(K * S, K)
    return edata, gdata


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


def gf_giekf_modulator_nmf(w, x, y, ss, mom, xt, kernel1, kernel2, num_lik_params, D, N, g_iter, l_iter, want_cov=False,
                           GradObj="off", balance_derivatives=False):
    """gf_giekf_modulator_nmf.m:1: log-scale parameters (:70-73), balanced model (:78-85), (m, P) carried across
    the global iterations (:127-131).  GradObj = 'on' with xt empty: energy and analytic gradient (:296-437), where
    the reference balances F, L, H, Pinf but NOT dF / dPinf (its balancing loop is commented out, :82-84);
    ``balance_derivatives`` runs that loop (then the result is the derivative of the energy)."""
    yall, return_ind = merge_inputs(x, y, xt)
    lik_param, param1, param2, Wnmf = ssmodel.unpack_log(w, num_lik_params, D, N)
    res = ss(x, param1, param2, kernel1, kernel2)
    F, L, Qc, H, Pinf = res[:5]
    F, L, H, Pinf, Tb = ssmodel.balance_ss(F, L, H, Pinf)
    if GradObj != "off" and not (xt is not None and np.size(xt) > 0):
        dF, dPinf = np.asarray(res[5], float), np.asarray(res[7], float)
        if balance_derivatives:                                       # the commented-out loop, :82-84
            Ti = np.linalg.inv(Tb)
            dF = np.stack([Ti @ dF[:, :, j] @ Tb for j in range(dF.shape[2])], axis=2)
            dPinf = np.stack([Ti @ dPinf[:, :, j] @ Ti.T for j in range(dPinf.shape[2])], axis=2)
        n = F.shape[0]
        dF = np.concatenate([np.zeros((n, n, 1)), dF], axis=2)        # :93-96
        dPinf = np.concatenate([np.zeros((n, n, 1)), dPinf], axis=2)
        s2 = math.exp(float(np.asarray(lik_param).ravel()[0]))
        edata, gdata = giekf_energy_grad(F, H, Pinf, dF, dPinf, s2, Wnmf, yall, D, N)
        ww = np.asarray(w, float).ravel()[:np.size(w) - D * N]
        return edata, gdata * np.exp(ww)                              # :432-433
    sigma2 = math.exp(float(np.asarray(lik_param).ravel()[0]))
    if xt is not None and np.size(xt) > 0:
        A, Q = ssmodel.lti_disc(F, L, Qc, 1.0)
        Eft, Varft, lb, ub, out = giekf_core(A, Q, H, Pinf, sigma2, Wnmf, yall, D, N, g_iter, l_iter,
                                             return_ind, want_cov, carry_cov=True)
        return Eft, Varft, None, lb, ub, out
    return giekf_energy(F, H, Pinf, sigma2, Wnmf, yall, D, N), np.zeros(np.size(w))
