"""Oracle: state-space model construction (test infrastructure only).

Restates matlab/ss_modulators_nmf.m (only the F,L,Qc,H,Pinf outputs),
matlab/unifying_prob_tf/cf_{exp,matern32,matern52,matern72}_to_ss.m,
matlab/unifying_prob_tf/lti_disc.m, the ``balance`` stanza of
matlab/ihgp_ep_modulator_nmf.m:81-87, and matlab/{sigmoid,inv_sigmoid,
lambda_map,unpack_params}.m.  Dense n-by-n matrices throughout, as in MATLAB.
"""
import math

import numpy as np
import scipy.linalg as sla


def cf_exp_to_ss(magnSigma2=1.0, lengthScale=1.0):
    """cf_exp_to_ss.m:90-110."""
    F = np.array([[-1.0 / lengthScale]])
    L = np.array([[1.0]])
    Qc = np.array([[2.0 * magnSigma2 / lengthScale]])
    H = np.array([[1.0]])
    Pinf = np.array([[magnSigma2]])
    return F, L, Qc, H, Pinf


def cf_matern32_to_ss(magnSigma2=1.0, lengthScale=1.0):
    """cf_matern32_to_ss.m:90-117."""
    lam = math.sqrt(3) / lengthScale
    F = np.array([[0.0, 1.0], [-lam ** 2, -2 * lam]])
    L = np.array([[0.0], [1.0]])
    Qc = np.array([[12 * math.sqrt(3) / lengthScale ** 3 * magnSigma2]])
    H = np.array([[1.0, 0.0]])
    Pinf = np.array([[magnSigma2, 0.0], [0.0, 3 * magnSigma2 / lengthScale ** 2]])
    return F, L, Qc, H, Pinf


def cf_matern52_to_ss(magnSigma2=1.0, lengthScale=1.0):
    """cf_matern52_to_ss.m:90-123."""
    lam = math.sqrt(5) / lengthScale
    F = np.array([[0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [-lam ** 3, -3 * lam ** 2, -3 * lam]])
    L = np.array([[0.0], [0.0], [1.0]])
    Qc = np.array([[magnSigma2 * 400 * math.sqrt(5) / 3 / lengthScale ** 5]])
    H = np.array([[1.0, 0.0, 0.0]])
    kappa = 5 / 3 * magnSigma2 / lengthScale ** 2
    Pinf = np.array([[magnSigma2, 0.0, -kappa], [0.0, kappa, 0.0],
                     [-kappa, 0.0, 25 * magnSigma2 / lengthScale ** 4]])
    return F, L, Qc, H, Pinf


def cf_matern72_to_ss(magnSigma2=1.0, lengthScale=1.0):
    """cf_matern72_to_ss.m:90-128."""
    lam = math.sqrt(7) / lengthScale
    F = np.array([[0.0, 1.0, 0.0, 0.0], [0.0, 0.0, 1.0, 0.0], [0.0, 0.0, 0.0, 1.0],
                  [-lam ** 4, -4 * lam ** 3, -6 * lam ** 2, -4 * lam]])
    L = np.array([[0.0], [0.0], [0.0], [1.0]])
    Qc = np.array([[magnSigma2 * 10976 * math.sqrt(7) / 5 / lengthScale ** 7]])
    H = np.array([[1.0, 0.0, 0.0, 0.0]])
    kappa = 7 / 5 * magnSigma2 / lengthScale ** 2
    kappa2 = 9.8 * magnSigma2 / lengthScale ** 4
    Pinf = np.array([[magnSigma2, 0.0, -kappa, 0.0], [0.0, kappa, 0.0, -kappa2],
                     [-kappa, 0.0, kappa2, 0.0], [0.0, -kappa2, 0.0, 343 * magnSigma2 / lengthScale ** 6]])
    return F, L, Qc, H, Pinf


_CF = {"exp": cf_exp_to_ss, "matern32": cf_matern32_to_ss,
       "matern52": cf_matern52_to_ss, "matern72": cf_matern72_to_ss}


def ss_modulators_nmf(w_subband, w_modulator, kernel1, kernel2):
    """ss_modulators_nmf.m:1-137 (derivative stacks omitted: the EP entry
    points discard them, gf_ep_modulator_nmf.m:78)."""
    w_subband = np.asarray(w_subband, float).ravel()
    w_modulator = np.asarray(w_modulator, float).ravel()
    D = w_subband.size // 3
    N = w_modulator.size // 2
    sig1, len1, omega = w_subband[:D], w_subband[D:2 * D], w_subband[2 * D:]
    sig2, len2 = w_modulator[:N], w_modulator[N:]
    cf1, cf2 = _CF[kernel1], _CF[kernel2]
    tau1 = cf1(1.0, 1.0)[0].shape[0]
    tau2 = 2

    # periodic subband (:23-78)
    F1s, L1s, Qc1s, H1s, P1s = [], [], [], [], []
    for d in range(D):
        F1d, L1d, Qc1d, H1d, P1d = cf1(sig1[d], len1[d])
        F1s.append(F1d); L1s.append(L1d); Qc1s.append(Qc1d); H1s.append(H1d); P1s.append(P1d)
    F1 = sla.block_diag(*F1s)
    L1 = np.vstack(L1s)                   # vertcat (:34)
    Qc1 = sla.block_diag(*Qc1s)
    H1 = sla.block_diag(*H1s)
    Pinf1 = sla.block_diag(*P1s)
    F_cos = sla.block_diag(*[np.array([[0.0, -omega[d]], [omega[d], 0.0]]) for d in range(D)])
    L_cos = np.eye(tau2 * D)
    F_cos_kron, L_sm, Qc_sm = [], [], []
    for d in range(D):
        i1 = slice(tau1 * d, tau1 * (d + 1))
        i2 = slice(tau2 * d, tau2 * (d + 1))
        F_cos_kron.append(np.kron(np.eye(tau1), F_cos[i2, i2]))
        L_sm.append(np.kron(L1[i1], L_cos[i2, i2]))
        Qc_sm.append(np.kron(Qc1[d:d + 1, d:d + 1], L_cos[i2, i2]))
    F_sm = np.kron(F1, np.eye(tau2)) + sla.block_diag(*F_cos_kron)
    L_sm = sla.block_diag(*L_sm)
    Qc_sm = sla.block_diag(*Qc_sm)
    H_sm = np.kron(H1, np.array([[1.0, 0.0]]))
    Pinf_sm = np.kron(Pinf1, np.eye(tau2))

    # slow varying modulator (:82-117)
    F2s, L2s, Qc2s, H2s, P2s = [], [], [], [], []
    for d in range(N):
        F2d, L2d, Qc2d, H2d, P2d = cf2(sig2[d], len2[d])
        F2s.append(F2d); L2s.append(L2d); Qc2s.append(Qc2d); H2s.append(H2d); P2s.append(P2d)

    # combine (:121-132)
    F = sla.block_diag(F_sm, *F2s)
    L = sla.block_diag(L_sm, *L2s)
    Qc = sla.block_diag(Qc_sm, *Qc2s)
    H = sla.block_diag(H_sm, *H2s)
    Pinf = sla.block_diag(Pinf_sm, *P2s)
    return F, L, Qc, H, Pinf


def ss_modulators(w, kernel1, kernel2):
    """ss_modulators.m:1-134: D carrier x modulator pairs, one parameter vector [var1; len1; omega; var2; len2] (:5-10).
    Line by line the construction of ss_modulators_nmf.m with N = D."""
    w = np.asarray(w, float).ravel()
    D = w.size // 5
    return ss_modulators_nmf(w[:3 * D], w[3 * D:], kernel1, kernel2)


def _cf_derivs(kernel, magnSigma2, lengthScale):
    """Derivative stacks (dF, dQc, dPinf), last axis = (magnSigma2, lengthScale), of one covariance function:
    cf_exp_to_ss.m:70-92, cf_matern32_to_ss.m:78-104, cf_matern52_to_ss.m:82-112, cf_matern72_to_ss.m:84-116."""
    s2, l = float(magnSigma2), float(lengthScale)
    F, L, Qc, H, Pinf = _CF[kernel](s2, l)
    t = F.shape[0]
    dF = np.zeros((t, t, 2)); dQc = np.zeros((1, 1, 2)); dPinf = np.zeros((t, t, 2))
    if kernel == "exp":
        dF[0, 0, 1] = 1 / l ** 2
        dQc[0, 0, 0] = 2 / l; dQc[0, 0, 1] = -2 * s2 / l ** 2
        dPinf[0, 0, 0] = 1.0
    elif kernel == "matern32":
        dF[1, :, 1] = [6 / l ** 3, 2 * math.sqrt(3) / l ** 2]
        dQc[0, 0, 0] = 12 * math.sqrt(3) / l ** 3; dQc[0, 0, 1] = -3 * 12 * math.sqrt(3) / l ** 4 * s2
        dPinf[:, :, 0] = [[1, 0], [0, 3 / l ** 2]]
        dPinf[:, :, 1] = [[0, 0], [0, -6 * s2 / l ** 3]]
    elif kernel == "matern52":
        dF[2, :, 1] = [15 * math.sqrt(5) / l ** 4, 30 / l ** 3, 3 * math.sqrt(5) / l ** 2]
        dQc[0, 0, 0] = 400 * math.sqrt(5) / 3 / l ** 5; dQc[0, 0, 1] = -s2 * 2000 * math.sqrt(5) / 3 / l ** 6
        kappa = 5 / 3 * s2 / l ** 2
        k2 = -2 * kappa / l
        dPinf[:, :, 0] = Pinf / s2
        dPinf[:, :, 1] = [[0, 0, -k2], [0, k2, 0], [-k2, 0, -100 * s2 / l ** 5]]
    elif kernel == "matern72":
        dF[3, :, 1] = [196 / l ** 5, 84 * math.sqrt(7) / l ** 4, 84 / l ** 3, 4 * math.sqrt(7) / l ** 2]
        dQc[0, 0, 0] = 10976 * math.sqrt(7) / 5 / l ** 7; dQc[0, 0, 1] = -s2 * 76832 * math.sqrt(7) / 5 / l ** 8
        dPinf[:, :, 0] = Pinf / s2
        dl = np.zeros(16); pv = Pinf.reshape(-1, order="F")         # MATLAB linear (column-major) indexing, :103-106
        dl[2::3] = -2 * pv[2::3] / l
        dl[1::3] = -4 * pv[1::3] / l
        dl[-1] = -3 * 2 * pv[-1] / l
        dPinf[:, :, 1] = dl.reshape((4, 4), order="F")
    else:
        raise ValueError(kernel)
    return dF, dQc, dPinf


def ss_modulators_nmf_derivs(w_subband, w_modulator, kernel1, kernel2):
    """Derivative stacks of ss_modulators_nmf.m (:25-47, :60-78, :84-110, :127-129): (dF, dQc, dPinf), each with
    3D + 2N slices ordered [d/dsig1 (D), d/dlen1 (D), d/domega (D), d/dsig2 (N), d/dlen2 (N)]."""
    w_subband = np.asarray(w_subband, float).ravel()
    w_modulator = np.asarray(w_modulator, float).ravel()
    D = w_subband.size // 3
    N = w_modulator.size // 2
    sig1, len1 = w_subband[:D], w_subband[D:2 * D]
    sig2, len2 = w_modulator[:N], w_modulator[N:]
    tau1 = _CF[kernel1](1.0, 1.0)[0].shape[0]
    tau3 = _CF[kernel2](1.0, 1.0)[0].shape[0]
    tau2 = 2
    nz, ng = tau1 * tau2 * D, tau3 * N
    n, nq = nz + ng, tau2 * D + N
    P = 3 * D + 2 * N
    dF = np.zeros((n, n, P)); dQc = np.zeros((nq, nq, P)); dPinf = np.zeros((n, n, P))
    I2 = np.eye(tau2)
    for d in range(D):
        dFd, dQd, dPd = _cf_derivs(kernel1, sig1[d], len1[d])
        z = slice(d * tau1 * tau2, (d + 1) * tau1 * tau2)
        q = slice(d * tau2, (d + 1) * tau2)
        for which, p in ((0, d), (1, D + d)):                         # kron(dF1, eye(tau2)), :62-69
            dF[z, z, p] = np.kron(dFd[:, :, which], I2)
            dQc[q, q, p] = np.kron(dQd[:, :, which], I2)
            dPinf[z, z, p] = np.kron(dPd[:, :, which], I2)
        dF[z, z, 2 * D + d] = np.kron(np.eye(tau1), np.array([[0.0, -1.0], [1.0, 0.0]]))      # d/domega, :43-58
    for j in range(N):
        dFd, dQd, dPd = _cf_derivs(kernel2, sig2[j], len2[j])
        g = slice(nz + j * tau3, nz + (j + 1) * tau3)
        for which, p in ((0, 3 * D + j), (1, 3 * D + N + j)):
            dF[g, g, p] = dFd[:, :, which]
            dQc[tau2 * D + j, tau2 * D + j, p] = dQd[0, 0, which]
            dPinf[g, g, p] = dPd[:, :, which]
    return dF, dQc, dPinf


def lti_disc(F, L=None, Q=None, dt=1.0):
    """lti_disc.m:60-82: A = expm(F dt); Q by matrix-fraction decomposition."""
    n = F.shape[0]
    if L is None:
        L = np.eye(n)
    if Q is None:
        Q = np.zeros((n, n))
    A = sla.expm(F * dt)
    Phi = np.block([[F, L @ Q @ L.T], [np.zeros((n, n)), -F.T]])
    AB = sla.expm(Phi * dt) @ np.vstack([np.zeros((n, n)), np.eye(n)])
    # AB(1:n,:)/AB(n+1:2n,:)  ==  AB1 * inv(AB2)
    Qd = np.linalg.solve(AB[n:, :].T, AB[:n, :].T).T
    return A, Qd


def balance_ss(F, L, H, Pinf):
    """ihgp_ep_modulator_nmf.m:81-87:
    [T,F]=balance(F); L=T\\L; H=H*T; LL=T\\chol(Pinf,'lower'); Pinf=LL*LL'."""
    Fb, T = sla.matrix_balance(F, permute=True, scale=True, separate=False)
    Lb = np.linalg.solve(T, L)
    Hb = H @ T
    LL = np.linalg.solve(T, np.linalg.cholesky(Pinf))
    return Fb, Lb, Hb, LL @ LL.T, T


def sigmoid(x, sig_range=(0.0, 20.0), c=0.0, a=1.0):
    """sigmoid.m:17-19."""
    lo, up = sig_range[0], sig_range[-1]
    return (up - lo) / (1 + np.exp(-a * (np.asarray(x, float) - c))) + lo


def inv_sigmoid(y, sig_range=(0.0, 20.0), c=0.0, a=1.0):
    """inv_sigmoid.m:17-23."""
    lo, up = sig_range[0], sig_range[-1]
    y = np.asarray(y, float)
    if np.any((up - y) / (y - lo) <= 0):
        raise ValueError("Error with inverse sigmoid transformation: parameter outside of user specified range")
    return c - np.log((up - y) / (y - lo)) / a


def unpack_log(w, num_lik_params, D, N):
    """gf_ep_modulator_nmf.m:72-75."""
    w = np.asarray(w, float).ravel()
    lik_param = w[:num_lik_params]
    param1 = np.exp(w[num_lik_params:num_lik_params + 3 * D])
    param2 = np.exp(w[num_lik_params + 3 * D:num_lik_params + 3 * D + 2 * N])
    Wnmf = np.exp(w[num_lik_params + 3 * D + 2 * N:]).reshape((D, N), order="F")
    return lik_param, param1, param2, Wnmf


def unpack_constraints(w, num_lik_params, D, N, constraints, w_fixed, tune_hypers):
    """gf_ep_modulator_nmf_constraints.m:75-110."""
    w = np.asarray(w, float).ravel()
    w_fixed = np.asarray(w_fixed, float).ravel()
    constraints = np.asarray(constraints, float)
    w_ind = 0
    wf_ind = 0
    if tune_hypers[0]:
        lik_param = w[:num_lik_params]; w_ind += num_lik_params
    else:
        lik_param = w_fixed[:num_lik_params]; wf_ind += num_lik_params
    param1, param2 = [], []
    for i in range(2, 7):               # MATLAB i = 2..6
        cnt = D if i <= 4 else N
        dst = param1 if i <= 4 else param2
        if tune_hypers[i - 1]:
            dst.append(sigmoid(w[w_ind:w_ind + cnt], constraints[i - 2])); w_ind += cnt
        else:
            dst.append(sigmoid(w_fixed[wf_ind:wf_ind + cnt], constraints[i - 2])); wf_ind += cnt
    if tune_hypers[6]:
        Wnmf = sigmoid(w[w_ind:], constraints[5]).reshape((D, N), order="F")
    else:
        Wnmf = sigmoid(w_fixed[wf_ind:], constraints[5]).reshape((D, N), order="F")
    return lik_param, np.concatenate(param1), np.concatenate(param2), Wnmf


def lambda_map(lin, kernel):
    """lambda_map.m:1-15."""
    c = {"exp": 1.0, "matern32": math.sqrt(3), "matern52": math.sqrt(5), "matern72": math.sqrt(7)}[kernel]
    return c / np.asarray(lin, float)
