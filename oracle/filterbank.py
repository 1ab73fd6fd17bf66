"""Oracle: stationary Kalman filter / RTS smoother of the probabilistic filter bank (test infrastructure only).

Restates matlab/unifying_prob_tf/kernel_ss_kalmanFastFB.m:46-151 (dense arithmetic in the reference's operation
order) and matlab/unifying_prob_tf/get_disc_model.m:1-73.  MATLAB's ``dare(A',H',Q,R)`` is
``scipy.linalg.solve_discrete_are(A.T, H.T, Q, R)``.
"""
import numpy as np
import scipy.linalg as sla

from . import ssmodel


def get_disc_model(lamx, varx, omega, D, kernel):
    """get_disc_model.m:1-73 (exp / matern32 / matern52)."""
    lamx = np.asarray(lamx, float); varx = np.asarray(varx, float); omega = np.asarray(omega, float)
    scale = {"exp": 1.0, "matern32": np.sqrt(3.0), "matern52": np.sqrt(5.0)}[kernel]
    lengthScale = scale / lamx                                                    # :9-19
    cf = {"exp": ssmodel.cf_exp_to_ss, "matern32": ssmodel.cf_matern32_to_ss, "matern52": ssmodel.cf_matern52_to_ss}[kernel]
    F1s, L1s, Qc1, H1s, P1s = [], [], [], [], []
    for d in range(D):                                                            # :26-35
        F1d, L1d, Qc1d, H1d, Pinf1d = cf(varx[d], lengthScale[d])[:5]
        F1s.append(np.atleast_2d(F1d)); L1s.append(np.reshape(L1d, (-1, 1))); Qc1.append(float(np.ravel(Qc1d)[0]))
        H1s.append(np.reshape(H1d, (1, -1))); P1s.append(np.atleast_2d(Pinf1d))
    tau1 = L1s[0].shape[0]
    F1 = sla.block_diag(*F1s); H1 = np.hstack(H1s); Pinf1 = sla.block_diag(*P1s)
    I2 = np.eye(2)
    F2k, Ls, Qcs = [], [], []
    for d in range(D):                                                            # :54-62
        F2d = np.array([[0.0, -omega[d]], [omega[d], 0.0]])
        F2k.append(np.kron(np.eye(tau1), F2d))
        Ls.append(np.kron(L1s[d], I2))
        Qcs.append(np.kron(np.array([[Qc1[d]]]), I2))
    F = np.kron(F1, I2) + sla.block_diag(*F2k)                                    # :63
    L = sla.block_diag(*Ls)
    Qc = sla.block_diag(*Qcs)
    H = np.kron(H1, np.array([[1.0, 0.0]]))                                       # :64
    Pinf = np.kron(Pinf1, I2)                                                     # :65
    A, Q = ssmodel.lti_disc(F, L, Qc, 1.0)                                        # :70
    return A, Q, H, Pinf, D * tau1, tau1


def kernel_ss_kalmanFastFB(A, Q, C, P0, K, vary, y, verbose=0, KF=0):
    """kernel_ss_kalmanFastFB.m:1-166 -> (lik, Xfin [1, n, T], Pfin [n, n, T])."""
    y = np.asarray(y, float).ravel()
    T = y.size
    n = A.shape[0]
    H = np.reshape(C, (1, n))
    R = float(vary)
    m = np.zeros(n)
    PP = sla.solve_discrete_are(A.T, H.T, Q, np.array([[R]]))                     # :50
    S = (H @ PP @ H.T).item() + R                                                   # :53
    Kg = (PP @ H.T / S).ravel()                                                   # :60
    AKHA = A - np.outer(Kg, H @ A)                                                # :63
    MS = np.zeros((n, T))
    PS = np.zeros((n, n, T))
    PF2 = PP - np.outer(Kg, H @ PP)                                               # :76
    HA = (H @ A).ravel()
    lik = 0.5 * np.log(2 * np.pi) * T + 0.5 * np.log(S) * T                       # :80
    for k in range(T):                                                            # :83-112
        if not np.isnan(y[k]):
            v = y[k] - HA @ m
            m = AKHA @ m + Kg * y[k]
            lik = lik + 0.5 * v ** 2 / S
        else:
            m = A @ m
        MS[:, k] = m
        PS[:, :, k] = PF2
    if KF != 1:
        G = np.linalg.solve(PP.T, (PF2 @ A.T).T).T                                # :126  PF2*A'/PP
        QQ = PF2 - G @ PP @ G.T
        QQ = (QQ + QQ.T) / 2
        P = sla.solve_discrete_lyapunov(G, QQ)                                    # :131  dare(G',0,QQ)
        for k in range(T - 2, -1, -1):                                            # :134-148
            m = MS[:, k] + G @ (m - A @ MS[:, k])
            MS[:, k] = m
            PS[:, :, k] = P
    return -lik, MS[None, :, :], PS
