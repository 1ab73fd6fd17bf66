"""Stationary probabilistic filter bank: the step before the EP path that initialises the subbands.

    get_disc_model            matlab/unifying_prob_tf/get_disc_model.m:1
    kernel_ss_kalmanFastFB    matlab/unifying_prob_tf/kernel_ss_kalmanFastFB.m:1

Same arguments and return values as the reference.  Host work is what the reference does once per call with MATLAB
built-ins (two ``dare`` solves, a few n-by-n products); the two loops over time run in the CUDA library (C ABI
``nsagp_fastfb``, csrc/fastfb.cuh).  There is no CPU path for them.
"""
import numpy as np
import scipy.linalg as sla

from . import _lib, ssmodel


def get_disc_model(lamx, varx, omega, D, kernel):
    """Discrete-time model of D quasi-periodic components (kernel x cosine).  Returns (A, Q, H, Pinf, K, tau1)."""
    lamx = np.asarray(lamx, float); varx = np.asarray(varx, float); omega = np.asarray(omega, float)
    scale = {"exp": 1.0, "matern32": 3.0 ** 0.5, "matern52": 5.0 ** 0.5}[kernel]
    order = ssmodel.kernel_order(kernel)
    n = 2 * order * D
    F = np.zeros((n, n)); L = np.zeros((n, 2 * D)); Qc = np.zeros((2 * D, 2 * D)); H = np.zeros((1, n)); Pinf = np.zeros((n, n))
    I2 = np.eye(2)
    for d in range(D):
        Fk, Lk, qk, Hk, Pk = ssmodel.kernel_sde(kernel, varx[d], scale / lamx[d])
        rot = np.array([[0.0, -omega[d]], [omega[d], 0.0]])
        sl = slice(2 * order * d, 2 * order * (d + 1))
        F[sl, sl] = np.kron(Fk, I2) + np.kron(np.eye(order), rot)
        L[sl, 2 * d:2 * d + 2] = np.kron(np.reshape(Lk, (-1, 1)), I2)
        Qc[2 * d:2 * d + 2, 2 * d:2 * d + 2] = float(np.ravel(qk)[0]) * I2
        H[0, sl] = np.kron(np.reshape(Hk, (1, -1)), np.array([[1.0, 0.0]]))
        Pinf[sl, sl] = np.kron(Pk, I2)
    A, Q = ssmodel.lti_disc(F, L, Qc, 1.0)
    return A, Q, H, Pinf, D * order, order


def kernel_ss_kalmanFastFB(A, Q, C, P0, K, vary, y, verbose=0, KF=0):
    """Infinite-horizon Kalman filter (KF = 1) / smoother (KF = 0) of the filter bank.
    Returns ``(lik, Xfin, Pfin)`` with Xfin [1, n, T] and Pfin [n, n, T] (one matrix replicated, as in the reference)."""
    A = _lib.as_f64(A); Q = _lib.as_f64(Q)
    y = _lib.as_f64(np.ravel(y))
    n, T = A.shape[0], y.size
    H = np.reshape(np.asarray(C, float), (1, n))
    R = float(vary)
    try:
        PP = sla.solve_discrete_are(A.T, H.T, Q, np.array([[R]]))                # dare(A',H',Q,R) (:50)
    except Exception as e:
        raise RuntimeError("Unstable DARE solution! (%s)" % e)                    # :55-57
    S = (H @ PP @ H.T).item() + R
    Kg = (PP @ H.T / S).ravel()
    AKHA = A - np.outer(Kg, H @ A)
    PF2 = PP - np.outer(Kg, H @ PP)
    HA = (H @ A).ravel()
    G = P = None
    if KF != 1:
        G = np.linalg.solve(PP.T, (PF2 @ A.T).T).T                                # PF2*A'/PP (:126)
        QQ = PF2 - G @ PP @ G.T
        QQ = (QQ + QQ.T) / 2
        P = sla.solve_discrete_lyapunov(G, QQ)                                    # dare(G',0,QQ) (:131)
    MS = np.empty((T, n))
    quad = np.zeros(1)
    fcol = lambda a: _lib.as_f64(np.asfortranarray(a).ravel(order="F"))
    Af, Ff, Gf = fcol(A), fcol(AKHA), (fcol(G) if G is not None else None)
    Kf, HAf = _lib.as_f64(Kg), _lib.as_f64(HA)
    _lib.check(_lib.lib().nsagp_fastfb(n, _lib.dptr(Af), _lib.dptr(Ff), _lib.dptr(Kf), _lib.dptr(HAf), S,
                                       _lib.dptr(Gf) if Gf is not None else None, _lib.dptr(y), T, _lib.dptr(MS), _lib.dptr(quad)))
    lik = 0.5 * np.log(2 * np.pi) * T + 0.5 * np.log(S) * T + quad[0]             # :80, :99
    Pfin = np.empty((n, n, T))
    Pfin[:] = PF2[:, :, None]
    if KF != 1:
        Pfin[:, :, :T - 1] = P[:, :, None]                                        # :147 (the last step keeps PF2)
    return -lik, MS.T[None, :, :], Pfin
