"""Pins for the filter-bank oracle (oracle/filterbank.py restates matlab/unifying_prob_tf/kernel_ss_kalmanFastFB.m):
away from the ends of a long signal the stationary filter / smoother must coincide with the ordinary time-varying
Kalman filter / RTS smoother of the same model, and the product's host-side model builder must agree with the oracle's."""
import importlib

import numpy as np

from conftest import rel_err


def _kalman_rts(A, Q, H, Pinf, R, y):
    n, T = A.shape[0], y.size
    m, P = np.zeros(n), Pinf.copy()
    MF, PF = np.zeros((n, T)), np.zeros((n, n, T))
    for k in range(T):
        m, P = A @ m, A @ P @ A.T + Q
        if not np.isnan(y[k]):
            S = (H @ P @ H.T).item() + R
            K = (P @ H.T / S).ravel()
            m = m + K * (y[k] - (H @ m).item())
            P = P - np.outer(K, H @ P)
        MF[:, k], PF[:, :, k] = m, P
    MS = MF.copy()
    for k in range(T - 2, -1, -1):
        Pp = A @ PF[:, :, k] @ A.T + Q
        G = np.linalg.solve(Pp.T, (PF[:, :, k] @ A.T).T).T
        MS[:, k] = MF[:, k] + G @ (MS[:, k + 1] - A @ MF[:, k])
    return MF, MS


def test_stationary_filterbank_equals_kalman_rts_in_the_interior():
    from oracle import filterbank as ofb
    rng = np.random.default_rng(1)
    D, T = 3, 1200
    lamx, varx, om = 1.0 / rng.uniform(20, 60, D), rng.uniform(0.3, 1.0, D), np.array([0.9, 0.5, 0.2])
    A, Q, H, Pinf, K, tau = ofb.get_disc_model(lamx, varx, om, D, "matern32")
    y = np.cos(0.5 * np.arange(T)) + 0.1 * rng.standard_normal(T)
    MF, MS = _kalman_rts(A, Q, H, Pinf, 0.01, y)
    _, Xf, _ = ofb.kernel_ss_kalmanFastFB(A, Q, H, Pinf, K, 0.01, y, 0, 1)
    _, Xs, _ = ofb.kernel_ss_kalmanFastFB(A, Q, H, Pinf, K, 0.01, y, 0, 0)
    mid = slice(500, 700)
    assert rel_err(Xf[0][:, mid], MF[:, mid]) < 1e-8
    assert rel_err(Xs[0][:, mid], MS[:, mid]) < 1e-8


def test_product_model_builder_matches_oracle(nsagp):
    from oracle import filterbank as ofb
    fb = importlib.import_module(nsagp.__name__ + ".filterbank")
    rng = np.random.default_rng(2)
    for kernel in ("exp", "matern32", "matern52"):
        D = 6
        lamx, varx, om = 1.0 / rng.uniform(40, 300, D), rng.uniform(0.2, 1.0, D), np.linspace(1.0, 0.1, D)
        a, b = ofb.get_disc_model(lamx, varx, om, D, kernel), fb.get_disc_model(lamx, varx, om, D, kernel)
        assert a[4:] == b[4:]
        for x, z in zip(a[:4], b[:4]):
            assert rel_err(z, x) < 1e-13
