"""Oracle: sigma-point rules (test infrastructure only -- see oracle/__init__.py).

Restates matlab/symmetric-cubature-rules/{utp,ut3,ut5,ut7,ut9}_ws.m, sym_set.m,
matlab/mvhermgauss.m and matlab/gauher.m.  Points are returned as in MATLAB:
``W`` is (S,), ``SX`` is (n, S) (one column per sigma point).
"""
import math

import numpy as np


def _nupk(n, k):
    # ut9_ws.m:102-104  prod((n-k+1):n); empty range -> 1
    out = 1.0
    for v in range(n - k + 1, n + 1):
        out *= v
    return out


def _ndownk(n, k):
    # ut9_ws.m:106-108
    return _nupk(n, k) / math.factorial(k)


def sym_set(n, gen):
    """sym_set.m:1-40.  ``nargin < 3`` is always true there, so nonzero = 0."""
    gen = list(gen)
    if len(gen) == 0:
        return np.zeros((n, 1))
    cols = []
    for i in range(n):  # MATLAB i = 1..n  -> python i = 0..n-1
        u = np.zeros(n)
        u[i] = gen[0]
        if len(gen) > 1:
            if abs(gen[0] - gen[1]) < np.finfo(float).eps:
                V = sym_set(n - (i + 1), gen[1:])
                for j in range(V.shape[1]):
                    u[i + 1:] = V[:, j]
                    cols.append(u.copy())
                    cols.append(-u)
            else:
                V = sym_set(n - 1, gen[1:])
                idx = [q for q in range(n) if q != i]
                for j in range(V.shape[1]):
                    u[idx] = V[:, j]
                    cols.append(u.copy())
                    cols.append(-u)
        else:
            cols.append(u.copy())
            cols.append(-u)
    if not cols:
        return np.zeros((n, 0))
    return np.stack(cols, axis=1)


def ut3_ws(n):
    """ut3_ws.m:1-27 (kappa forced to 0 at line 9)."""
    kappa = 0.0
    W = np.empty(2 * n + 1)
    W[0] = kappa / (n + kappa)
    W[1:] = 1.0 / (2.0 * (n + kappa))
    SX = np.concatenate([np.zeros((n, 1)), np.eye(n), -np.eye(n)], axis=1)
    SX = math.sqrt(n + kappa) * SX
    return W, SX


def ut5_ws(n):
    """ut5_ws.m:1-40."""
    I0, I2, I4, I22 = 1.0, 1.0, 3.0, 1.0
    u = math.sqrt(I4 / I2)
    A0 = I0 - n * (I2 / I4) ** 2 * (I4 - 0.5 * (n - 1) * I22)
    A1 = 0.5 * (I2 / I4) ** 2 * (I4 - (n - 1) * I22)
    A11 = 0.25 * (I2 / I4) ** 2 * I22
    U0 = sym_set(n, [])
    U1 = sym_set(n, [u])
    U2 = sym_set(n, [u, u])
    SX = np.concatenate([U0, U1, U2], axis=1)
    W = np.concatenate([A0 * np.ones(U0.shape[1]), A1 * np.ones(U1.shape[1]),
                        A11 * np.ones(U2.shape[1])])
    return W, SX


def _pos_roots(coeffs):
    tmp = np.roots(coeffs)
    tmp = tmp[tmp > 0]
    return float(np.real(tmp[0])), float(np.real(tmp[1]))


def ut7_ws(n):
    """ut7_ws.m:1-52."""
    I222, I22, I24, I2, I6, I4, I0 = 1.0, 1.0, 3.0, 1.0, 15.0, 3.0, 1.0
    u, v = _pos_roots([I2 ** 2 - I0 * I4, 0, -(I2 * I4 - I0 * I6), 0, I4 ** 2 - I2 * I6])
    u2 = u * u; u4 = u2 * u2; u6 = u4 * u2
    v2 = v * v; v4 = v2 * v2; v6 = v4 * v2
    A111 = I222 / 8 / u6
    tmp = 0.25 * np.linalg.solve(np.array([[u4, v4], [u6, v6]]),
                                 np.array([I22, I24]) - 8 * (n - 2) * np.array([u4, u6]) * A111)
    A11, A22 = tmp
    tmp = -2 * (n - 1) * np.array([A11, A22]) + 0.5 * np.linalg.solve(
        np.array([[u2, v2], [u4, v4]]),
        np.array([I2, I4]) - 8 * (n - 1) * (n - 2) / 2 * np.array([u2, u4]) * A111)
    A1, A2 = tmp
    A0 = I0 - 2 * n * (A1 + A2) - 4 * n * (n - 1) / 2 * (A11 + A22) - 8 * n * (n - 1) * (n - 2) / 6 * A111
    sets = [sym_set(n, []), sym_set(n, [u]), sym_set(n, [v]), sym_set(n, [u, u]),
            sym_set(n, [v, v]), sym_set(n, [u, u, u])]
    ws = [A0, A1, A2, A11, A22, A111]
    SX = np.concatenate(sets, axis=1)
    W = np.concatenate([a * np.ones(s.shape[1]) for a, s in zip(ws, sets)])
    return W, SX


def ut9_ws(n):
    """ut9_ws.m:26-100, including the doubled minus sign of lines 78-79
    (``... - ...`` continued by ``-8*ndownk(n,3)*(A111+A222)``), which makes the
    weights sum to something other than 1 for n >= 3."""
    I2222 = 1.0; I224 = 3.0; I222 = 1.0; I44 = 9.0; I26 = 15.0; I24 = 3.0; I22 = 1.0
    I8 = 105.0; I6 = 15.0; I4 = 3.0; I2 = 1.0; I0 = 1.0
    u, v = _pos_roots([I4 ** 2 - I2 * I6, 0, -(I4 * I6 - I2 * I8), 0, I6 ** 2 - I4 * I8])
    u2 = u * u; u4 = u2 * u2; u6 = u4 * u2; u8 = u4 * u4
    v2 = v * v; v4 = v2 * v2; v6 = v4 * v2; v8 = v4 * v4
    A1111 = I2222 / 16 / u8
    M68 = np.array([[u6, v6], [u8, v8]])
    tmp = 1 / 8 * np.linalg.solve(M68, np.array([I222, I224]) - 16 * (n - 3) * A1111 * np.array([u6, u8]))
    A111, A222 = tmp
    A12 = (I26 - I44) / (4 * u2 * v2 * (u2 - v2) ** 2)
    tmp = -2 * (n - 2) * np.array([A111, A222]) + 1 / 4 * np.linalg.solve(
        M68,
        np.array([I24, I26]) - 4 * np.array([u4 * v2 + u2 * v4, u6 * v2 + u2 * v6]) * A12
        - 16 * _ndownk(n - 2, 2) * np.array([u6, u8]) * A1111)
    A11, A22 = tmp
    tmp = (-2 * (n - 1) * np.array([A11 + A12, A22 + A12])
           - 4 * _ndownk(n - 1, 2) * np.array([A111, A222])
           + 0.5 * np.linalg.solve(np.array([[u2, v2], [u4, v4]]),
                                   np.array([I2, I4]) - 16 * _ndownk(n - 1, 3) * np.array([u2, u4]) * A1111))
    A1, A2 = tmp
    # NB: "- -8*..." is the reference's literal arithmetic (ut9_ws.m:78-79)
    A0 = (I0 - 2 * n * (A1 + A2) - 4 * _ndownk(n, 2) * (A11 + 2 * A12 + A22)
          - -8 * _ndownk(n, 3) * (A111 + A222) - 16 * _ndownk(n, 4) * A1111)
    sets = [sym_set(n, []), sym_set(n, [u]), sym_set(n, [v]), sym_set(n, [u, u]),
            sym_set(n, [u, v]), sym_set(n, [v, v]), sym_set(n, [u, u, u]),
            sym_set(n, [v, v, v]), sym_set(n, [u, u, u, u])]
    ws = [A0, A1, A2, A11, A12, A22, A111, A222, A1111]
    SX = np.concatenate(sets, axis=1)
    W = np.concatenate([a * np.ones(s.shape[1]) for a, s in zip(ws, sets)])
    return W, SX


def utp_ws(p, n):
    """utp_ws.m:3-14."""
    if p == 3:
        return ut3_ws(n)
    if p == 5:
        return ut5_ws(n)
    if p == 7:
        return ut7_ws(n)
    if p == 9:
        return ut9_ws(n)
    raise ValueError("Not implemented")


_GAUHER20_X = np.array([
    -7.619048541679757, -6.510590157013656, -5.578738805893203,
    -4.734581334046057, -3.943967350657318, -3.18901481655339,
    -2.458663611172367, -1.745247320814127, -1.042945348802751,
    -0.346964157081356, 0.346964157081356, 1.042945348802751,
    1.745247320814127, 2.458663611172367, 3.18901481655339,
    3.943967350657316, 4.734581334046057, 5.578738805893202,
    6.510590157013653, 7.619048541679757])
_GAUHER20_W = np.array([
    0.000000000000126, 0.000000000248206, 0.000000061274903,
    0.00000440212109, 0.000128826279962, 0.00183010313108,
    0.013997837447101, 0.061506372063977, 0.161739333984,
    0.260793063449555, 0.260793063449555, 0.161739333984,
    0.061506372063977, 0.013997837447101, 0.00183010313108,
    0.000128826279962, 0.00000440212109, 0.000000061274903,
    0.000000000248206, 0.000000000000126])


def gauher(N):
    """gauher.m:34-54: 20-point table verbatim, else Golub-Welsch on the Jacobi matrix."""
    if N == 20:
        return _GAUHER20_X.copy(), _GAUHER20_W.copy()
    b = np.sqrt(np.arange(1, N) / 2.0)
    J = np.diag(b, 1) + np.diag(b, -1)
    D, V = np.linalg.eigh(J)  # MATLAB eig of a symmetric matrix: ascending eigenvalues
    w = V[0, :] ** 2
    x = math.sqrt(2) * D
    return x, w


def mvhermgauss_unit(D, N):
    """Unit-scale part of mvhermgauss.m:15-23: tensor grid via ndgrid (first
    dimension varies fastest).  Returns (x_loc (N^D, D), wn (N^D,))."""
    t, w = gauher(N)
    grids = np.meshgrid(*([np.arange(N)] * D), indexing="ij")
    idx = [g.reshape(-1, order="F") for g in grids]  # x(:) is column-major
    x_loc = np.stack([t[i] for i in idx], axis=1)
    w_loc = np.stack([w[i] for i in idx], axis=1)
    wn = np.prod(w_loc, axis=1)
    return x_loc, wn


def mvhermgauss(mu, s2, N):
    """mvhermgauss.m:1-24.  Returns xn (N^D, D), wn (N^D,)."""
    mu = np.asarray(mu, float).ravel()
    s2 = np.asarray(s2, float).ravel()
    x_loc, wn = mvhermgauss_unit(mu.size, N)
    sig = np.diag(np.sqrt(s2))
    xn = x_loc @ sig + np.tile(mu, (x_loc.shape[0], 1))
    return xn, wn
