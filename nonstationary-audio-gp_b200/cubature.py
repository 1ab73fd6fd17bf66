"""Sigma-point tables for the modulator integral, computed once on the host and
uploaded to the GPU (the reference recomputes them inside every ``mom`` call,
matlab/likModulatorNMFPower.m:33).

Interface mirrors matlab/symmetric-cubature-rules/utp_ws.m:3-14
(``[W,SX] = utp_ws(p,n)``) and matlab/mvhermgauss.m / gauher.m.  The rules are
the McNamee-Stenger fully symmetric formulas; generators are the closed-form
roots of He4 (p=7) and He5 (p=9).  The 9th-order centre weight reproduces the
reference's doubled minus sign (ut9_ws.m:78-79) so that results match it.
"""
import functools
import itertools
import math

import numpy as np


def _comb(n, k):
    return math.comb(n, k) if n >= k >= 0 else 0


def _orbit(n, gens):
    """All points with the generator magnitudes ``gens`` placed on distinct axes,
    all sign patterns; enumerated in the reference's order (sym_set.m): leading
    axis ascending, (+,-) pairs innermost."""
    gens = list(gens)
    if not gens:
        return np.zeros((n, 1))
    out = []
    g0, rest = gens[0], gens[1:]
    same = bool(rest) and abs(g0 - rest[0]) < np.finfo(float).eps
    for i in range(n):
        if not rest:
            tails = [np.zeros(n - 1)]
            free = [q for q in range(n) if q != i]
        elif same:
            sub = _orbit(n - i - 1, rest)
            tails = [sub[:, j] for j in range(sub.shape[1])]
            free = list(range(i + 1, n))
        else:
            sub = _orbit(n - 1, rest)
            tails = [sub[:, j] for j in range(sub.shape[1])]
            free = [q for q in range(n) if q != i]
        for t in tails:
            p = np.zeros(n)
            p[i] = g0
            p[free] = t[:len(free)]
            out.append(p)
            out.append(-p)
    if not out:
        return np.zeros((n, 0))
    return np.stack(out, axis=1)


def _assemble(n, parts):
    pts = [_orbit(n, g) for _, g in parts]
    SX = np.concatenate(pts, axis=1)
    W = np.concatenate([np.full(p.shape[1], a) for (a, _), p in zip(parts, pts)])
    return W, SX


def _rule3(n):
    # kappa = 0 (ut3_ws.m:9): centre weight 0, points at sqrt(n)
    # ut3_ws.m:23 orders the points [0, +axes, -axes]
    u = math.sqrt(n)
    SX = u * np.concatenate([np.zeros((n, 1)), np.eye(n), -np.eye(n)], axis=1)
    W = np.concatenate([[0.0], np.full(2 * n, 1.0 / (2 * n))])
    return W, SX


def _rule5(n):
    u = math.sqrt(3.0)
    A0 = 1 - n / 9.0 * (3 - 0.5 * (n - 1))
    A1 = (3 - (n - 1)) / 18.0
    A11 = 1.0 / 36.0
    return _assemble(n, [(A0, []), (A1, [u]), (A11, [u, u])])


def _solve2(a, b, c, d, r0, r1):
    det = a * d - b * c
    return (d * r0 - b * r1) / det, (a * r1 - c * r0) / det


def _rule7(n):
    u = math.sqrt(3 + math.sqrt(6)); v = math.sqrt(3 - math.sqrt(6))
    u2, v2 = u * u, v * v
    u4, v4 = u2 * u2, v2 * v2
    u6, v6 = u4 * u2, v4 * v2
    A111 = 1.0 / 8 / u6
    x, y = _solve2(u4, v4, u6, v6, 1 - 8 * (n - 2) * u4 * A111, 3 - 8 * (n - 2) * u6 * A111)
    A11, A22 = 0.25 * x, 0.25 * y
    x, y = _solve2(u2, v2, u4, v4, 1 - 4 * (n - 1) * (n - 2) * u2 * A111, 3 - 4 * (n - 1) * (n - 2) * u4 * A111)
    A1 = -2 * (n - 1) * A11 + 0.5 * x
    A2 = -2 * (n - 1) * A22 + 0.5 * y
    A0 = 1 - 2 * n * (A1 + A2) - 4 * _comb(n, 2) * (A11 + A22) - 8 * _comb(n, 3) * A111
    return _assemble(n, [(A0, []), (A1, [u]), (A2, [v]), (A11, [u, u]), (A22, [v, v]), (A111, [u, u, u])])


def _rule9(n):
    u = math.sqrt(5 + math.sqrt(10)); v = math.sqrt(5 - math.sqrt(10))
    u2, v2 = u * u, v * v
    u4, v4 = u2 * u2, v2 * v2
    u6, v6 = u4 * u2, v4 * v2
    u8, v8 = u4 * u4, v4 * v4
    A1111 = 1.0 / 16 / u8
    x, y = _solve2(u6, v6, u8, v8, 1 - 16 * (n - 3) * A1111 * u6, 3 - 16 * (n - 3) * A1111 * u8)
    A111, A222 = x / 8, y / 8
    A12 = (15.0 - 9.0) / (4 * u2 * v2 * (u2 - v2) ** 2)
    c2 = _comb(n - 2, 2)
    x, y = _solve2(u6, v6, u8, v8,
                   3 - 4 * (u4 * v2 + u2 * v4) * A12 - 16 * c2 * u6 * A1111,
                   15 - 4 * (u6 * v2 + u2 * v6) * A12 - 16 * c2 * u8 * A1111)
    A11 = -2 * (n - 2) * A111 + x / 4
    A22 = -2 * (n - 2) * A222 + y / 4
    c3 = _comb(n - 1, 3)
    x, y = _solve2(u2, v2, u4, v4, 1 - 16 * c3 * u2 * A1111, 3 - 16 * c3 * u4 * A1111)
    A1 = -2 * (n - 1) * (A11 + A12) - 4 * _comb(n - 1, 2) * A111 + 0.5 * x
    A2 = -2 * (n - 1) * (A22 + A12) - 4 * _comb(n - 1, 2) * A222 + 0.5 * y
    # Reference quirk (ut9_ws.m:78-79): the 3-generator term enters with a PLUS
    # sign, so sum(W) != 1 for n >= 3.  Kept on purpose -- results must match.
    A0 = (1 - 2 * n * (A1 + A2) - 4 * _comb(n, 2) * (A11 + 2 * A12 + A22)
          + 8 * _comb(n, 3) * (A111 + A222) - 16 * _comb(n, 4) * A1111)
    return _assemble(n, [(A0, []), (A1, [u]), (A2, [v]), (A11, [u, u]), (A12, [u, v]), (A22, [v, v]),
                         (A111, [u, u, u]), (A222, [v, v, v]), (A1111, [u, u, u, u])])


@functools.lru_cache(maxsize=None)
def _utp_cached(p, n):
    rule = {3: _rule3, 5: _rule5, 7: _rule7, 9: _rule9}.get(p)
    if rule is None:
        raise ValueError("Not implemented")          # utp_ws.m:13
    if n < 2:
        raise ValueError("symmetric cubature needs at least 2 modulators")
    W, SX = rule(n)
    W.setflags(write=False); SX.setflags(write=False)
    return W, SX


def utp_ws(p, n):
    """``[W,SX] = utp_ws(p,n)``: weights (S,), unit sigma points (n,S)."""
    return _utp_cached(int(p), int(n))


_GH20_X = (7.619048541679757, 6.510590157013656, 5.578738805893203, 4.734581334046057,
           3.943967350657318, 3.18901481655339, 2.458663611172367, 1.745247320814127,
           1.042945348802751, 0.346964157081356)
_GH20_XR = (0.346964157081356, 1.042945348802751, 1.745247320814127, 2.458663611172367,
            3.18901481655339, 3.943967350657316, 4.734581334046057, 5.578738805893202,
            6.510590157013653, 7.619048541679757)
_GH20_W = (0.000000000000126, 0.000000000248206, 0.000000061274903, 0.00000440212109,
           0.000128826279962, 0.00183010313108, 0.013997837447101, 0.061506372063977,
           0.161739333984, 0.260793063449555)


@functools.lru_cache(maxsize=None)
def gauher(N):
    """``[x,w] = gauher(N)`` (gauher.m:34-54), probabilists' weight.  N = 20 returns
    the reference's 15-digit table (which is not exactly symmetric)."""
    N = int(N)
    if N == 20:
        x = np.array([-v for v in _GH20_X] + list(_GH20_XR))
        w = np.array(list(_GH20_W) + list(reversed(_GH20_W)))
    else:
        off = np.sqrt(np.arange(1, N) / 2.0)
        lam, vec = np.linalg.eigh(np.diag(off, 1) + np.diag(off, -1))
        x = math.sqrt(2.0) * lam
        w = vec[0] ** 2
    x.setflags(write=False); w.setflags(write=False)
    return x, w


@functools.lru_cache(maxsize=None)
def mvhermgauss_unit(N, p):
    """Unit tensor Gauss-Hermite grid used by mvhermgauss.m:15-23: returns
    (wn (p^N,), xn_unscaled (N, p^N)), first dimension varying fastest."""
    x, w = gauher(p)
    S = p ** N
    xn = np.empty((N, S)); wn = np.ones(S)
    for j, combo in enumerate(itertools.product(range(p), repeat=N)):
        combo = combo[::-1]                # ndgrid: dimension 1 is the fastest index
        for d in range(N):
            xn[d, j] = x[combo[d]]
            wn[j] *= w[combo[d]]
    xn.setflags(write=False); wn.setflags(write=False)
    return wn, xn
