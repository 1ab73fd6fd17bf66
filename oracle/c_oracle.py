"""ctypes binding of oracle/c/libnsagp_oracle.so -- the second, independent (plain C)
restatement of the infinite-horizon hot path.  TEST INFRASTRUCTURE ONLY (see
oracle/__init__.py): used by tests/ to cross-check the NumPy oracle and by bench.py's
cpu_baseline / --impl reference legs as the timed CPU port.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "c", "libnsagp_oracle.so")
dp = C.POINTER(C.c_double)


class Problem(C.Structure):
    _fields_ = [("D", C.c_int), ("N", C.c_int), ("n", C.c_int), ("ilist", C.POINTER(C.c_int)),
                ("A", dp), ("H", dp), ("Pinf", dp), ("nr", C.c_int), ("r", dp), ("PP", dp), ("PG", dp),
                ("kind", C.c_int), ("lik_param", C.c_double), ("shift", C.c_double), ("W", dp), ("S", C.c_int),
                ("wn", dp), ("xn", dp), ("alpha", C.c_double), ("damping", dp), ("ep_itts", C.c_int)]


_lib = None


def build():
    subprocess.run(["make", "-C", os.path.join(HERE, "c"), "-s"], check=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        L.oracle_mom.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, dp, C.c_int, dp, dp, C.c_double,
                                 C.c_double, dp, dp, dp, dp, dp]
        L.oracle_ihgp_predict.argtypes = [C.POINTER(Problem), dp, C.c_long, dp, dp, dp, dp, dp, dp, dp,
                                          C.POINTER(C.c_long), dp]
        L.oracle_ihgp_nlz.argtypes = [C.POINTER(Problem), dp, C.c_long, C.c_int, dp, dp, dp, dp, dp]
        _lib = L
    return _lib


def _f(a):
    return np.asfortranarray(np.asarray(a, float))


def _p(a):
    return a.ctypes.data_as(dp)


def mom(kind, lik_param, shift, W, wn, xn_unscaled, alpha, y, mu, s2):
    """[lZ, dlZ, d2lZ] of likModulatorNMFPower (kind 0) / likModulatorPreCalcwn (kind 1)."""
    W = _f(W); D, N = W.shape
    wn = _f(np.ravel(wn)); xn = _f(xn_unscaled)
    mu = _f(np.ravel(mu)); s2 = _f(np.ravel(s2))
    lZ = C.c_double(); d1 = np.empty(D + N); d2 = np.empty(D + N)
    rc = lib().oracle_mom(kind, float(lik_param), float(shift), D, N, _p(W), wn.size, _p(wn), _p(xn), float(alpha),
                          float(y), _p(mu), _p(s2), C.byref(lZ), _p(d1), _p(d2))
    assert rc == 0
    return lZ.value, d1, d2


class IhgpProblem:
    """Dense model + tables in the oracle's own terms (A, H, Pinf dense; PPlist/PGlist as
    oracle.ihgp_ep.ihgp_setup returns them)."""

    def __init__(self, A, H, Pinf, tabs, kind, lik_param, shift, W, wn, xn_unscaled, alpha, damping, ep_itts):
        self.keep = k = {}
        k["A"], k["H"], k["Pinf"], k["W"] = _f(A), _f(H), _f(Pinf), _f(W)
        k["ilist"] = np.ascontiguousarray(tabs["ilist"], dtype=np.int32)
        k["r"] = _f(tabs["r"])
        k["PP"] = np.ascontiguousarray(np.concatenate([np.asarray(t).ravel() for t in tabs["PPlist"]]))
        k["PG"] = None if tabs.get("PGlist") is None else \
            np.ascontiguousarray(np.concatenate([np.asarray(t).ravel() for t in tabs["PGlist"]]))
        k["wn"] = _f(np.ravel(wn)); k["xn"] = _f(xn_unscaled)
        k["damping"] = _f(np.atleast_1d(damping))
        D, N = k["W"].shape
        self.M, self.n = D + N, k["A"].shape[0]
        p = Problem()
        p.D, p.N, p.n = D, N, self.n
        p.ilist = k["ilist"].ctypes.data_as(C.POINTER(C.c_int))
        p.A, p.H, p.Pinf = _p(k["A"]), _p(k["H"]), _p(k["Pinf"])
        p.nr, p.r, p.PP = k["r"].size, _p(k["r"]), _p(k["PP"])
        p.PG = _p(k["PG"]) if k["PG"] is not None else None
        p.kind, p.lik_param, p.shift = int(kind), float(np.ravel(lik_param)[0]), float(shift)
        p.W, p.S, p.wn, p.xn = _p(k["W"]), k["wn"].size, _p(k["wn"]), _p(k["xn"])
        p.alpha, p.damping, p.ep_itts = float(alpha), _p(k["damping"]), int(ep_itts)
        self.p = p

    def predict(self, y):
        y = _f(np.ravel(y)); T = y.size; M, n, itts = self.M, self.n, self.p.ep_itts
        MS = np.empty((n, T), order="F"); tt = np.empty((M, T), order="F"); tn = np.empty((M, T), order="F")
        R = np.empty((M, T), order="F"); nlZ = np.empty(itts); Eft = np.empty((M, T), order="F")
        Varft = np.empty(M); neg = C.c_long(); md = np.empty(itts)
        rc = lib().oracle_ihgp_predict(C.byref(self.p), _p(y), T, _p(MS), _p(tt), _p(tn), _p(R), _p(nlZ), _p(Eft),
                                       _p(Varft), C.byref(neg), _p(md))
        assert rc == 0
        return dict(MS=MS, ttau=tt, tnu=tn, R=R, nlZ=nlZ, Eft=Eft, Varft=np.tile(Varft[:, None], (1, T)),
                    n_negcav=int(neg.value), maxDiffM=md)

    def nlz(self, y, running=False):
        y = _f(np.ravel(y)); T = y.size; M = self.M
        tt = np.empty((M, T), order="F"); tn = np.empty((M, T), order="F"); R = np.empty((M, T), order="F")
        lZk = np.empty(T); e = C.c_double()
        rc = lib().oracle_ihgp_nlz(C.byref(self.p), _p(y), T, int(running), C.byref(e), _p(tt), _p(tn), _p(R), _p(lZk))
        assert rc == 0
        return e.value, dict(ttau=tt, tnu=tn, R=R, lZ=lZk)
