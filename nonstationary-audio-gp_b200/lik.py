"""Likelihood / moment-matching closures.

In the reference ``mom`` is a function handle built by the caller
(matlab/demo_toy_modulators_nmf.m:78-81, matlab/experiments/train_model.m:186-190):

    mom = @(hyp,mu,s2,nmfW,ep_frac,yall,k) feval(likfunc,link,hyp,yall(k),mu,s2,nmfW,p,ep_frac,'infEP');

A handle cannot cross into CUDA, so here ``mom`` is a small descriptor object
the entry points recognise: which likelihood file, which link,
which sigma-point rule.  It is still callable with the handle's seven
arguments -- the call runs the same CUDA kernel for a single step.
"""
import math

import numpy as np

from . import _lib, cubature


class Softplus:
    """``link = @(g) log(1+exp(g-shift))`` (demo_toy_modulators_nmf.m:11 shift 0,
    experiments/train_model.m:38 shift 1)."""

    def __init__(self, shift=0.0):
        self.shift = float(shift)

    def __call__(self, g):
        return np.log(1 + np.exp(np.asarray(g, float) - self.shift))


class Moments:
    """Descriptor of ``mom``.  kind 0 = likModulatorNMFPower, 1 = likModulatorPreCalcwn."""

    def __init__(self, kind, link, wn, xn_unscaled, p=None):
        if not isinstance(link, Softplus):
            raise TypeError("the GPU path implements the softplus link log(1+exp(g-c)); pass lik.Softplus(c)")
        self.kind = int(kind)
        self.link = link
        self.p = p
        self.wn = _lib.as_f64(np.asarray(wn).ravel())
        xn = np.asarray(xn_unscaled, float)
        if xn.ndim != 2 or xn.shape[1] != self.wn.size:
            raise ValueError("xn_unscaled must be N-by-S with S = numel(wn)")
        self.N = xn.shape[0]
        self.S = xn.shape[1]
        self.xn = np.asfortranarray(xn)           # N-by-S column-major, as utp_ws returns it

    def c_lik(self, lik_param, W, keep):
        """Fill an nsagp_lik; ``keep`` collects the arrays that must stay alive."""
        W = np.asfortranarray(np.asarray(W, float))
        if W.shape[1] != self.N:
            raise ValueError("W has %d columns but the sigma points are %d-dimensional" % (W.shape[1], self.N))
        keep.extend([W, self.wn, self.xn])
        L = _lib.Lik()
        L.kind = self.kind
        L.sn2 = math.exp(float(np.asarray(lik_param, float).ravel()[0]))
        L.link_shift = self.link.shift
        L.W = W.ctypes.data_as(_lib.c_double_p)
        L.S = self.S
        L.wn = _lib.dptr(self.wn)
        L.xn = self.xn.ctypes.data_as(_lib.c_double_p)
        return L

    def batch(self, hyp, y, mu, s2, nmfW, ep_frac, warp_form=False):
        """Moments for T independent steps: mu, s2 are M-by-T.  Returns (lZ[T], dlZ[M,T], d2lZ[M,T])."""
        W = np.asarray(nmfW, float)
        D, N = W.shape
        y = _lib.as_f64(np.atleast_1d(y).ravel())
        T = y.size
        mu = _lib.as_f64(np.asarray(mu, float).reshape(D + N, T).T)     # time-major, M contiguous
        s2 = _lib.as_f64(np.asarray(s2, float).reshape(D + N, T).T)
        lZ = np.empty(T); d1 = np.empty((T, D + N)); d2 = np.empty((T, D + N))
        keep = []
        L = self.c_lik(hyp, W, keep)
        fn = _lib.lib().nsagp_mom_batch_warp if warp_form else _lib.lib().nsagp_mom_batch
        _lib.check(fn(L, D, N, float(ep_frac), T, _lib.dptr(y), _lib.dptr(mu), _lib.dptr(s2),
                      _lib.dptr(lZ), _lib.dptr(d1), _lib.dptr(d2)))
        return lZ, d1.T.copy(), d2.T.copy()

    def __call__(self, hyp, mu, s2, nmfW, ep_frac, yall, k):
        """The handle's signature (k is 0-based).  Returns (lZ, dlZ[M], d2lZ[M])."""
        lZ, d1, d2 = self.batch(hyp, np.asarray(yall, float).ravel()[k], np.asarray(mu, float).reshape(-1, 1),
                                np.asarray(s2, float).reshape(-1, 1), nmfW, ep_frac)
        return float(lZ[0]), d1[:, 0], d2[:, 0]


def _rule(p, N):
    if p in (3, 5, 7, 9):
        return cubature.utp_ws(p, N)                # likModulatorNMFPower.m:32-35
    wn, xn = cubature.mvhermgauss_unit(N, p)        # :41 tensor Gauss-Hermite with p points per dimension
    return wn, xn


def likModulatorNMFPower(link, p_cubature, N):
    """``mom`` for matlab/likModulatorNMFPower.m with cubature order ``p_cubature``
    (3/5/7/9: symmetric rule; anything else: p-point Gauss-Hermite per dimension)."""
    wn, xn = _rule(int(p_cubature), int(N))
    return Moments(0, link, wn, xn, p=int(p_cubature))


def likModulatorPower(link, p_cubature, D):
    """``mom`` for matlab/likModulatorPower.m, the likelihood of the model without NMF weights (``gf_ep_modulator``,
    D carrier x modulator pairs): the arithmetic of likModulatorNMFPower with W = I.  One deviation: the floor under
    the tilted normaliser Z is the NMF files' 1e-10 (csrc/common.cuh kJitter), not this file's 1e-8 (:29) -- it
    differs only at steps whose Z is below 1e-8."""
    return likModulatorNMFPower(link, p_cubature, D)


def likModulatorPreCalcwn(link, wn, xn_unscaled):
    """``mom`` for matlab/experiments/likModulatorPreCalcwn.m (spectrogram model,
    Power-EP constant, caller-supplied sigma points)."""
    return Moments(1, link, wn, xn_unscaled)
