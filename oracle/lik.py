"""Oracle: tilted-distribution moments (test infrastructure only).

Restates matlab/likModulatorNMFPower.m:28-87 and
matlab/experiments/likModulatorPreCalcwn.m:28-86 ('infEP' branch, 3 outputs).
Shapes follow MATLAB: xn is (S, N), link_xn_W is (S, D), sums run over rows.
"""
import math

import numpy as np

from . import cubature


def softplus_link(shift=0.0):
    """demo_toy_modulators_nmf.m:11 ``link = @(g) log(1+exp(g))``; the experiments
    use ``log(1+exp(g-1))`` (experiments/train_model.m:38) -> shift = 1."""
    def link(g):
        return np.log(1 + np.exp(g - shift))
    return link


def _normpdf(y, mu, sigma):
    # MATLAB normpdf: exp(-0.5*((x-mu)./sigma).^2) ./ (sqrt(2*pi).*sigma)
    return np.exp(-0.5 * ((y - mu) / sigma) ** 2) / (math.sqrt(2 * math.pi) * sigma)


def _moments(link_xn_W, xn, wn, y, mu_z, mu_g, s2_z, s2_g, sn2, ep_fraction, pEP_const, jitter=1e-10):
    D = mu_z.size
    N = mu_g.size
    sn2_link_xn2_s2_z = sn2 / ep_fraction + (link_xn_W ** 2) @ s2_z      # :45
    link_xn_mu_z = link_xn_W @ mu_z                                       # :46
    xn_mu_g_s2_g = (xn - mu_g[None, :]) / s2_g[None, :]                   # :47
    normy_xn = _normpdf(y, link_xn_mu_z, np.sqrt(sn2_link_xn2_s2_z))      # :51-53
    Z = pEP_const * max(np.sum(wn * normy_xn), jitter)                    # :55
    if np.isnan(np.sum(wn * normy_xn)):
        # MATLAB max(NaN, jitter) = jitter
        Z = pEP_const * jitter
    Zinv = 1 / Z
    lZ = math.log(Z)
    dlZ = np.empty(D + N)
    d2lZ = np.empty(D + N)
    dZ_integrand1 = link_xn_W * ((y - link_xn_mu_z) / sn2_link_xn2_s2_z * normy_xn)[:, None]   # :59-60
    dZ1 = np.sum(wn[:, None] * dZ_integrand1, axis=0)
    dlZ[:D] = Zinv * pEP_const * dZ1                                      # :63
    dZ_integrand2 = xn_mu_g_s2_g * normy_xn[:, None]                      # :65
    dZ2 = np.sum(wn[:, None] * dZ_integrand2, axis=0)
    dlZ[D:] = Zinv * pEP_const * dZ2                                      # :68
    d2Z_integrand1 = link_xn_W ** 2 * ((((y - link_xn_mu_z) / sn2_link_xn2_s2_z) ** 2
                                        - (1 / sn2_link_xn2_s2_z)) * normy_xn)[:, None]       # :71-74
    d2lZ[:D] = -dlZ[:D] ** 2 + Zinv * pEP_const * np.sum(wn[:, None] * d2Z_integrand1, axis=0)
    d2Z_integrand2 = (xn_mu_g_s2_g ** 2 - (1 / s2_g[None, :])) * normy_xn[:, None]            # :77-79
    d2lZ[D:] = -dlZ[D:] ** 2 + Zinv * pEP_const * np.sum(wn[:, None] * d2Z_integrand2, axis=0)
    return lZ, dlZ, d2lZ


def _points(mu_g, s2_g, xn_unscaled):
    # xn = (mu_g + diag(s2_g.^0.5)*xn_unscaled)'      likModulatorNMFPower.m:34
    # NB: a negative cavity variance makes MATLAB go complex here;
    # the oracle lets NumPy produce NaN instead and callers flag the step.
    with np.errstate(invalid="ignore"):
        sd = s2_g ** 0.5
    return (mu_g[:, None] + sd[:, None] * xn_unscaled).T


def likModulatorNMFPower(link, hyp, y, mu, s2, W, p, ep_fraction):
    """likModulatorNMFPower.m:28-87.  Returns (lZ, dlZ[M], d2lZ[M])."""
    sn2 = math.exp(float(np.asarray(hyp).ravel()[0]))
    W = np.asarray(W, float)
    D, N = W.shape
    mu = np.asarray(mu, float).ravel(); s2 = np.asarray(s2, float).ravel()
    mu_z, mu_g = mu[:D], mu[D:]
    s2_z, s2_g = s2[:D], s2[D:]
    if p in (3, 5, 7, 9):
        wn, xn_unscaled = cubature.utp_ws(p, N)        # recomputed per call in the reference (:33)
        xn = _points(mu_g, s2_g, xn_unscaled)
    else:
        xn, wn = cubature.mvhermgauss(mu_g, s2_g, p)   # :41
    with np.errstate(all="ignore"):
        link_xn_W = link(xn) @ W.T                                            # :44
        return _moments(link_xn_W, xn, wn, y, mu_z, mu_g, s2_z, s2_g, sn2, ep_fraction, 1.0)   # pEP_const = 1 (:49)


def likModulatorPower(link, hyp, y, mu, s2, p, ep_fraction):
    """likModulatorPower.m:28-90: the likelihood of the model without NMF weights, y = sum_d z_d link(g_d) -- the
    arithmetic of likModulatorNMFPower with W = I (D pairs, mu = [mu_z; mu_g]) and a floor of 1e-8 (:29) under Z."""
    sn2 = math.exp(float(np.asarray(hyp).ravel()[0]))
    mu = np.asarray(mu, float).ravel(); s2 = np.asarray(s2, float).ravel()
    D = mu.size // 2
    mu_z, mu_g = mu[:D], mu[D:]
    s2_z, s2_g = s2[:D], s2[D:]
    if p in (3, 5, 7, 9):
        wn, xn_unscaled = cubature.utp_ws(p, D)                                # :34
        xn = _points(mu_g, s2_g, xn_unscaled)
    else:
        xn, wn = cubature.mvhermgauss(mu_g, s2_g, p)
    with np.errstate(all="ignore"):
        return _moments(link(xn), xn, wn, y, mu_z, mu_g, s2_z, s2_g, sn2, ep_fraction, 1.0, jitter=1e-8)


def likModulatorPreCalcwn(link, hyp, y, mu, s2, W, ep_fraction, wn, xn_unscaled):
    """experiments/likModulatorPreCalcwn.m:28-86 ("sqrt" model, Power-EP constant)."""
    sn2 = math.exp(float(np.asarray(hyp).ravel()[0]))
    W = np.asarray(W, float)
    D, N = W.shape
    mu = np.asarray(mu, float).ravel(); s2 = np.asarray(s2, float).ravel()
    mu_z, mu_g = mu[:D], mu[D:]
    s2_z, s2_g = s2[:D], s2[D:]
    xn = _points(mu_g, s2_g, np.asarray(xn_unscaled, float))                 # :34
    wn = np.asarray(wn, float).ravel()
    with np.errstate(all="ignore"):
        link_xn_W = np.sqrt(link(xn) @ W.T)                                   # :44
        pEP_const = (2 * math.pi * sn2) ** (0.5 * (1 - ep_fraction)) * ep_fraction ** (-0.5)   # :48
        return _moments(link_xn_W, xn, wn, y, mu_z, mu_g, s2_z, s2_g, sn2, ep_fraction, pEP_const)


def make_mom(kind, link, p=None, wn=None, xn_unscaled=None):
    """Build the ``mom(hyp,mu,s2,nmfW,ep_frac,yall,k)`` closure the demos build
    (demo_toy_modulators_nmf.m:81; experiments/train_model.m:213 for PreCalcwn).
    ``k`` is 0-based here."""
    if kind == "power":
        def mom(hyp, mu, s2, nmfW, ep_frac, yall, k):
            return likModulatorNMFPower(link, hyp, yall[k], mu, s2, nmfW, p, ep_frac)
    elif kind == "precalc":
        def mom(hyp, mu, s2, nmfW, ep_frac, yall, k):
            return likModulatorPreCalcwn(link, hyp, yall[k], mu, s2, nmfW, ep_frac, wn, xn_unscaled)
    else:
        raise ValueError(kind)
    return mom
