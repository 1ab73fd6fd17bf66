"""CPU: pin the oracle's tilted-distribution moments (oracle/lik.py, restating likModulatorNMFPower.m:28-87 and
experiments/likModulatorPreCalcwn.m:28-86) against their DEFINITION, with no analytic marginalisation and no
derivative formulas: Z(mu) = E_{z,g ~ N(mu, diag s2)}[ N(y | f(z,g), sn2)^alpha ] integrated by a dense product
Gauss-Hermite rule over ALL of (z, g), dlZ and d2lZ as central differences of log Z in mu.  The reference ships no
values for this path, so identities of this kind are what the oracle can be held to."""
import math

import numpy as np
import pytest

from oracle import cubature as oc
from oracle import lik as ol

Q = 40                                  # nodes per dimension of the oracle-side rule over g (converged to ~1e-12)
QB = 36                                 # nodes per dimension of the brute-force rule over (z, g)


def _brute_lZ(kind, link, sn2, alpha, y, mu, s2, W):
    """log E[ p(y|z,g)^alpha ] (precalc: the power of the density itself; power: the reference drops the Power-EP
    constant, likModulatorNMFPower.m:49, i.e. integrates N(y | f, sn2/alpha))."""
    D, N = W.shape
    t, w = np.polynomial.hermite_e.hermegauss(QB)
    w = w / math.sqrt(2 * math.pi)
    grids = np.meshgrid(*([np.arange(QB)] * (D + N)), indexing="ij")
    x = np.stack([mu[i] + math.sqrt(s2[i]) * t[g.ravel()] for i, g in enumerate(grids)], axis=1)
    wt = np.prod(np.stack([w[g.ravel()] for g in grids], axis=1), axis=1)
    z, g = x[:, :D], x[:, D:]
    a = link(g) @ W.T
    if kind == "precalc":
        a = np.sqrt(a)
    f = np.sum(z * a, axis=1)
    if kind == "precalc":
        dens = (np.exp(-0.5 * (y - f) ** 2 / sn2) / math.sqrt(2 * math.pi * sn2)) ** alpha
    else:
        v = sn2 / alpha
        dens = np.exp(-0.5 * (y - f) ** 2 / v) / math.sqrt(2 * math.pi * v)
    return math.log(np.sum(wt * dens))


def _oracle(kind, link, hyp, alpha, y, mu, s2, W):
    N = W.shape[1]
    if kind == "power":
        return ol.likModulatorNMFPower(link, hyp, y, mu, s2, W, Q, alpha)
    xn, wn = oc.mvhermgauss(np.zeros(N), np.ones(N), Q)
    return ol.likModulatorPreCalcwn(link, hyp, y, mu, s2, W, alpha, wn, xn.T)


@pytest.mark.parametrize("kind", ["power", "precalc"])
@pytest.mark.parametrize("alpha", [1.0, 0.75])
@pytest.mark.parametrize("shift", [0.0, 1.0])
def test_moments_are_the_derivatives_of_the_tilted_normaliser(kind, alpha, shift):
    rng = np.random.default_rng(11)
    D, N = 2, 1
    W = rng.uniform(0.3, 1.0, (D, N))
    mu = np.concatenate([rng.normal(0, 0.5, D), rng.normal(0.3, 0.3, N)])
    s2 = np.concatenate([rng.uniform(0.2, 0.5, D), rng.uniform(0.1, 0.3, N)])
    sn2, y = 0.3, 0.4
    link = ol.softplus_link(shift)
    hyp = np.log([sn2])
    lZ, dlZ, d2lZ = _oracle(kind, link, hyp, alpha, y, mu, s2, W)
    b = lambda m: _brute_lZ(kind, link, sn2, alpha, y, m, s2, W)
    assert abs(lZ - b(mu)) < 1e-9 * max(1.0, abs(lZ))
    h = 1e-3
    for i in range(D + N):
        e = np.zeros(D + N); e[i] = h
        lp, lm, l0 = b(mu + e), b(mu - e), b(mu)
        # five-point stencils would be tighter; central differences at h = 1e-3 carry O(h^2) ~ 1e-6 truncation
        assert abs(dlZ[i] - (lp - lm) / (2 * h)) < 5e-6 * max(1.0, abs(dlZ[i]))
        assert abs(d2lZ[i] - (lp - 2 * l0 + lm) / h ** 2) < 5e-5 * max(1.0, abs(d2lZ[i]))


def test_two_modulators_share_a_subband():
    """N = 2 modulators mixed by W into D = 1 subband: the NMF product link(g) W' (likModulatorNMFPower.m:44)."""
    rng = np.random.default_rng(5)
    W = rng.uniform(0.3, 1.0, (1, 2))
    mu = np.array([0.4, -0.2, 0.5]); s2 = np.array([0.3, 0.2, 0.15])
    link = ol.softplus_link(1.0)
    for kind in ("power", "precalc"):
        lZ, dlZ, _ = _oracle(kind, link, np.log([0.2]), 0.75, 0.3, mu, s2, W)
        b = lambda m: _brute_lZ(kind, link, 0.2, 0.75, 0.3, m, s2, W)
        assert abs(lZ - b(mu)) < 1e-9
        for i in range(3):
            e = np.zeros(3); e[i] = 1e-3
            assert abs(dlZ[i] - (b(mu + e) - b(mu - e)) / 2e-3) < 5e-6


def test_ep_with_the_real_likelihood_is_exact_gp_regression_when_the_modulator_is_frozen():
    """One subband, one modulator whose prior variance is ~0: y_k = a z_k + noise with a = W softplus(0), so the whole
    chain  cubature -> likModulatorNMFPower -> ADF filter -> RTS smoother  (gf_ep_modulator_nmf.m:126-267) must return
    the dense GP-regression posterior of z and the Gaussian evidence of y."""
    from oracle import gf_ep, ssmodel as oss
    T, sn2 = 30, 0.05
    ws = np.array([0.8, 12.0, 0.6])
    wm = np.array([1e-12, 20.0])
    F, L, Qc, H, Pinf = oss.ss_modulators_nmf(ws, wm, "matern32", "matern52")
    A, Q = oss.lti_disc(F, L, Qc, 1.0)
    n = A.shape[0]
    W = np.array([[0.7]])
    a = 0.7 * math.log(2.0)
    rng = np.random.default_rng(3)
    y = rng.normal(0, 0.5, T)
    mom = ol.make_mom("power", ol.softplus_link(0.0), p=9)
    Eft, Varft, lb, ub, out = gf_ep.gf_ep_core(A, Q, H, Pinf, np.log([sn2]), W, y, mom, 1.0, [1.0], 1, True, np.arange(T))
    Apow = [np.eye(n)]
    for _ in range(T):
        Apow.append(A @ Apow[-1])
    h = H[0]
    K = np.array([[h @ (Apow[s - t] @ Pinf if s >= t else Pinf @ Apow[t - s].T) @ h for t in range(T)] for s in range(T)])
    S = a * a * K + sn2 * np.eye(T)
    mean = a * K @ np.linalg.solve(S, y)
    var = np.diag(K - a * a * K @ np.linalg.solve(S, K))
    assert np.allclose(Eft[0], mean, rtol=1e-7, atol=1e-9)
    assert np.allclose(Varft[0], var, rtol=1e-7, atol=1e-9)
    _, logdet = np.linalg.slogdet(S)
    ref = -0.5 * y @ np.linalg.solve(S, y) - 0.5 * logdet - 0.5 * T * math.log(2 * math.pi)
    assert abs(-out["nlZ"][0] - ref) < 1e-7 * abs(ref)


def test_oracle_entry_points_are_exact_on_the_frozen_modulator_models(nsagp):
    """The same closed-form cases the CUDA path is held to in tests/test_gpu_exact.py, for the oracle: Power EP
    (alpha = 0.75, sqrt model) posterior, the EKF entry point over several subbands (posterior + energy), the nlZ mode."""
    from test_gpu_exact import exact_case, exact_case_subbands
    from oracle import gf_ep, giekf, ssmodel as oss
    ss = lambda x, p1, p2, k1, k2: oss.ss_modulators_nmf(p1, p2, k1, k2)
    a = math.sqrt(1.1 * math.log1p(math.exp(-1.0)))
    hyp, y, mean, var, _ = exact_case(nsagp, a=a)
    t = np.arange(1.0, y.size + 1.0)
    wn, xn = oc.utp_ws(9, 2)
    mom = ol.make_mom("precalc", ol.softplus_link(1.0), wn=wn, xn_unscaled=xn)
    E, V = gf_ep.gf_ep_modulator_nmf(hyp.pack_log(), t, y, ss, mom, t, "matern32", "matern52", 1, 1, 2, 0.75, np.ones(3), 3)[:2]
    assert np.allclose(E[0], mean, rtol=1e-8, atol=1e-10) and np.allclose(V[0], var, rtol=1e-8, atol=1e-10)
    hyp, y, mean, var, lml = exact_case_subbands(nsagp, D=4, k1="exp")
    t = np.arange(1.0, y.size + 1.0)
    E, V = giekf.gf_giekf_modulator_nmf(hyp.pack_log(), t, y, ss, None, t, "exp", "matern52", 1, 4, 2, 1, 1)[:2]
    assert np.allclose(E[:4], mean, rtol=1e-8, atol=1e-10) and np.allclose(V[:4], var, rtol=1e-8, atol=1e-10)
    e = giekf.gf_giekf_modulator_nmf(hyp.pack_log(), t, y, ss, None, None, "exp", "matern52", 1, 4, 2, 1, 1)[0]
    assert abs(e + lml) < 1e-8 * abs(lml)
    hyp, y, _, _, lml = exact_case(nsagp)
    t = np.arange(1.0, y.size + 1.0)
    mom = ol.make_mom("power", ol.softplus_link(0.0), p=9)
    nlZ = gf_ep.gf_ep_modulator_nmf(hyp.pack_log(), t, y, ss, mom, None, "matern32", "matern52", 1, 1, 2, 1.0, np.ones(1), 1)[0]
    assert abs(float(np.ravel(nlZ)[0]) + lml) < 1e-8 * abs(lml)


def test_oracle_non_nmf_entry_point_is_exact_with_a_frozen_modulator():
    """gf_ep_modulator.m (no NMF weights, balanced model, likModulatorPower): one carrier x modulator pair whose
    modulator has prior variance ~0 is y = log(2) z + noise; the balancing transformation must not change the answer."""
    from oracle import gf_ep, ssmodel as oss
    T, sn2 = 30, 0.05
    par = np.array([0.8, 12.0, 0.6, 1e-12, 20.0])
    w = np.log(np.concatenate([[sn2], par]))
    F, L, Qc, H, Pinf = oss.ss_modulators(par, "matern32", "matern52")[:5]
    A, Q = oss.lti_disc(F, L, Qc, 1.0)
    n = A.shape[0]
    a = math.log(2.0)
    y = np.random.default_rng(8).normal(0, 0.5, T)
    t = np.arange(1.0, T + 1.0)
    link = ol.softplus_link(0.0)
    mom = lambda hyp, mu, s2, ep_frac, yall, k: ol.likModulatorPower(link, hyp, yall[k], mu, s2, 9, ep_frac)
    ss = lambda x, pr, k1, k2: oss.ss_modulators(pr, k1, k2)
    E, V, _, _, _, out = gf_ep.gf_ep_modulator(w, t, y, ss, mom, t, "matern32", "matern52", 1, 1.0, np.ones(2), 2)
    Apow = [np.eye(n)]
    for _ in range(T):
        Apow.append(A @ Apow[-1])
    h = H[0]
    K = np.array([[h @ (Apow[s - u] @ Pinf if s >= u else Pinf @ Apow[u - s].T) @ h for u in range(T)] for s in range(T)])
    S = a * a * K + sn2 * np.eye(T)
    assert np.allclose(E[0], a * K @ np.linalg.solve(S, y), rtol=1e-7, atol=1e-9)
    assert np.allclose(V[0], np.diag(K - a * a * K @ np.linalg.solve(S, K)), rtol=1e-7, atol=1e-9)
    _, logdet = np.linalg.slogdet(S)
    ref = -0.5 * y @ np.linalg.solve(S, y) - 0.5 * logdet - 0.5 * T * math.log(2 * math.pi)
    assert abs(-out["nlZ"][0] - ref) < 1e-7 * abs(ref)
