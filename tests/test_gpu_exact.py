"""GPU parity against MATHEMATICS rather than against the oracle: with the modulators frozen (prior variance ~0) the
model of gf_ep_modulator_nmf.m is y_k = a z_k + noise, a = sum_n W_n softplus(0), and the whole chain
cubature -> likModulatorNMFPower -> ADF pass -> smoother (-> site updates -> frozen-site passes) must return the dense
GP-regression posterior of z and the Gaussian evidence of y.  The only thing borrowed from oracle/ is the state-space
model used to build the dense kernel matrix."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def exact_case(nsagp, T=40, sn2=0.05, seed=3, a=None):
    from oracle import ssmodel as oss
    hyp = nsagp.synth.Hypers(sn2, var_fast=np.array([0.8]), len_fast=np.array([12.0]), omega=np.array([0.6]),
                             var_slow=np.array([1e-12, 1e-12]), len_slow=np.array([20.0, 35.0]), W=np.array([[0.7, 0.4]]))
    F, L, Qc, H, Pinf = oss.ss_modulators_nmf(hyp.w_sub(), hyp.w_mod(), "matern32", "matern52")
    A, Q = oss.lti_disc(F, L, Qc, 1.0)
    n = A.shape[0]
    a = (0.7 + 0.4) * math.log(2.0) if a is None else a
    y = np.random.default_rng(seed).normal(0, 0.5, T)
    Apow = [np.eye(n)]
    for _ in range(T):
        Apow.append(A @ Apow[-1])
    h = H[0]
    K = np.array([[h @ (Apow[s - t] @ Pinf if s >= t else Pinf @ Apow[t - s].T) @ h for t in range(T)] for s in range(T)])
    S = a * a * K + sn2 * np.eye(T)
    mean = a * K @ np.linalg.solve(S, y)
    var = np.diag(K - a * a * K @ np.linalg.solve(S, K))
    _, logdet = np.linalg.slogdet(S)
    lml = -0.5 * y @ np.linalg.solve(S, y) - 0.5 * logdet - 0.5 * T * math.log(2 * math.pi)
    return hyp, y, mean, var, lml


@pytest.mark.parametrize("itts", [1, 3])
def test_gfep_equals_dense_gp_regression_with_frozen_modulators(nsagp, gpu_lib, itts):
    hyp, y, mean, var, lml = exact_case(nsagp)
    T = y.size
    t = np.arange(1.0, T + 1.0)
    mom = nsagp.likModulatorNMFPower(nsagp.Softplus(0.0), 9, 2)
    ss = lambda x, p1, p2, k1, k2: nsagp.ss_modulators_nmf(p1, p2, k1, k2)
    E, V, _, _, _, out = nsagp.gf_ep_modulator_nmf(hyp.pack_log(), t, y, ss, mom, t, "matern32", "matern52", 1, 1, 2,
                                                  1.0, np.ones(itts), itts)
    assert np.allclose(E[0], mean, rtol=1e-6, atol=1e-8)
    assert np.allclose(V[0], var, rtol=1e-6, atol=1e-8)
    # the first (ADF) sweep's evidence is the exact one; the later entries of the reference are -sum(lZ) of the smoother-side
    # site update (gf_ep_modulator_nmf.m:277), whose cavities condition on all the other observations: not the evidence
    # of y, the oracle shows the same offset -- the posterior above stays
    # exact because at alpha = 1 the sites of a Gaussian likelihood are a fixed point of the update
    assert abs(-np.asarray(out["nlZ"]).ravel()[0] - lml) < 1e-6 * abs(lml)


@pytest.mark.parametrize("itts", [1, 3])
def test_gfep_power_ep_posterior_is_exact_for_a_gaussian_likelihood(nsagp, gpu_lib, itts):
    """likModulatorPreCalcwn ("sqrt" model, shifted link, alpha = 0.75): a = sqrt(sum_n W_n softplus(-1)).  Power EP
    reproduces a Gaussian factor exactly whatever alpha is, so the posterior is still the GP regression (the Power-EP
    energy is not the evidence, so nlZ is not compared)."""
    a = math.sqrt((0.7 + 0.4) * math.log1p(math.exp(-1.0)))
    hyp, y, mean, var, _ = exact_case(nsagp, a=a)
    T = y.size
    t = np.arange(1.0, T + 1.0)
    wn, xn = nsagp.utp_ws(9, 2)
    mom = nsagp.likModulatorPreCalcwn(nsagp.Softplus(1.0), wn, xn)
    ss = lambda x, p1, p2, k1, k2: nsagp.ss_modulators_nmf(p1, p2, k1, k2)
    E, V, _, _, _, out = nsagp.gf_ep_modulator_nmf(hyp.pack_log(), t, y, ss, mom, t, "matern32", "matern52", 1, 1, 2,
                                                  0.75, np.ones(itts), itts)
    assert np.allclose(E[0], mean, rtol=1e-6, atol=1e-8)
    assert np.allclose(V[0], var, rtol=1e-6, atol=1e-8)


def test_giekf_is_the_kalman_smoother_of_the_linear_model(nsagp, gpu_lib):
    """gf_giekf_modulator_nmf with one global and one local iteration on the same frozen-modulator model: the
    linearisation is exact, so the smoothed z is the GP regression and the energy is minus the Gaussian evidence."""
    hyp, y, mean, var, lml = exact_case(nsagp)
    T = y.size
    t = np.arange(1.0, T + 1.0)
    ss = lambda x, p1, p2, k1, k2: nsagp.ss_modulators_nmf(p1, p2, k1, k2)
    E, V = nsagp.gf_giekf_modulator_nmf(hyp.pack_log(), t, y, ss, None, t, "matern32", "matern52", 1, 1, 2, 1, 1)[:2]
    assert np.allclose(E[0], mean, rtol=1e-6, atol=1e-8)
    assert np.allclose(V[0], var, rtol=1e-6, atol=1e-8)
    e = nsagp.gf_giekf_modulator_nmf(hyp.pack_log(), t, y, ss, None, None, "matern32", "matern52", 1, 1, 2, 1, 1)[0]
    assert abs(e + lml) < 1e-6 * abs(lml)


def exact_case_subbands(nsagp, D=4, T=36, sn2=0.02, seed=9, k1="exp", k2="matern52"):
    """D subbands, N = 2 frozen modulators: y = sum_d a_d z_d + noise, a_d = sum_n W_dn log 2.  Exact for the EKF's joint
    update (not for EP, whose sites factorise over the latents)."""
    from oracle import ssmodel as oss
    rng = np.random.default_rng(seed)
    hyp = nsagp.synth.Hypers(sn2, var_fast=rng.uniform(0.3, 1.0, D), len_fast=rng.uniform(5.0, 30.0, D),
                             omega=np.linspace(1.0, 0.2, D), var_slow=np.array([1e-12, 1e-12]),
                             len_slow=np.array([20.0, 35.0]), W=rng.uniform(0.2, 1.0, (D, 2)))
    F, L, Qc, H, Pinf = oss.ss_modulators_nmf(hyp.w_sub(), hyp.w_mod(), k1, k2)
    A, Q = oss.lti_disc(F, L, Qc, 1.0)
    n = A.shape[0]
    a = hyp.W.sum(axis=1) * math.log(2.0)
    y = rng.normal(0, 0.5, T)
    Apow = [np.eye(n)]
    for _ in range(T):
        Apow.append(A @ Apow[-1])
    C = lambda s, t: Apow[s - t] @ Pinf if s >= t else Pinf @ Apow[t - s].T
    Kd = [np.array([[H[d] @ C(s, t) @ H[d] for t in range(T)] for s in range(T)]) for d in range(D)]
    S = sum(a[d] ** 2 * Kd[d] for d in range(D)) + sn2 * np.eye(T)
    mean = np.stack([a[d] * Kd[d] @ np.linalg.solve(S, y) for d in range(D)])
    var = np.stack([np.diag(Kd[d] - a[d] ** 2 * Kd[d] @ np.linalg.solve(S, Kd[d])) for d in range(D)])
    _, logdet = np.linalg.slogdet(S)
    lml = -0.5 * y @ np.linalg.solve(S, y) - 0.5 * logdet - 0.5 * T * math.log(2 * math.pi)
    return hyp, y, mean, var, lml


@pytest.mark.parametrize("D,k1", [(4, "exp"), (6, "matern32")])
def test_giekf_joint_update_is_exact_over_several_subbands(nsagp, gpu_lib, D, k1):
    hyp, y, mean, var, lml = exact_case_subbands(nsagp, D=D, k1=k1)
    T = y.size
    t = np.arange(1.0, T + 1.0)
    ss = lambda x, p1, p2, k1_, k2_: nsagp.ss_modulators_nmf(p1, p2, k1_, k2_)
    E, V = nsagp.gf_giekf_modulator_nmf(hyp.pack_log(), t, y, ss, None, t, k1, "matern52", 1, D, 2, 1, 1)[:2]
    assert np.allclose(E[:D], mean, rtol=1e-6, atol=1e-8)
    assert np.allclose(V[:D], var, rtol=1e-6, atol=1e-8)
    e = nsagp.gf_giekf_modulator_nmf(hyp.pack_log(), t, y, ss, None, None, k1, "matern52", 1, D, 2, 1, 1)[0]
    assert abs(e + lml) < 1e-6 * abs(lml)


def test_gfep_nlz_mode_returns_the_gaussian_evidence(nsagp, gpu_lib):
    """xt empty: the entry point's first output is nlZ (gf_ep_modulator_nmf.m:357-533)."""
    hyp, y, _, _, lml = exact_case(nsagp)
    t = np.arange(1.0, y.size + 1.0)
    mom = nsagp.likModulatorNMFPower(nsagp.Softplus(0.0), 9, 2)
    ss = lambda x, p1, p2, k1, k2: nsagp.ss_modulators_nmf(p1, p2, k1, k2)
    nlZ = nsagp.gf_ep_modulator_nmf(hyp.pack_log(), t, y, ss, mom, None, "matern32", "matern52", 1, 1, 2, 1.0, np.ones(1), 1)[0]
    assert abs(float(np.ravel(nlZ)[0]) + lml) < 1e-6 * abs(lml)


def test_filterbank_equals_the_time_varying_kalman_smoother_in_the_interior(nsagp, gpu_lib):
    """kernel_ss_kalmanFastFB (stationary gains): away from the ends of a long signal its filter and smoother means are
    those of the ordinary Kalman filter / RTS smoother of the same model (a textbook recursion, tests/test_oracle_filterbank.py)."""
    import importlib
    from test_oracle_filterbank import _kalman_rts
    from conftest import rel_err
    fb = importlib.import_module(nsagp.__name__ + ".filterbank")
    rng = np.random.default_rng(1)
    D, T = 3, 1200
    lamx, varx, om = 1.0 / rng.uniform(20, 60, D), rng.uniform(0.3, 1.0, D), np.array([0.9, 0.5, 0.2])
    A, Q, H, Pinf, K, tau = fb.get_disc_model(lamx, varx, om, D, "matern32")
    y = np.cos(0.5 * np.arange(T)) + 0.1 * rng.standard_normal(T)
    MF, MS = _kalman_rts(A, Q, H, Pinf, 0.01, y)
    _, Xf, _ = fb.kernel_ss_kalmanFastFB(A, Q, H, Pinf, K, 0.01, y, 0, 1)
    _, Xs, _ = fb.kernel_ss_kalmanFastFB(A, Q, H, Pinf, K, 0.01, y, 0, 0)
    mid = slice(500, 700)
    assert rel_err(Xf[0][:, mid], MF[:, mid]) < 1e-6
    assert rel_err(Xs[0][:, mid], MS[:, mid]) < 1e-6
