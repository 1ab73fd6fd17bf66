"""GPU parity for the model WITHOUT NMF weights, matlab/gf_ep_modulator.m (SURVEY.md section 8 "next" row N4, the
non-NMF half): D carrier x modulator pairs, y = sum_d z_d softplus(g_d), likelihood matlab/likModulatorPower.m, state
space matlab/ss_modulators.m -- the configuration of matlab/demo_toy_modulators.m (D = 2, matern32 x matern52,
ep_fraction 0.5, ep_itts 5, damping 0.3, 9th-degree cubature) and two variations, predict and nlZ modes, against the
oracle's restatement of that file.  Tolerance 1e-8 for the sequential first pass, 1e-6 after the scans."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _problem(nsagp, D, T, k1, k2, seed, w_lik):
    from importlib import import_module
    synth = import_module(nsagp.__name__ + ".synth")
    rng = np.random.default_rng(seed)
    hyp = synth.Hypers(w_lik,
                       var_fast=np.array([0.1, 0.1, 0.4, 0.2][:D]), len_fast=np.array([50.0, 40.0, 20.0, 30.0][:D]),
                       omega=np.array([np.pi / 4, np.pi / 6, np.pi / 8, np.pi / 10][:D]),
                       var_slow=np.array([2.0, 3.0, 4.0, 2.5][:D]), len_slow=np.array([500.0, 700.0, 90.0, 300.0][:D]),
                       W=np.eye(D))                                          # demo_toy_modulators.m:12-16
    y, _, _ = synth.sample_signal(hyp, k1, k2, T, rng)
    w = np.log(np.concatenate([[w_lik], hyp.var_fast, hyp.len_fast, hyp.omega, hyp.var_slow, hyp.len_slow]))
    return w, y, np.arange(1.0, T + 1.0)


@pytest.mark.parametrize("D,k1,k2,p", [(2, "matern32", "matern52", 9),      # the demo
                                       (3, "exp", "matern32", 7),
                                       (4, "matern32", "matern52", 5)])
def test_gf_ep_modulator_matches_oracle(nsagp, gpu_lib, D, k1, k2, p):
    from oracle import gf_ep, lik as olik, ssmodel as oss
    T = 400
    w, y, t = _problem(nsagp, D, T, k1, k2, seed=123, w_lik=1e-3)
    link = olik.softplus_link(0.0)
    seen_Z = []

    def mom_ref(hyp, mu, s2, ep_frac, yall, k):
        out = olik.likModulatorPower(link, hyp, yall[k], mu, s2, p, ep_frac)
        seen_Z.append(out[0])
        return out

    ss_ref = lambda x, pr, a, b: oss.ss_modulators(pr, a, b)
    ss_gpu = lambda x, pr, a, b: nsagp.ss_modulators(pr, a, b)
    mom_gpu = nsagp.likModulatorPower(nsagp.Softplus(0.0), p, D)
    damping = np.linspace(0.3, 0.3, 5)
    Eo, Vo, _, lbo, ubo, oo = gf_ep.gf_ep_modulator(w, t, y, ss_ref, mom_ref, t, k1, k2, 1, 0.5, damping, 5)
    Eg, Vg, _, lbg, ubg, og = nsagp.gf_ep_modulator(w, t, y, ss_gpu, mom_gpu, t, k1, k2, 1, 0.5, damping, 5)
    # the one documented deviation (floor under Z: 1e-10 here, 1e-8 in likModulatorPower.m:29) must not be in play
    assert np.exp(min(seen_Z)) > 1e-7
    assert Eg.shape == Eo.shape == (2 * D, T)
    assert rel_err(Eg, Eo) < 1e-6 and rel_err(Vg, Vo) < 1e-6 and rel_err(lbg, lbo) < 1e-6 and rel_err(ubg, ubo) < 1e-6
    assert rel_err(og["ttau"], oo["ttau"]) < 1e-6 and rel_err(og["lZ"], oo["lZ"]) < 1e-6
    assert rel_err(og["nlZ"], oo["nlZ"]) < 1e-6
    assert abs(og["nlZ"][0] - oo["nlZ"][0]) < 1e-8 * abs(oo["nlZ"][0])       # the sequential first pass
    # nlZ mode (:360-538): what fminunc minimises in demo_toy_modulators.m:99
    eo, go = gf_ep.gf_ep_modulator(w, t, y, ss_ref, mom_ref, None, k1, k2, 1, 0.5, damping, 3)
    eg, gg = nsagp.gf_ep_modulator(w, t, y, ss_gpu, mom_gpu, None, k1, k2, 1, 0.5, damping, 3)
    assert abs(eg - eo) < 1e-6 * abs(eo) and gg.shape == go.shape == w.shape and not np.any(gg)


def test_gf_ep_modulator_rejects_more_pairs_than_the_cubature_supports(nsagp, gpu_lib):
    D = 5
    with pytest.raises((ValueError, nsagp._lib.NsagpError)):
        mom = nsagp.likModulatorPower(nsagp.Softplus(0.0), 9, D)
        w = np.zeros(1 + 5 * D)
        t = np.arange(1.0, 51.0)
        nsagp.gf_ep_modulator(w, t, np.zeros(50), lambda x, pr, a, b: nsagp.ss_modulators(pr, a, b), mom, t,
                              "exp", "exp", 1, 0.5, [0.3], 1)
