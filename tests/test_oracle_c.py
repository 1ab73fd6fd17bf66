"""The plain-C restatement (oracle/c/nsagp_oracle.c) against the NumPy oracle and the
committed golden fixtures: two independently written restatements of
matlab/ihgp_ep_modulator_nmf.m and matlab/likModulatorNMFPower.m must agree to rounding."""
import numpy as np
import pytest

from conftest import make_problem, rel_err


def test_c_mom_matches_numpy_oracle(nsagp):
    from oracle import c_oracle, cubature as ocub, lik as olik
    rng = np.random.default_rng(4)
    for kind, D, N, p, shift, alpha in [("power", 5, 2, 9, 0.0, 0.5), ("precalc", 16, 3, 9, 1.0, 0.75),
                                        ("power", 3, 3, 7, 0.0, 1.0)]:
        W = rng.uniform(0.05, 0.6, (D, N))
        wn, xn = ocub.utp_ws(p, N)
        mom = olik.make_mom(kind, olik.softplus_link(shift), p=p, wn=wn, xn_unscaled=xn)
        for _ in range(20):
            mu = np.concatenate([rng.normal(0, 0.3, D), rng.normal(0, 1.5, N)])
            s2 = np.concatenate([rng.uniform(1e-3, 0.1, D), rng.uniform(0.05, 2.0, N)])
            y = rng.normal(0, 0.3)
            lo, d1o, d2o = mom(np.log([1e-3]), mu, s2, W, alpha, np.array([y]), 0)
            lc, d1c, d2c = c_oracle.mom(1 if kind == "precalc" else 0, np.log(1e-3), shift, W, wn, xn, alpha, y, mu, s2)
            assert abs(lc - lo) < 1e-12 * max(1.0, abs(lo))
            assert rel_err(d1c, d1o) < 1e-11 and rel_err(d2c, d2o) < 1e-11
    # missing sample: NaN moments, lZ = log(pEP * jitter) (MATLAB max(NaN, jitter) = jitter)
    lc, d1c, d2c = c_oracle.mom(0, np.log(1e-3), 0.0, W, wn, xn, 1.0, np.nan, mu, s2)
    assert abs(lc - np.log(1e-10)) < 1e-12 and np.all(np.isnan(d1c)) and np.all(np.isnan(d2c))


def _c_problem(nsagp, pb, alpha, damping, itts, want_smoother=True):
    from oracle import c_oracle, cubature as ocub, ihgp_ep, ssmodel as oss
    D, N = pb["D"], pb["N"]
    lik_param, p1, p2, W = oss.unpack_log(pb["w"], 1, D, N)
    A, Q, H, Pinf = ihgp_ep._model(lik_param, p1, p2, pb["ss_ref"], pb["t"], pb["kernel1"], pb["kernel2"])
    tabs = ihgp_ep.ihgp_setup(A, Q, H, want_smoother=want_smoother)
    mg = pb["mom_gpu"]
    return c_oracle.IhgpProblem(A, H, Pinf, tabs, mg.kind, lik_param, mg.link.shift, W, mg.wn, mg.xn, alpha, damping, itts)


@pytest.mark.parametrize("kind,gaps,itts", [("power", False, 3), ("precalc", True, 3), ("precalc", False, 1)])
def test_c_ihgp_matches_numpy_oracle(nsagp, kind, gaps, itts):
    from oracle import ihgp_ep
    D, N, T = (4, 2, 200) if kind == "power" else (6, 3, 220)
    k1, k2 = ("matern32", "matern52") if kind == "power" else ("exp", "matern52")
    shift, alpha = (0.0, 0.5) if kind == "power" else (1.0, 0.75)
    pb = make_problem(nsagp, D, N, T, k1, k2, seed=31, kind=kind, p=9, shift=shift, gaps=gaps)
    damping = np.linspace(0.5, 0.3, itts)
    cp = _c_problem(nsagp, pb, alpha, damping, itts)
    rc = cp.predict(pb["y"])
    Eo, Vo, _, _, _, oo = ihgp_ep.ihgp_ep_modulator_nmf(pb["w"], pb["t"], pb["y"], pb["ss_ref"], pb["mom_ref"], pb["t"],
                                                        k1, k2, 1, D, N, alpha, damping, itts)
    tol = 1e-9
    assert rel_err(rc["nlZ"], oo["nlZ"]) < tol
    assert rel_err(rc["Eft"], Eo) < tol and rel_err(rc["Varft"], Vo) < tol
    assert rel_err(rc["ttau"], oo["ttau"]) < tol and rel_err(rc["tnu"], oo["tnu"]) < tol
    assert rel_err(rc["R"], oo["R"]) < tol and rel_err(rc["MS"], oo["MS"]) < tol
    assert rc["n_negcav"] == oo["n_negcav"]
    assert rel_err(rc["maxDiffM"], oo["maxDiffM"]) < 1e-7
    # nlZ mode, plain and running-site (_constraints) variants
    e_c, _ = _c_problem(nsagp, pb, alpha, damping[:1], 1, want_smoother=False).nlz(pb["y"], running=False)
    e_o, _ = ihgp_ep.ihgp_ep_modulator_nmf(pb["w"], pb["t"], pb["y"], pb["ss_ref"], pb["mom_ref"], None, k1, k2, 1, D, N,
                                           alpha, damping[:1], 1)
    assert abs(e_c - e_o) < tol * abs(e_o)
    from oracle import ssmodel as oss
    lik_param, p1, p2, W = oss.unpack_log(pb["w"], 1, D, N)
    A, Q, H, Pinf = ihgp_ep._model(lik_param, p1, p2, pb["ss_ref"], pb["t"], k1, k2)
    e_o2, _ = ihgp_ep.ihgp_nlz_core(A, Q, H, Pinf, lik_param, W, pb["y"], pb["mom_ref"], damping[:1], running_sites=True)
    e_c2, _ = _c_problem(nsagp, pb, alpha, damping[:1], 1, want_smoother=False).nlz(pb["y"], running=True)
    assert abs(e_c2 - e_o2) < tol * abs(e_o2)


@pytest.mark.parametrize("name", ["ihgp_demo_small", "ihgp_c2_small"])
def test_c_ihgp_reproduces_golden(nsagp, name):
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    D, N, T = int(g["D"]), int(g["N"]), int(g["T"])
    from oracle import ssmodel as oss
    kind = str(g["kind"])
    wn, xn = nsagp.utp_ws(int(g["p"]), N)
    if kind == "power":
        mg = nsagp.likModulatorNMFPower(nsagp.Softplus(float(g["shift"])), int(g["p"]), N)
    else:
        mg = nsagp.likModulatorPreCalcwn(nsagp.Softplus(float(g["shift"])), wn, xn)
    pb = dict(D=D, N=N, w=g["w"], t=g["t"], kernel1=str(g["kernel1"]), kernel2=str(g["kernel2"]), mom_gpu=mg,
              ss_ref=lambda x, p1, p2, a, b: oss.ss_modulators_nmf(p1, p2, a, b))
    rc = _c_problem(nsagp, pb, float(g["alpha"]), g["damping"], int(g["itts"])).predict(g["y"])
    tol = 1e-9
    assert rel_err(rc["Eft"], g["Eft"]) < tol and rel_err(rc["Varft"], g["Varft"]) < tol
    assert rel_err(rc["ttau"], g["ttau"]) < tol and rel_err(rc["tnu"], g["tnu"]) < tol
    assert rel_err(rc["R"], g["R"]) < tol and rel_err(rc["MS"], g["MS"]) < tol
    assert rel_err(rc["nlZ"], g["nlZ"]) < tol and rc["n_negcav"] == int(g["n_negcav"])
