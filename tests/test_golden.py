"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py).

CPU: the oracle reproduces its committed vectors (regression pin).
GPU: the CUDA path, called through the reference-shaped entry points, matches the
same committed vectors (1e-6: everything here went through the scan kernels)."""
import importlib
import os

import numpy as np
import pytest

from conftest import rel_err

HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = ["gfep_demo_small", "gfep_c3_small", "ihgp_demo_small", "ihgp_c2_small"]


def _load(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


def _call(g, impl, nsagp, xt):
    from oracle import cubature as ocub, gf_ep, ihgp_ep, lik as olik, ssmodel as oss
    D, N, p, shift = int(g["D"]), int(g["N"]), int(g["p"]), float(g["shift"])
    k1, k2, kind, entry = str(g["kernel1"]), str(g["kernel2"]), str(g["kind"]), str(g["entry"])
    if impl == "oracle":
        if kind == "power":
            mom = olik.make_mom("power", olik.softplus_link(shift), p=p)
        else:
            wo, xo = ocub.utp_ws(p, N)
            mom = olik.make_mom("precalc", olik.softplus_link(shift), wn=wo, xn_unscaled=xo)
        ss = lambda x, p1, p2, a, b: oss.ss_modulators_nmf(p1, p2, a, b)
        fn = gf_ep.gf_ep_modulator_nmf if entry == "gf_ep" else ihgp_ep.ihgp_ep_modulator_nmf
    else:
        if kind == "power":
            mom = nsagp.likModulatorNMFPower(nsagp.Softplus(shift), p, N)
        else:
            mom = nsagp.likModulatorPreCalcwn(nsagp.Softplus(shift), *nsagp.utp_ws(p, N))
        ss = lambda x, p1, p2, a, b: nsagp.ss_modulators_nmf(p1, p2, a, b)
        fn = nsagp.gf_ep_modulator_nmf if entry == "gf_ep" else nsagp.ihgp_ep_modulator_nmf
    return fn(g["w"], g["t"], g["y"], ss, mom, xt, k1, k2, 1, D, N, float(g["alpha"]), g["damping"], int(g["itts"]))


def _compare(res, nlz, g, tol):
    Eft, Varft, _, lb, ub, out = res
    assert rel_err(Eft, g["Eft"]) < tol and rel_err(Varft, g["Varft"]) < tol
    assert rel_err(lb, g["lb"]) < tol and rel_err(ub, g["ub"]) < tol
    assert rel_err(out["ttau"], g["ttau"]) < tol and rel_err(out["tnu"], g["tnu"]) < tol
    assert rel_err(out["R"], g["R"]) < tol and rel_err(out["MS"], g["MS"]) < tol
    assert rel_err(out["nlZ"], g["nlZ"]) < tol
    assert abs(nlz - float(g["nlz_mode"])) < tol * abs(float(g["nlz_mode"]))
    assert int(out["n_negcav"]) == int(g["n_negcav"])


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_golden(nsagp, name):
    g = _load(name)
    res = _call(g, "oracle", nsagp, g["t"])
    nlz, grad = _call(g, "oracle", nsagp, None)
    _compare(res, nlz, g, 1e-9)
    assert np.all(grad == 0)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_matches_golden(nsagp, gpu_lib, name):
    g = _load(name)
    res = _call(g, "cuda", nsagp, g["t"])
    nlz, grad = _call(g, "cuda", nsagp, None)
    _compare(res, nlz, g, 1e-6)
    assert np.all(grad == 0)
