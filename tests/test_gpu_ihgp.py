"""GPU parity: ihgp_ep_modulator_nmf[_constraints] through the C ABI vs the oracle
restatement of matlab/ihgp_ep_modulator_nmf.m on the same seeded inputs.

Tolerances (north_star): 1e-8 relative for the sequential ADF pass (same-order
arithmetic), 1e-6 for everything that goes through the re-associated scans
(filter passes >= 2 and all smoother passes)."""
import numpy as np
import pytest

from conftest import make_problem, rel_err

pytestmark = pytest.mark.gpu
TOL_SEQ = 1e-8
TOL_SCAN = 1e-6


def _args(pb, which, xt, alpha, damping, itts):
    return (pb["w"], pb["t"], pb["y"], pb["ss_" + which], pb["mom_" + which], xt, pb["kernel1"], pb["kernel2"],
            1, pb["D"], pb["N"], alpha, damping, itts)


CASES = [
    # D, N, T, k1, k2, kind, p, shift, alpha, itts, gaps
    (4, 2, 300, "matern32", "matern52", "power", 9, 0.0, 0.5, 1, False),
    (4, 2, 300, "matern32", "matern52", "power", 9, 0.0, 0.5, 3, False),
    (6, 3, 500, "exp", "matern52", "precalc", 9, 1.0, 0.75, 4, False),     # C2 likelihood/kernels, small
    (6, 3, 500, "exp", "matern52", "precalc", 9, 1.0, 0.75, 3, True),      # missing-data gaps (NaN moments)
    (5, 2, 260, "matern52", "matern32", "power", 7, 0.0, 0.5, 2, False),   # 6x6 subband blocks
    (3, 2, 200, "matern72", "exp", "power", 5, 0.0, 1.0, 2, False),        # 8x8 and 1x1 blocks
    (16, 3, 200, "exp", "matern52", "precalc", 9, 1.0, 0.75, 2, False),    # C2's own shape (D = 16, N = 3, S = 77)
]


@pytest.mark.parametrize("adf_form", [0, 1, 2])   # one CTA per signal / one warp per signal / half-width CTA (two per SM)
@pytest.mark.parametrize("D,N,T,k1,k2,kind,p,shift,alpha,itts,gaps", CASES)
def test_ihgp_predict_matches_oracle(nsagp, gpu_lib, D, N, T, k1, k2, kind, p, shift, alpha, itts, gaps, adf_form):
    from oracle import ihgp_ep
    pb = make_problem(nsagp, D, N, T, k1, k2, seed=11 + D + T, kind=kind, p=p, shift=shift, gaps=gaps)
    damping = np.linspace(0.5, 0.3, itts)
    Eo, Vo, _, lbo, ubo, oo = ihgp_ep.ihgp_ep_modulator_nmf(*_args(pb, "ref", pb["t"], alpha, damping, itts))
    Eg, Vg, Cg, lbg, ubg, og = nsagp.ihgp_ep_modulator_nmf(*_args(pb, "gpu", pb["t"], alpha, damping, itts), adf_form=adf_form)
    assert Cg is None
    tol = TOL_SEQ if itts == 1 else TOL_SCAN
    assert rel_err(og["nlZ"], oo["nlZ"]) < tol
    assert rel_err(Eg, Eo) < TOL_SCAN
    assert rel_err(Vg, Vo) < TOL_SCAN
    assert rel_err(lbg, lbo) < TOL_SCAN and rel_err(ubg, ubo) < TOL_SCAN
    assert rel_err(og["ttau"], oo["ttau"]) < tol
    # tnu keeps NaN at missing samples in the reference (never cleaned): patterns must agree
    assert rel_err(og["tnu"], oo["tnu"]) < tol
    assert rel_err(og["R"], oo["R"]) < tol
    assert rel_err(og["MF"], oo["MF"]) < TOL_SCAN
    assert rel_err(og["MS"], oo["MS"]) < TOL_SCAN
    assert rel_err(og["maxDiffM"], oo["maxDiffM"]) < 1e-5
    assert og["n_negcav"] == oo["n_negcav"]


@pytest.mark.parametrize("constrained", [False, True])
def test_ihgp_nlz_matches_oracle(nsagp, gpu_lib, constrained):
    from oracle import ihgp_ep
    pb = make_problem(nsagp, 5, 2, 400, "matern32", "matern52", seed=5, kind="power", p=9)
    damping = [0.5]
    if not constrained:
        eo, go = ihgp_ep.ihgp_ep_modulator_nmf(*_args(pb, "ref", None, 0.5, damping, 1))
        eg, gg = nsagp.ihgp_ep_modulator_nmf(*_args(pb, "gpu", None, 0.5, damping, 1))
    else:
        hyp = pb["hyp"]
        cons = np.array([[0.0, 0.1], [50.0, 1000.0], [0.0, 3.2], [0.0, 20.0], [100.0, 3000.0], [0.0, 1.25]])
        parts = [hyp.var_fast, hyp.len_fast, hyp.omega, hyp.var_slow, hyp.len_slow, hyp.W.reshape(-1, order="F")]
        wc = np.concatenate([np.log([hyp.w_lik])] + [nsagp.inv_sigmoid(v, c) for v, c in zip(parts, cons)])
        tune = [1, 1, 0, 0, 1, 0, 1]
        idx = np.cumsum([0, 1, 5, 5, 5, 2, 2, 10])
        w = np.concatenate([wc[idx[i]:idx[i + 1]] for i in range(7) if tune[i]])
        wf = np.concatenate([wc[idx[i]:idx[i + 1]] for i in range(7) if not tune[i]])
        a = list(_args(pb, "ref", None, 0.5, damping, 1)); a[0] = w
        eo, go = ihgp_ep.ihgp_ep_modulator_nmf_constraints(*a, cons, wf, tune)
        a = list(_args(pb, "gpu", None, 0.5, damping, 1)); a[0] = w
        eg, gg = nsagp.ihgp_ep_modulator_nmf_constraints(*a, cons, wf, tune)
    assert abs(eg - eo) < TOL_SEQ * abs(eo)
    assert np.all(gg == 0) and gg.shape == go.shape      # the reference's gradient is identically zero


def test_native_tables_end_to_end(nsagp, gpu_lib):
    """The library's own table routine (nsagp_ihgp_tables, doubling algorithm) instead of SciPy's generic Riccati
    solver: the tables differ by the solvers' accuracy (<= 1e-10 in the max norm, test_host_logic.py), the
    posterior by what that perturbation propagates to -- bounded here, not bit-level."""
    from conftest import make_problem, rel_err
    D, N, T = 4, 2, 300
    pb = make_problem(nsagp, D, N, T, "matern32", "matern52", seed=5, kind="power", p=9)
    out = {}
    for native in (False, True):
        nsagp.tables.DEFAULT_NATIVE = native
        try:
            out[native] = nsagp.ihgp_ep_modulator_nmf(pb["w"], pb["t"], pb["y"], pb["ss_gpu"], pb["mom_gpu"], pb["t"], "matern32",
                                                      "matern52", 1, D, N, 0.5, [0.5, 0.5, 0.5], 3)
        finally:
            nsagp.tables.DEFAULT_NATIVE = False
    e = rel_err(out[True][0], out[False][0]); v = rel_err(out[True][1], out[False][1])
    print("native vs SciPy tables: Eft %.2e Varft %.2e" % (e, v))
    assert e < 1e-5 and v < 1e-5
