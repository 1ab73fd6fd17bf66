"""GPU parity: gf_ep_modulator_nmf[_constraints] through the C ABI vs the oracle
restatement of matlab/gf_ep_modulator_nmf.m (dense n-by-n arithmetic) on the same
seeded inputs.

Tolerances (north_star): 1e-8 relative for the sequential filter pass, 1e-6 for
what goes through the re-associated smoother scan."""
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
    (4, 2, 300, "matern32", "matern52", "power", 9, 0.0, 0.5, 3, False),     # demo_toy_modulators_nmf, small
    (6, 3, 400, "exp", "matern52", "precalc", 9, 1.0, 0.75, 3, False),       # C3 likelihood/kernels, small
    (6, 3, 400, "exp", "matern52", "precalc", 9, 1.0, 0.75, 3, True),        # missing-data gaps
    (5, 2, 260, "matern52", "matern32", "power", 7, 0.0, 0.5, 2, False),     # 6x6 subband blocks
    (3, 2, 200, "matern72", "exp", "power", 5, 0.0, 1.0, 2, False),          # 8x8 and 1x1 blocks
    (16, 3, 200, "exp", "matern52", "precalc", 9, 1.0, 0.75, 2, False),      # C3's own shape (M = 19 lanes per chunk in the scans)
]


def _dense_blocks(packed, sizes):
    """(sum b^2, T) packed block covariances -> list over blocks of (b, b, T)."""
    out, off = [], 0
    for b in sizes:
        out.append(packed[off:off + b * b].reshape((b, b, -1), order="F"))
        off += b * b
    return out


@pytest.mark.parametrize("adf_form", [0, 1, 2])   # one CTA per signal / one warp per signal / half-width CTA (two per SM)
@pytest.mark.parametrize("D,N,T,k1,k2,kind,p,shift,alpha,itts,gaps", CASES)
def test_gfep_predict_matches_oracle(nsagp, gpu_lib, D, N, T, k1, k2, kind, p, shift, alpha, itts, gaps, adf_form):
    from oracle import gf_ep
    pb = make_problem(nsagp, D, N, T, k1, k2, seed=21 + D + T, kind=kind, p=p, shift=shift, gaps=gaps)
    damping = np.linspace(0.5, 0.3, itts)
    Eo, Vo, _, lbo, ubo, oo = gf_ep.gf_ep_modulator_nmf(*_args(pb, "ref", pb["t"], alpha, damping, itts), want_cov=True)
    Eg, Vg, Cg, lbg, ubg, og = nsagp.gf_ep_modulator_nmf(*_args(pb, "gpu", pb["t"], alpha, damping, itts), debug_cov=True,
                                                      adf_form=adf_form)
    assert Cg is None
    tol = TOL_SEQ if itts == 1 else TOL_SCAN
    assert rel_err(og["nlZ"], oo["nlZ"]) < tol
    assert rel_err(og["lZ"], oo["lZ"]) < tol
    assert rel_err(Eg, Eo) < TOL_SCAN
    assert rel_err(Vg, Vo) < TOL_SCAN
    assert rel_err(lbg, lbo) < TOL_SCAN and rel_err(ubg, ubo) < TOL_SCAN
    assert rel_err(og["ttau"], oo["ttau"]) < tol
    assert rel_err(og["tnu"], oo["tnu"]) < tol
    assert rel_err(og["R"], oo["R"]) < tol
    assert rel_err(og["MF"], oo["MF"]) < tol
    assert rel_err(og["MS"], oo["MS"]) < TOL_SCAN
    assert rel_err(og["maxDiffM"], oo["maxDiffM"]) < 1e-5
    assert rel_err(og["maxDiffP"], oo["maxDiffP"]) < 1e-5
    assert og["n_negcav"] == oo["n_negcav"]
    # block covariances against the diagonal blocks of the oracle's dense n x n x T
    mdl = nsagp.to_block_model(*_dense_model(nsagp, pb), D, N)
    st = mdl.starts()
    for name in ("PF", "PS"):
        blocks = _dense_blocks(og[name], mdl.block_sizes())
        for i, blk in enumerate(blocks):
            ref = oo[name][st[i]:st[i + 1], st[i]:st[i + 1], :]
            assert rel_err(blk, ref) < TOL_SCAN
    # the reference's dense covariance is exactly block diagonal: nothing is lost
    mask = np.ones(oo["PS"].shape[:2], bool)
    for i in range(D + N):
        mask[st[i]:st[i + 1], st[i]:st[i + 1]] = False
    assert np.all(oo["PS"][mask] == 0.0)


def _dense_model(nsagp, pb):
    hyp = pb["hyp"]
    F, L, Qc, H, Pinf = nsagp.ss_modulators_nmf(hyp.w_sub(), hyp.w_mod(), pb["kernel1"], pb["kernel2"])[:5]
    A, Q = nsagp.lti_disc(F, L, Qc, 1.0)
    return A, Q, H, Pinf


@pytest.mark.parametrize("itts", [1, 3])
def test_gfep_nlz_matches_oracle(nsagp, gpu_lib, itts):
    from oracle import gf_ep
    pb = make_problem(nsagp, 5, 2, 350, "matern32", "matern52", seed=8, kind="power", p=9)
    damping = np.linspace(0.5, 0.4, itts)
    eo, go = gf_ep.gf_ep_modulator_nmf(*_args(pb, "ref", None, 0.5, damping, itts))
    eg, gg = nsagp.gf_ep_modulator_nmf(*_args(pb, "gpu", None, 0.5, damping, itts))
    assert abs(eg - eo) < (TOL_SEQ if itts == 1 else TOL_SCAN) * abs(eo)
    assert np.all(gg == 0) and gg.shape == go.shape


def test_gfep_constraints_matches_oracle(nsagp, gpu_lib):
    from oracle import gf_ep
    pb = make_problem(nsagp, 5, 2, 300, "matern32", "matern52", seed=9, kind="power", p=9)
    hyp = pb["hyp"]
    cons = np.array([[0.0, 0.1], [50.0, 1000.0], [0.0, 3.2], [0.0, 20.0], [100.0, 3000.0], [0.0, 1.25]])
    parts = [hyp.var_fast, hyp.len_fast, hyp.omega, hyp.var_slow, hyp.len_slow, hyp.W.reshape(-1, order="F")]
    wc = np.concatenate([np.log([hyp.w_lik])] + [nsagp.inv_sigmoid(v, c) for v, c in zip(parts, cons)])
    tune = [0, 1, 1, 0, 1, 1, 0]
    idx = np.cumsum([0, 1, 5, 5, 5, 2, 2, 10])
    w = np.concatenate([wc[idx[i]:idx[i + 1]] for i in range(7) if tune[i]])
    wf = np.concatenate([wc[idx[i]:idx[i + 1]] for i in range(7) if not tune[i]])
    damping = [0.5, 0.5]
    a = list(_args(pb, "ref", pb["t"], 0.5, damping, 2)); a[0] = w
    Eo, Vo, _, _, _, oo = gf_ep.gf_ep_modulator_nmf_constraints(*a, cons, wf, tune)
    a = list(_args(pb, "gpu", pb["t"], 0.5, damping, 2)); a[0] = w
    Eg, Vg, _, _, _, og = nsagp.gf_ep_modulator_nmf_constraints(*a, cons, wf, tune)
    assert rel_err(Eg, Eo) < TOL_SCAN and rel_err(Vg, Vo) < TOL_SCAN
    assert rel_err(og["nlZ"], oo["nlZ"]) < TOL_SCAN
    a = list(_args(pb, "ref", None, 0.5, damping, 2)); a[0] = w
    eo, _ = gf_ep.gf_ep_modulator_nmf_constraints(*a, cons, wf, tune)
    a = list(_args(pb, "gpu", None, 0.5, damping, 2)); a[0] = w
    eg, _ = nsagp.gf_ep_modulator_nmf_constraints(*a, cons, wf, tune)
    assert abs(eg - eo) < TOL_SCAN * abs(eo)
