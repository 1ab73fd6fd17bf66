"""GPU checks at BASELINE.json's full sizes (C2: D=16, N=3, S=77, T=100000), where the CPU
oracle would take hours.  Size-independent properties stand in for it:

* two independent implementations of the sequential pass (one CTA per signal with the fused
  update and straight-line math, vs one warp per signal with the library math and the
  reference's literal update order) agree to the sequential tolerance at full length;
* causality: the filtered means and sites of the first T' steps do not depend on what follows;
* determinism: two runs give identical bits;
* the oracle itself is compared on a prefix short enough for it to finish in seconds.
"""
import importlib

import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu
D, N, T = 16, 3, 100000
K1, K2 = "exp", "matern52"


@pytest.fixture(scope="module")
def c2(nsagp):
    rng = np.random.default_rng(2026)
    hyp = nsagp.synth.speech_hypers(D, N, rng)
    y, _, _ = nsagp.synth.sample_signal(hyp, K1, K2, T, rng, link_shift=1.0, sqrt_model=True)
    F, L, Qc, H, Pinf = nsagp.ss_modulators_nmf(hyp.w_sub(), hyp.w_mod(), K1, K2)[:5]
    F, L, H, Pinf = nsagp.ssmodel.balance(F, L, H, Pinf)
    A, Q = nsagp.lti_disc(F, L, Qc, 1.0)
    Q = (Q + Q.T) / 2
    mdl = nsagp.to_block_model(A, Q, H, Pinf, D, N)
    tabs = nsagp.tables.build_tables(mdl, want_smoother=True)
    wn, xn = nsagp.utp_ws(9, N)
    mom = nsagp.likModulatorPreCalcwn(nsagp.Softplus(1.0), wn, xn)
    return dict(hyp=hyp, y=y, mdl=mdl, tabs=tabs, mom=mom, lik_param=np.log([hyp.w_lik]))


def _run(nsagp, c2, y, itts, form, names):
    L = nsagp._lib
    damp = np.linspace(0.05, 0.1, max(itts, 1))
    with nsagp.Plan(L.KIND_IHGP, [c2["mdl"]], [(c2["mom"], c2["lik_param"], c2["hyp"].W)], 0.75, damp, itts, y[None, :],
                    L.MODE_PREDICT, tables=[c2["tabs"]]) as plan:
        plan.set_adf_form(form)
        plan.run()
        return plan.fetch(0, names)


def test_c2_full_length_two_implementations_agree(nsagp, gpu_lib, c2):
    names = ("Eft", "nlZ", "ttau", "tnu", "R", "MS", "n_negcav")
    a = _run(nsagp, c2, c2["y"], 2, 0, names)
    b = _run(nsagp, c2, c2["y"], 2, 1, names)
    assert rel_err(a["nlZ"], b["nlZ"]) < 1e-8
    assert rel_err(a["Eft"], b["Eft"]) < 1e-6 and rel_err(a["MS"], b["MS"]) < 1e-6
    assert rel_err(a["ttau"], b["ttau"]) < 1e-6 and rel_err(a["R"], b["R"]) < 1e-6
    assert a["n_negcav"] == b["n_negcav"]
    assert np.all(np.isfinite(a["Eft"])) and np.all(a["ttau"] >= 0)


def test_c2_full_length_causality_and_determinism(nsagp, gpu_lib, c2):
    names = ("MF", "ttau", "tnu", "R", "lZ")
    full = _run(nsagp, c2, c2["y"], 1, 0, names)
    again = _run(nsagp, c2, c2["y"], 1, 0, names)
    for k in names:
        assert np.array_equal(full[k], again[k], equal_nan=True), k            # bit-identical
    Tp = 12345
    part = _run(nsagp, c2, c2["y"][:Tp], 1, 0, names)
    # the ADF pass is causal: a prefix of the signal gives the prefix of the result, bit for bit
    # (except the last step, whose moments the reference recomputes in every pass)
    for k in ("MF", "ttau", "tnu", "R"):
        assert np.array_equal(full[k][:, :Tp - 1], part[k][:, :Tp - 1], equal_nan=True), k
    assert np.array_equal(full["lZ"][:Tp - 1], part["lZ"][:Tp - 1])


def test_c2_missing_data_semantics_at_scale(nsagp, gpu_lib, c2):
    """NaN samples: the IHGP filter does not skip them (ihgp_ep_modulator_nmf.m:253-271); the sites
    of a missing step are ttau = 0, tnu = NaN, R = Inf and the step adds log(pEP 1e-10) to lZ."""
    y = c2["y"].copy()
    gaps = np.zeros(T, bool)
    rng = np.random.default_rng(5)
    for s in rng.integers(100, T - 400, 30):
        gaps[s:s + int(rng.integers(10, 320))] = True
    y[gaps] = np.nan
    r = _run(nsagp, c2, y, 1, 0, ("ttau", "tnu", "R", "lZ", "MF"))
    g = gaps.copy(); g[-1] = False
    assert np.all(r["ttau"][:, g] == 0) and np.all(np.isnan(r["tnu"][:, g])) and np.all(np.isinf(r["R"][:, g]))
    assert np.allclose(r["lZ"][g], np.log(1e-10))                              # alpha = 1 in the filter: pEP = 1
    assert np.all(np.isfinite(r["MF"]))
    b = _run(nsagp, c2, y, 1, 1, ("ttau", "MF"))
    assert rel_err(r["MF"], b["MF"]) < 1e-8 and rel_err(r["ttau"], b["ttau"]) < 1e-8


def test_c2_prefix_against_oracle(nsagp, gpu_lib, c2):
    """The C restatement of the reference on the first 1500 samples of the full-size problem."""
    from oracle import c_oracle, cubature as ocub, ihgp_ep, ssmodel as oss
    Tp, itts = 1500, 3
    hyp = c2["hyp"]
    damp = np.linspace(0.05, 0.1, itts)
    lik_param, p1, p2, W = oss.unpack_log(hyp.pack_log(), 1, D, N)
    ss = lambda x, a, b, k1, k2: oss.ss_modulators_nmf(a, b, k1, k2)
    A, Q, H, Pinf = ihgp_ep._model(lik_param, p1, p2, ss, np.arange(1.0, Tp + 1), K1, K2)
    tabs = ihgp_ep.ihgp_setup(A, Q, H)
    wo, xo = ocub.utp_ws(9, N)
    ref = c_oracle.IhgpProblem(A, H, Pinf, tabs, 1, lik_param, 1.0, W, wo, xo, 0.75, damp, itts).predict(c2["y"][:Tp])
    L = nsagp._lib
    with nsagp.Plan(L.KIND_IHGP, [c2["mdl"]], [(c2["mom"], c2["lik_param"], hyp.W)], 0.75, damp, itts, c2["y"][None, :Tp],
                    L.MODE_PREDICT, tables=[c2["tabs"]]) as plan:
        plan.run()
        got = plan.fetch(0, ("Eft", "nlZ", "ttau", "MS"))
    assert rel_err(got["nlZ"], ref["nlZ"]) < 1e-6
    assert rel_err(got["Eft"], ref["Eft"]) < 1e-6 and rel_err(got["MS"], ref["MS"]) < 1e-6
    assert rel_err(got["ttau"], ref["ttau"]) < 1e-6


# ----------------------------------------------------------------------------- C4 (iterated EKF, dense n = 73)
def _c4_problem(nsagp, T, seed=3):
    rng = np.random.default_rng(seed)
    Dk, Nk = 32, 3
    hyp = nsagp.synth.speech_hypers(Dk, Nk, rng, w_lik=1e-2)
    y, _, _ = nsagp.synth.sample_signal(hyp, K1, K2, T, rng)
    for s in range(0, T, 20000):                              # six gaps of 10..320 samples per 20k samples
        for j, g in enumerate((10, 20, 40, 80, 160, 320)):
            a = s + 1500 + 3000 * j
            y[a:min(a + g, T)] = np.nan
    return hyp, y, Dk, Nk


def _c4_run(nsagp, hyp, y, Dk, Nk, form, chunk_len=0, seg_chunks=0, g_iter=1, cov=False):
    L = nsagp._lib
    ss = lambda x, p1, p2, k1, k2: nsagp.ss_modulators_nmf(p1, p2, k1, k2)
    t = np.arange(1.0, y.size + 1.0)
    L.check(L.lib().nsagp_giekf_config(form, chunk_len, seg_chunks))
    try:
        return nsagp.gf_giekf_modulator_nmf(hyp.pack_log(), t, y, ss, None, t, K1, K2, 1, Dk, Nk, g_iter, 1, debug_cov=cov)
    finally:
        L.check(L.lib().nsagp_giekf_config(0, 0, 0))


def test_c4_scan_smoother_matches_sequential_kernels_at_n73(nsagp, gpu_lib):
    """BASELINE config C4's shape (D=32 exp subbands, N=3 matern52 modulators: dense n = 73, missing-data gaps).  The
    scan smoother on the FP64 tensor cores (several chunks, several segments) against the first-generation kernels
    (sequential Cholesky / solves / products): two independent implementations, 1e-6 class, observed ~1e-12."""
    hyp, y, Dk, Nk = _c4_problem(nsagp, 6000)
    Ea, Va, _, lba, uba, oa = _c4_run(nsagp, hyp, y, Dk, Nk, 1, cov=True)
    Eb, Vb, _, lbb, ubb, ob = _c4_run(nsagp, hyp, y, Dk, Nk, 2, chunk_len=37, seg_chunks=50, cov=True)
    assert rel_err(Eb, Ea) < 1e-9 and rel_err(Vb, Va) < 1e-9 and rel_err(ob["MS"], oa["MS"]) < 1e-9
    assert rel_err(ob["PF"], oa["PF"]) < 1e-8 and rel_err(ob["PS"], oa["PS"]) < 1e-9
    assert rel_err(lbb, lba) < 1e-8 and rel_err(ubb, uba) < 1e-8


def test_c4_long_signal_properties(nsagp, gpu_lib):
    """T = 60 000 (three gap patterns): finite, positive marginal variances; smoothing never increases a marginal
    variance; inside a gap the filtered variance grows and the smoothed one is bounded by the prior's; two runs with
    different chunking of the scan agree to rounding."""
    T = 60000
    hyp, y, Dk, Nk = _c4_problem(nsagp, T, seed=8)
    E1, V1, _, _, _, o1 = _c4_run(nsagp, hyp, y, Dk, Nk, 2)
    E2, V2, _, _, _, o2 = _c4_run(nsagp, hyp, y, Dk, Nk, 2, chunk_len=50, seg_chunks=120)
    assert np.all(np.isfinite(E1)) and np.all(V1 > 0)
    assert rel_err(E2, E1) < 1e-9 and rel_err(V2, V1) < 1e-9
    # filtered marginal variances from a filter-only view: rerun with the sequential form on a prefix is too slow at
    # this size; use the smoothed-vs-prior bound instead: var_smoothed <= prior variance h Pinf h'
    F, Lm, Qc, H, Pinf = nsagp.ss_modulators_nmf(hyp.w_sub(), hyp.w_mod(), K1, K2)[:5]
    F, Lm, H, Pinf = nsagp.ssmodel.balance(F, Lm, H, Pinf)
    prior = np.einsum("ij,jk,ik->i", H, Pinf, H)
    assert np.all(V1 <= prior[:, None] * (1 + 1e-9))
    gap = np.isnan(y)
    # a subband's marginal is far less certain in the middle of the longest gap than right before it
    k_mid = 1500 + 3000 * 5 + 160
    assert gap[k_mid] and not gap[1500 + 3000 * 5 - 1]
    assert np.all(V1[:Dk, k_mid] > V1[:Dk, 1500 + 3000 * 5 - 5])
