"""GPU parity: gf_giekf_modulator_nmf_constraints (iterated EKF + dense RTS smoother) through
the C ABI vs the oracle restatement of matlab/gf_giekf_modulator_nmf_constraints.m and
matlab/iekf_update1.m on the same seeded inputs.  Tolerance 1e-8 for the sequential smoother
(same-order arithmetic; the dense Cholesky/solves differ from LAPACK's operation order by rounding
only) and the looser 1e-6 class for the re-associated scan smoother (csrc/ekfscan.cuh; observed ~1e-12).
Both smoother forms run every case; the scan with chunk lengths / segment sizes small enough that the
test signals span several chunks and several segments."""
import numpy as np
import pytest

from conftest import make_problem, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-8


def _constrained(nsagp, pb, D, N):
    hyp = pb["hyp"]
    cons = np.array([[0.0, 0.1], [50.0, 1000.0], [0.0, 3.2], [0.0, 20.0], [100.0, 3000.0], [0.0, 1.25]])
    parts = [hyp.var_fast, hyp.len_fast, hyp.omega, hyp.var_slow, hyp.len_slow, hyp.W.reshape(-1, order="F")]
    wc = np.concatenate([np.log([hyp.w_lik])] + [nsagp.inv_sigmoid(v, c) for v, c in zip(parts, cons)])
    tune = [1, 1, 0, 1, 0, 1, 1]
    idx = np.cumsum([0, 1, D, D, D, N, N, D * N])
    w = np.concatenate([wc[idx[i]:idx[i + 1]] for i in range(7) if tune[i]])
    wf = np.concatenate([wc[idx[i]:idx[i + 1]] for i in range(7) if not tune[i]])
    return w, wf, cons, tune


CASES = [
    # D, N, T, k1, k2, g_iter, l_iter, gaps
    (4, 2, 150, "matern32", "matern52", 1, 1, False),
    (4, 2, 150, "matern32", "matern52", 3, 2, False),
    (6, 3, 160, "exp", "matern52", 2, 1, True),           # C4 kernels, missing-data gaps
    (3, 2, 90, "matern72", "exp", 2, 1, False),           # 8x8 and 1x1 blocks
]


FORMS = [
    # smoother_form, chunk_len, chunks_per_segment, tolerance
    (1, 0, 0, TOL),             # sequential kernel
    (2, 0, 0, 1e-6),            # scan, defaults (one chunk at these lengths... T > 64: several)
    (2, 7, 3, 1e-6),            # scan, several chunks per segment, several segments, ragged last chunk
    (2, 1, 4, 1e-6),            # scan, one step per chunk
    (2, 1000, 1, 1e-6),         # scan, the whole signal in one chunk
]


@pytest.fixture
def giekf_form(nsagp):
    L = nsagp._lib
    yield lambda form, cl, sc: L.check(L.lib().nsagp_giekf_config(form, cl, sc))
    L.check(L.lib().nsagp_giekf_config(0, 0, 0))


@pytest.mark.parametrize("form,chunk_len,seg_chunks,tol", FORMS)
@pytest.mark.parametrize("D,N,T,k1,k2,g_iter,l_iter,gaps", CASES)
def test_giekf_predict_matches_oracle(nsagp, gpu_lib, giekf_form, D, N, T, k1, k2, g_iter, l_iter, gaps, form, chunk_len,
                                      seg_chunks, tol):
    from oracle import giekf
    giekf_form(form, chunk_len, seg_chunks)
    TOL = tol
    pb = make_problem(nsagp, D, N, T, k1, k2, seed=41 + D + T, kind="power", p=9, gaps=gaps, w_lik=1e-2)
    w, wf, cons, tune = _constrained(nsagp, pb, D, N)
    Eo, Vo, _, lbo, ubo, oo = giekf.gf_giekf_modulator_nmf_constraints(
        w, pb["t"], pb["y"], pb["ss_ref"], None, pb["t"], k1, k2, 1, D, N, g_iter, l_iter, cons, wf, tune, want_cov=True)
    Eg, Vg, Cg, lbg, ubg, og = nsagp.gf_giekf_modulator_nmf_constraints(
        w, pb["t"], pb["y"], pb["ss_gpu"], None, pb["t"], k1, k2, 1, D, N, g_iter, l_iter, cons, wf, tune, debug_cov=True)
    assert Cg is None
    assert rel_err(Eg, Eo) < TOL and rel_err(Vg, Vo) < TOL
    assert rel_err(lbg, lbo) < TOL and rel_err(ubg, ubo) < TOL
    assert rel_err(og["MF"], oo["MF"]) < TOL and rel_err(og["MS"], oo["MS"]) < TOL
    assert rel_err(og["PF"], oo["PF"]) < TOL and rel_err(og["PS"], oo["PS"]) < TOL
    if form == 2:               # the re-association costs rounding only
        assert rel_err(og["PS"], oo["PS"]) < 1e-9 and rel_err(og["MS"], oo["MS"]) < 1e-9
    if g_iter > 1:              # diagnostic over the marginal variances (include/nsagp.h)
        assert np.all(np.isfinite(og["maxDiffP"])) and og["maxDiffP"][0] > 0


def test_giekf_scan_single_step_signal(nsagp, gpu_lib, giekf_form):
    """T = 1 and T = 2: no / one smoothing element."""
    from oracle import giekf
    D, N, k1, k2 = 3, 2, "exp", "matern52"
    for T in (1, 2):
        pb = make_problem(nsagp, D, N, T, k1, k2, seed=5, kind="power", p=9, w_lik=1e-2)
        w, wf, cons, tune = _constrained(nsagp, pb, D, N)
        Eo, Vo, _, _, _, oo = giekf.gf_giekf_modulator_nmf_constraints(
            w, pb["t"], pb["y"], pb["ss_ref"], None, pb["t"], k1, k2, 1, D, N, 2, 1, cons, wf, tune, want_cov=True)
        giekf_form(2, 0, 0)
        Eg, Vg, _, _, _, og = nsagp.gf_giekf_modulator_nmf_constraints(
            w, pb["t"], pb["y"], pb["ss_gpu"], None, pb["t"], k1, k2, 1, D, N, 2, 1, cons, wf, tune, debug_cov=True)
        assert rel_err(Eg, Eo) < 1e-8 and rel_err(Vg, Vo) < 1e-8 and rel_err(og["PS"], oo["PS"]) < 1e-8


def test_giekf_energy_matches_oracle(nsagp, gpu_lib):
    from oracle import giekf
    D, N, T = 5, 2, 300
    pb = make_problem(nsagp, D, N, T, "matern32", "matern52", seed=12, kind="power", p=9, w_lik=1e-2)
    w, wf, cons, tune = _constrained(nsagp, pb, D, N)
    eo, go = giekf.gf_giekf_modulator_nmf_constraints(w, pb["t"], pb["y"], pb["ss_ref"], None, None, "matern32", "matern52",
                                                      1, D, N, 1, 1, cons, wf, tune)
    eg, gg = nsagp.gf_giekf_modulator_nmf_constraints(w, pb["t"], pb["y"], pb["ss_gpu"], None, None, "matern32", "matern52",
                                                      1, D, N, 1, 1, cons, wf, tune)
    assert abs(eg - eo) < TOL * abs(eo)
    assert np.all(gg == 0) and gg.shape == go.shape
    # a missing sample makes the reference's energy NaN (no isnan test in its nlZ branch)
    y2 = pb["y"].copy(); y2[10] = np.nan
    eg2, _ = nsagp.gf_giekf_modulator_nmf_constraints(w, pb["t"], y2, pb["ss_gpu"], None, None, "matern32", "matern52",
                                                      1, D, N, 1, 1, cons, wf, tune)
    assert np.isnan(eg2)


@pytest.mark.parametrize("form", [1, 2])
def test_giekf_plain_entry_carries_covariance(nsagp, gpu_lib, giekf_form, form):
    """gf_giekf_modulator_nmf (no constraints): log-scale w, (m, P) carried across global iterations."""
    from oracle import giekf
    D, N, T, k1, k2 = 5, 2, 140, "exp", "matern52"
    pb = make_problem(nsagp, D, N, T, k1, k2, seed=77, kind="power", p=9, gaps=True, w_lik=1e-2)
    w = pb["hyp"].pack_log()
    Eo, Vo, _, lbo, ubo, oo = giekf.gf_giekf_modulator_nmf(w, pb["t"], pb["y"], pb["ss_ref"], None, pb["t"], k1, k2, 1, D, N, 3, 1,
                                                           want_cov=True)
    giekf_form(form, 9, 4)
    Eg, Vg, _, lbg, ubg, og = nsagp.gf_giekf_modulator_nmf(w, pb["t"], pb["y"], pb["ss_gpu"], None, pb["t"], k1, k2, 1, D, N, 3, 1,
                                                           debug_cov=True)
    tol = TOL if form == 1 else 1e-6
    assert rel_err(Eg, Eo) < tol and rel_err(Vg, Vo) < tol and rel_err(lbg, lbo) < tol and rel_err(ubg, ubo) < tol
    assert rel_err(og["MS"], oo["MS"]) < tol and rel_err(og["PS"], oo["PS"]) < tol and rel_err(og["PF"], oo["PF"]) < tol
    # the carried covariance matters: with P reset every iteration the filtered covariance of step 1 would be the prior's
    oc = giekf.giekf_core(*_dense_model(pb, k1, k2, D, N), pb["y"], D, N, 3, 1, np.arange(T), want_cov=True)[4]
    assert rel_err(oo["PF"][:, :, 0], oc["PF"][:, :, 0]) > 1e-3


def _dense_model(pb, k1, k2, D, N):
    import math
    from oracle import ssmodel
    lik_param, param1, param2, Wnmf = ssmodel.unpack_log(pb["hyp"].pack_log(), 1, D, N)
    F, L, Qc, H, Pinf = pb["ss_ref"](pb["t"], param1, param2, k1, k2)[:5]
    F, L, H, Pinf, _ = ssmodel.balance_ss(F, L, H, Pinf)
    A, Q = ssmodel.lti_disc(F, L, Qc, 1.0)
    return A, Q, H, Pinf, math.exp(float(np.ravel(lik_param)[0])), Wnmf


@pytest.mark.parametrize("g_iter", [1, 3])
def test_giekf_c4_shape_n73_matches_oracle(nsagp, gpu_lib, giekf_form, g_iter):
    """BASELINE config C4's own shape (D = 32 exp subbands, N = 3 matern52 modulators: dense n = 73) with
    missing-data gaps, against the oracle (not against another GPU kernel), both smoother forms."""
    from oracle import giekf
    D, N, T, k1, k2 = 32, 3, 300, "exp", "matern52"
    pb = make_problem(nsagp, D, N, T, k1, k2, seed=404, kind="power", p=9, gaps=True, w_lik=1e-2, speech=True)
    w = pb["hyp"].pack_log()
    Eo, Vo, _, lbo, ubo, oo = giekf.gf_giekf_modulator_nmf(w, pb["t"], pb["y"], pb["ss_ref"], None, pb["t"], k1, k2, 1, D, N,
                                                           g_iter, 1, want_cov=True)
    assert oo["MS"].shape[0] == 73 and np.isnan(pb["y"]).sum() > 10
    for form, cl, sc, tol in ((1, 0, 0, TOL), (2, 0, 0, 1e-6), (2, 23, 5, 1e-6)):
        giekf_form(form, cl, sc)
        Eg, Vg, _, lbg, ubg, og = nsagp.gf_giekf_modulator_nmf(w, pb["t"], pb["y"], pb["ss_gpu"], None, pb["t"], k1, k2, 1, D, N,
                                                               g_iter, 1, debug_cov=True)
        assert rel_err(Eg, Eo) < tol and rel_err(Vg, Vo) < tol, form
        assert rel_err(lbg, lbo) < tol and rel_err(ubg, ubo) < tol
        assert rel_err(og["MF"], oo["MF"]) < tol and rel_err(og["MS"], oo["MS"]) < tol
        assert rel_err(og["PF"], oo["PF"]) < tol and rel_err(og["PS"], oo["PS"]) < tol


@pytest.mark.parametrize("g_iter", [1, 2])
def test_giekf_c4_shape_n137_matches_oracle(nsagp, gpu_lib, giekf_form, g_iter):
    """BASELINE config C4's second shape (D = 32 matern32 subbands, N = 3 matern52 modulators: dense n = 137,
    cf_matern32_to_ss.m:93-117) with missing-data gaps: the large-state smoother (csrc/ekfbig.cuh, operands in HBM / L2)
    against the oracle, default chunking and a chunking with several chunks and segments."""
    from oracle import giekf
    D, N, T, k1, k2 = 32, 3, 140, "matern32", "matern52"
    pb = make_problem(nsagp, D, N, T, k1, k2, seed=137, kind="power", p=9, gaps=True, w_lik=1e-2, speech=True)
    w = pb["hyp"].pack_log()
    Eo, Vo, _, lbo, ubo, oo = giekf.gf_giekf_modulator_nmf(w, pb["t"], pb["y"], pb["ss_ref"], None, pb["t"], k1, k2, 1, D, N,
                                                           g_iter, 1, want_cov=True)
    assert oo["MS"].shape[0] == 137
    for cl, sc in ((0, 0), (9, 4)):
        giekf_form(0 if cl == 0 else 2, cl, sc)
        Eg, Vg, _, lbg, ubg, og = nsagp.gf_giekf_modulator_nmf(w, pb["t"], pb["y"], pb["ss_gpu"], None, pb["t"], k1, k2, 1, D, N,
                                                               g_iter, 1, debug_cov=True)
        assert rel_err(og["MF"], oo["MF"]) < TOL and rel_err(og["PF"], oo["PF"]) < TOL          # the sequential filter
        assert rel_err(Eg, Eo) < 1e-6 and rel_err(Vg, Vo) < 1e-6
        assert rel_err(lbg, lbo) < 1e-6 and rel_err(ubg, ubo) < 1e-6
        assert rel_err(og["MS"], oo["MS"]) < 1e-6 and rel_err(og["PS"], oo["PS"]) < 1e-6
