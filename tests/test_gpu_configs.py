"""GPU parity for the BASELINE.json configurations that no other test compares with the oracle:

* C5 -- batched plans with B > 1 DIFFERENT problems (different hyper-parameters AND different clips), both plan
  kinds, through the plan API, through the one-call C entry points nsagp_ep_ihgp_batch / nsagp_ep_full_batch, and
  through batch.gpu_evaluator + finite_difference_gradient (what fminunc does around the nlZ mode,
  matlab/demo_toy_modulators_nmf.m:100-104), each against per-problem oracle calls;
* C3 -- gf_ep_modulator_nmf at the FULL length T = 500 000: the first filter pass is causal, so the first 1 000 steps
  of MF / ttau / tnu / lZ must equal the oracle's run on that prefix; plus the matern32-subband variant (n = 73);
* C2 -- the bench's exact EP schedule (ep_itts = 20, damping linspace(0.01, 0.1, 20)) on a prefix of the benched
  signal against the C restatement.

Tolerances (north_star): 1e-8 for the sequential pass, 1e-6 for what went through a re-associated scan.
"""
import ctypes as C

import numpy as np
import pytest

from conftest import make_problem, rel_err

pytestmark = pytest.mark.gpu
TOL_SEQ, TOL_SCAN = 1e-8, 1e-6
SCAN_TILE_DEFAULT = (16, 256)      # the library's default (csrc/api.cu)


def _args(pb, which, xt, alpha, damping, itts):
    return (pb["w"], pb["t"], pb["y"], pb["ss_" + which], pb["mom_" + which], xt, pb["kernel1"], pb["kernel2"],
            1, pb["D"], pb["N"], alpha, damping, itts)


# ------------------------------------------------------------------------------------------ C5: B > 1
def _batch_problems(nsagp, B, D, N, T, k1, k2, kind, p, shift):
    """B problems that differ in hyper-parameters AND in the clip (different seeds)."""
    return [make_problem(nsagp, D, N, T, k1, k2, seed=300 + 7 * b, kind=kind, p=p, shift=shift, speech=(b % 2 == 1))
            for b in range(B)]


def _models(nsagp, pbs, kind_id, k1, k2, D, N, smoother=False):
    from importlib import import_module
    entry = import_module(nsagp.__name__ + ".entry")
    L = nsagp._lib
    models, liks, tabs = [], [], []
    for pb in pbs:
        lik_param, p1, p2, W = entry._unpack_log(pb["w"], 1, D, N)
        mdl, _ = entry._discrete_model(pb["ss_gpu"], None, p1, p2, k1, k2, D, N, balance=(kind_id == L.KIND_IHGP),
                                       symmetrise_Q=(kind_id == L.KIND_IHGP))
        models.append(mdl); liks.append((pb["mom_gpu"], lik_param, W))
        if kind_id == L.KIND_IHGP:
            tabs.append(nsagp.tables.build_tables(mdl, want_smoother=smoother))
    return models, liks, (tabs if kind_id == L.KIND_IHGP else None)


@pytest.mark.parametrize("adf_form", [0, 1, 2, 3])
def test_batch_ihgp_nlz_matches_per_problem_oracle(nsagp, gpu_lib, adf_form):
    from oracle import ihgp_ep
    L = nsagp._lib
    B, D, N, T, k1, k2 = 5, 6, 3, 350, "exp", "matern52"
    pbs = _batch_problems(nsagp, B, D, N, T, k1, k2, "precalc", 9, 1.0)
    damping = [0.3]
    ref = [ihgp_ep.ihgp_ep_modulator_nmf(*_args(pb, "ref", None, 0.75, damping, 1))[0] for pb in pbs]
    models, liks, tabs = _models(nsagp, pbs, L.KIND_IHGP, k1, k2, D, N)
    y = np.stack([pb["y"] for pb in pbs])
    with nsagp.Plan(L.KIND_IHGP, models, liks, 0.75, damping, 1, y, L.MODE_NLZ, tables=tabs) as plan:
        plan.set_adf_form(adf_form)
        plan.run()
        got = [plan.fetch(b, ("edata",))["edata"] for b in range(B)]
        sites = [plan.fetch(b, ("ttau", "lZ")) for b in range(B)]
    assert len(set(np.round(ref, 6))) == B                      # the problems really differ
    for b in range(B):
        assert abs(got[b] - ref[b]) < TOL_SEQ * abs(ref[b]), (b, got[b], ref[b])
        assert abs(-np.sum(sites[b]["lZ"]) - ref[b]) < TOL_SEQ * abs(ref[b])
    # the same through the per-signal entry point: batch index b must pick up problem b's tables and likelihood
    for b in (0, B - 1):
        e1, _ = nsagp.ihgp_ep_modulator_nmf(*_args(pbs[b], "gpu", None, 0.75, damping, 1), adf_form=adf_form)
        assert abs(e1 - got[b]) <= 1e-12 * abs(e1)


def test_batch_ihgp_running_sites_matches_per_problem_oracle(nsagp, gpu_lib):
    """NSAGP_MODE_NLZ_RUNNING (ihgp_ep_modulator_nmf_constraints.m:568-615) with B > 1."""
    from oracle import ihgp_ep
    L = nsagp._lib
    B, D, N, T, k1, k2 = 4, 5, 2, 300, "matern32", "matern52"
    pbs = _batch_problems(nsagp, B, D, N, T, k1, k2, "power", 9, 0.0)
    damping = [0.5]
    cons = np.array([[0.0, 0.1], [50.0, 1000.0], [0.0, 3.2], [0.0, 20.0], [100.0, 3000.0], [0.0, 1.25]])
    tune = [1, 1, 1, 1, 1, 1, 1]
    ref = []
    for pb in pbs:
        hyp = pb["hyp"]
        parts = [hyp.var_fast, np.clip(hyp.len_fast, 51, 999), hyp.omega, hyp.var_slow, np.clip(hyp.len_slow, 101, 2999),
                 hyp.W.reshape(-1, order="F")]
        if not all(np.all((v > c[0]) & (v < c[1])) for v, c in zip(parts, cons)):
            pytest.skip("synthetic hyper-parameters outside the constraint boxes")
        wc = np.concatenate([np.log([hyp.w_lik])] + [nsagp.inv_sigmoid(v, c) for v, c in zip(parts, cons)])
        a = list(_args(pb, "ref", None, 0.5, damping, 1)); a[0] = wc
        ref.append(ihgp_ep.ihgp_ep_modulator_nmf_constraints(*a, cons, np.zeros(0), tune)[0])
        # the product problem must see the same (clipped) parameters
        pb["w"] = np.concatenate([np.log([hyp.w_lik]), np.log(np.concatenate(parts))])
    models, liks, tabs = _models(nsagp, pbs, L.KIND_IHGP, k1, k2, D, N)
    y = np.stack([pb["y"] for pb in pbs])
    with nsagp.Plan(L.KIND_IHGP, models, liks, 0.5, damping, 1, y, L.MODE_NLZ_RUNNING, tables=tabs) as plan:
        plan.run()
        got = [plan.fetch(b, ("edata",))["edata"] for b in range(B)]
    for b in range(B):
        assert abs(got[b] - ref[b]) < TOL_SEQ * abs(ref[b]), (b, got[b], ref[b])


@pytest.mark.parametrize("itts", [1, 3])
def test_batch_full_nlz_matches_per_problem_oracle(nsagp, gpu_lib, itts):
    from oracle import gf_ep
    L = nsagp._lib
    B, D, N, T, k1, k2 = 5, 5, 2, 300, "matern32", "matern52"
    pbs = _batch_problems(nsagp, B, D, N, T, k1, k2, "power", 9, 0.0)
    damping = np.linspace(0.5, 0.4, itts)
    ref = [gf_ep.gf_ep_modulator_nmf(*_args(pb, "ref", None, 0.5, damping, itts))[0] for pb in pbs]
    models, liks, _ = _models(nsagp, pbs, L.KIND_FULL, k1, k2, D, N)
    y = np.stack([pb["y"] for pb in pbs])
    tol = TOL_SEQ if itts == 1 else TOL_SCAN
    with nsagp.Plan(L.KIND_FULL, models, liks, 0.5, damping, itts, y, L.MODE_NLZ) as plan:
        plan.run()
        got = [plan.fetch(b, ("edata",))["edata"] for b in range(B)]
    for b in range(B):
        assert abs(got[b] - ref[b]) < tol * abs(ref[b]), (b, got[b], ref[b])


@pytest.mark.parametrize("kind_name", ["ihgp", "full"])
def test_batch_one_call_c_entry_points(nsagp, gpu_lib, kind_name):
    """nsagp_ep_ihgp_batch / nsagp_ep_full_batch with HOST buffers: B = 3 problems in predict mode, every output
    of every problem against the per-problem oracle."""
    from oracle import gf_ep, ihgp_ep
    lm = nsagp._lib
    L = lm.lib()
    B, D, N, T, k1, k2, itts = 3, 4, 2, 260, "matern32", "matern52", 3
    pbs = _batch_problems(nsagp, B, D, N, T, k1, k2, "power", 9, 0.0)
    damping = lm.as_f64(np.linspace(0.5, 0.3, itts))
    kind_id = lm.KIND_IHGP if kind_name == "ihgp" else lm.KIND_FULL
    models, liks, tabs = _models(nsagp, pbs, kind_id, k1, k2, D, N, smoother=True)
    keep = []
    cm, cl = (lm.Model * B)(), (lm.Lik * B)()
    ct = (lm.Tables * B)() if tabs else None
    outs = (lm.Outputs * B)()
    bufs = []
    M = D + N
    for b in range(B):
        mdl = models[b]
        arrs = [lm.as_f64(a) for a in (mdl.A, mdl.Q, mdl.Pinf, mdl.h)]
        keep.extend(arrs)
        cm[b].D, cm[b].N, cm[b].bz, cm[b].bg = mdl.D, mdl.N, mdl.bz, mdl.bg
        cm[b].A, cm[b].Q, cm[b].Pinf, cm[b].h = [lm.dptr(a) for a in arrs]
        cl[b] = liks[b][0].c_lik(liks[b][1], liks[b][2], keep)
        if ct is not None:
            pp, pg = tabs[b].packed()
            r, pp, pg = lm.as_f64(tabs[b].r), lm.as_f64(pp), lm.as_f64(pg)
            keep.extend([r, pp, pg])
            ct[b].nr, ct[b].r, ct[b].PP, ct[b].PG = r.size, lm.dptr(r), lm.dptr(pp), lm.dptr(pg)
        o = dict(Eft=np.empty((T, M)), Varft=np.empty((T, M)), ttau=np.empty((T, M)), nlZ=np.empty(itts), MS=np.empty((T, mdl.n)))
        for k, a in o.items():
            setattr(outs[b], k, lm.dptr(a))
        bufs.append(o)
    y = lm.as_f64(np.stack([pb["y"] for pb in pbs]))
    ep = lm.Ep(0.5, lm.dptr(damping), itts)
    if kind_id == lm.KIND_IHGP:
        lm.check(L.nsagp_ep_ihgp_batch(B, cm, cl, C.byref(ep), ct, lm.dptr(y), T, lm.MODE_PREDICT, outs))
    else:
        lm.check(L.nsagp_ep_full_batch(B, cm, cl, C.byref(ep), lm.dptr(y), T, lm.MODE_PREDICT, outs))
    for b, pb in enumerate(pbs):
        fn = ihgp_ep.ihgp_ep_modulator_nmf if kind_id == lm.KIND_IHGP else gf_ep.gf_ep_modulator_nmf
        Eo, Vo, _, _, _, oo = fn(*_args(pb, "ref", pb["t"], 0.5, damping, itts))
        assert rel_err(bufs[b]["Eft"].T, Eo) < TOL_SCAN and rel_err(bufs[b]["Varft"].T, Vo) < TOL_SCAN
        assert rel_err(bufs[b]["ttau"].T, oo["ttau"]) < TOL_SCAN and rel_err(bufs[b]["MS"].T, oo["MS"]) < TOL_SCAN
        assert rel_err(bufs[b]["nlZ"], oo["nlZ"]) < TOL_SCAN


@pytest.mark.parametrize("kind_name", ["ihgp", "full"])
def test_gpu_evaluator_and_fd_gradient_match_oracle(nsagp, gpu_lib, kind_name):
    """batch.gpu_evaluator + finite_difference_gradient: 1 + numel(w) evaluations as one batched plan; every
    value against the oracle's nlZ at the same perturbed parameter vector."""
    from importlib import import_module
    from oracle import gf_ep, ihgp_ep
    batch = import_module(nsagp.__name__ + ".batch")
    L = nsagp._lib
    D, N, T, k1, k2 = 3, 2, 250, "exp", "matern52"
    pb = make_problem(nsagp, D, N, T, k1, k2, seed=17, kind="power", p=7)
    itts = 1 if kind_name == "ihgp" else 2
    damping = [0.5, 0.5][:max(itts, 1)]
    kind_id = L.KIND_IHGP if kind_name == "ihgp" else L.KIND_FULL
    ev = batch.gpu_evaluator(kind_id, pb["ss_gpu"], pb["mom_gpu"], k1, k2, 1, D, N, 0.5, damping, itts)
    h = 1e-5
    f0, g = batch.finite_difference_gradient(pb["w"], pb["y"], ev, h=h)
    fn = ihgp_ep.ihgp_ep_modulator_nmf if kind_name == "ihgp" else gf_ep.gf_ep_modulator_nmf
    ref = lambda w: fn(w, pb["t"], pb["y"], pb["ss_ref"], pb["mom_ref"], None, k1, k2, 1, D, N, 0.5, damping, itts)[0]
    tol = TOL_SEQ if itts == 1 else TOL_SCAN
    r0 = ref(pb["w"])
    assert abs(f0 - r0) < tol * abs(r0)
    assert g.shape == pb["w"].shape
    for i in range(pb["w"].size):
        wi = pb["w"].copy(); wi[i] += h
        ri = ref(wi)
        gi = f0 + g[i] * h                                          # the evaluator's value at the perturbed point
        assert abs(gi - ri) < 10 * tol * abs(ri), (i, gi, ri)
    assert np.any(np.abs(g) > 1e-3)                                 # the objective does depend on the parameters


# ------------------------------------------------------------------------------------------ C3
def test_c3_full_length_prefix_against_oracle(nsagp, gpu_lib):
    """gf_ep_modulator_nmf at BASELINE's C3 size (D=16 exp subbands, N=3 matern52 modulators, n=41, T=500 000): the
    first filter pass is causal, so the first Tp steps of its outputs are what the oracle computes on the prefix."""
    from oracle import gf_ep, cubature as ocub, lik as olik, ssmodel as oss
    L = nsagp._lib
    D, N, T, Tp, k1, k2 = 16, 3, 500000, 1000, "exp", "matern52"
    rng = np.random.default_rng(2027)
    hyp = nsagp.synth.speech_hypers(D, N, rng)
    y, _, _ = nsagp.synth.sample_signal(hyp, k1, k2, T, rng, link_shift=1.0, sqrt_model=True)
    F, Lm, Qc, H, Pinf = nsagp.ss_modulators_nmf(hyp.w_sub(), hyp.w_mod(), k1, k2)[:5]
    A, Q = nsagp.lti_disc(F, Lm, Qc, 1.0)
    mdl = nsagp.to_block_model(A, Q, H, Pinf, D, N)
    wn, xn = nsagp.utp_ws(9, N)
    mom = nsagp.likModulatorPreCalcwn(nsagp.Softplus(1.0), wn, xn)
    damp = [0.05]
    with nsagp.Plan(L.KIND_FULL, [mdl], [(mom, np.log([hyp.w_lik]), hyp.W)], 0.75, damp, 1, y[None, :], L.MODE_PREDICT) as plan:
        plan.run()
        got = plan.fetch(0, ("MF", "ttau", "tnu", "lZ", "Eft", "Varft", "nlZ"))
    assert np.all(np.isfinite(got["Eft"])) and np.all(got["Varft"] > 0) and np.all(got["ttau"] >= 0)
    wo, xo = ocub.utp_ws(9, N)
    mom_ref = olik.make_mom("precalc", olik.softplus_link(1.0), wn=wo, xn_unscaled=xo)
    ss = lambda x, p1, p2, a, b: oss.ss_modulators_nmf(p1, p2, a, b)
    t = np.arange(1.0, Tp + 1.0)
    _, _, _, _, _, oo = gf_ep.gf_ep_modulator_nmf(hyp.pack_log(), t, y[:Tp], ss, mom_ref, t, k1, k2, 1, D, N, 0.75, damp, 1)
    for k in ("MF", "ttau", "tnu"):
        assert rel_err(got[k][:, :Tp], oo[k]) < TOL_SEQ, k
    assert rel_err(got["lZ"][:Tp], np.ravel(oo["lZ"])) < TOL_SEQ


@pytest.mark.parametrize("itts", [1, 3])
def test_c3_matern32_variant_n73_against_oracle(nsagp, gpu_lib, itts):
    """C3's second shape: matern32 subbands (4x4 blocks), n = 16*4 + 3*3 = 73."""
    from oracle import gf_ep
    D, N, T, k1, k2 = 16, 3, 300, "matern32", "matern52"
    pb = make_problem(nsagp, D, N, T, k1, k2, seed=73, kind="precalc", p=9, shift=1.0, speech=True)
    damping = np.linspace(0.3, 0.2, itts)
    Eo, Vo, _, _, _, oo = gf_ep.gf_ep_modulator_nmf(*_args(pb, "ref", pb["t"], 0.75, damping, itts))
    Eg, Vg, _, _, _, og = nsagp.gf_ep_modulator_nmf(*_args(pb, "gpu", pb["t"], 0.75, damping, itts))
    tol = TOL_SEQ if itts == 1 else TOL_SCAN
    assert oo["MS"].shape[0] == 73
    assert rel_err(og["nlZ"], oo["nlZ"]) < tol and rel_err(og["ttau"], oo["ttau"]) < tol
    assert rel_err(og["MF"], oo["MF"]) < tol
    assert rel_err(Eg, Eo) < TOL_SCAN and rel_err(Vg, Vo) < TOL_SCAN and rel_err(og["MS"], oo["MS"]) < TOL_SCAN


# ------------------------------------------------------------------------------------------ C2, bench schedule
def test_c2_bench_schedule_prefix_against_oracle(nsagp, gpu_lib):
    """The bench's exact EP schedule (ep_itts = 20, damping = linspace(0.01, 0.1, 20), alpha = 0.75) on the first
    2 000 samples of the benched signal (seed 2026) against the C restatement of the reference."""
    import bench
    r = bench.prefix_check(nsagp, seed=2026, Tp=2000, itts=20)
    assert r["lZ_rel_err"] < TOL_SCAN and r["Eft_rel_err"] < TOL_SCAN and r["ttau_rel_err"] < TOL_SCAN
    assert r["lZ_rel_err_first_sweep"] < TOL_SEQ


# ------------------------------------------------------------------------------------------ family-specialised scans
@pytest.fixture(params=[1, 0], ids=["one-tile", "launch-per-family"])
def family_scans(nsagp, request):
    """The family-specialised form of the frozen-site scans both as one CTA tile for the two families (default) and as
    one launch per family."""
    L = nsagp._lib
    L.check(L.lib().nsagp_scan_config(0))
    L.check(L.lib().nsagp_scan_merge(request.param))
    yield
    L.check(L.lib().nsagp_scan_merge(1))


@pytest.mark.parametrize("D,N,k1,k2", [(16, 3, "exp", "matern52"),        # (BM, bz, bg) = (3, 2, 3): C2 / C3
                                        (5, 2, "matern32", "matern52"),    # (4, 4, 3): the demo's kernels
                                        (4, 2, "matern52", "matern52"),    # (6, 6, 3)
                                        (4, 2, "exp", "exp")])             # no specialised pair: padded path
@pytest.mark.parametrize("entry_name", ["ihgp", "gfep"])
def test_family_specialised_scans_match_oracle(nsagp, gpu_lib, family_scans, entry_name, D, N, k1, k2):
    """Subband and modulator latents as separate launches, each at its own block size (csrc/scan.cuh), against the
    oracle: predict mode, 3 EP iterations, gaps."""
    from oracle import gf_ep, ihgp_ep
    T = 330
    pb = make_problem(nsagp, D, N, T, k1, k2, seed=90 + D, kind="precalc", p=9, shift=1.0, gaps=True)
    damping = np.linspace(0.4, 0.2, 3)
    if entry_name == "ihgp":
        ref, got = ihgp_ep.ihgp_ep_modulator_nmf, nsagp.ihgp_ep_modulator_nmf
    else:
        ref, got = gf_ep.gf_ep_modulator_nmf, nsagp.gf_ep_modulator_nmf
    Eo, Vo, _, _, _, oo = ref(*_args(pb, "ref", pb["t"], 0.75, damping, 3))
    Eg, Vg, _, _, _, og = got(*_args(pb, "gpu", pb["t"], 0.75, damping, 3))
    assert rel_err(Eg, Eo) < TOL_SCAN and rel_err(Vg, Vo) < TOL_SCAN
    assert rel_err(og["nlZ"], oo["nlZ"]) < TOL_SCAN and rel_err(og["MS"], oo["MS"]) < TOL_SCAN
    assert rel_err(og["ttau"], oo["ttau"]) < TOL_SCAN and rel_err(og["MF"], oo["MF"]) < TOL_SCAN
    assert og["n_negcav"] == oo["n_negcav"]


# ------------------------------------------------------------------------------------------ site-update forms
@pytest.mark.parametrize("entry_name", ["ihgp", "gfep"])
@pytest.mark.parametrize("D,N,p", [(16, 3, 9), (6, 2, 9), (20, 4, 7), (5, 2, 3)])
def test_site_update_forms_agree(nsagp, gpu_lib, entry_name, D, N, p):
    """The smoother-side site update in its three forms (csrc/siteupd.cuh: sigma points two at a time / one at a time;
    csrc/ihgp.cuh: one thread per step) against the oracle and against each other: D > 16 (eight subbands per lane),
    N = 2 .. 4 (every split of the modulators' g-sums over the four lanes), S = 5 .. 97 (always odd: the last round has
    one point), gaps."""
    from oracle import gf_ep, ihgp_ep
    T = 200
    pb = make_problem(nsagp, D, N, T, "exp", "matern52", seed=300 + D, kind="precalc", p=p, shift=1.0, gaps=True)
    damping = np.linspace(0.4, 0.2, 3)
    if entry_name == "ihgp":
        ref, got = ihgp_ep.ihgp_ep_modulator_nmf, nsagp.ihgp_ep_modulator_nmf
    else:
        ref, got = gf_ep.gf_ep_modulator_nmf, nsagp.gf_ep_modulator_nmf
    Eo, Vo, _, _, _, oo = ref(*_args(pb, "ref", pb["t"], 0.75, damping, 3))
    L = nsagp._lib
    res = {}
    try:
        for form in (0, 2, 1):
            L.check(L.lib().nsagp_site_config(form))
            Eg, Vg, _, _, _, og = got(*_args(pb, "gpu", pb["t"], 0.75, damping, 3))
            res[form] = (Eg, og["ttau"], og["tnu"], og["nlZ"])
            assert rel_err(Eg, Eo) < TOL_SCAN and rel_err(Vg, Vo) < TOL_SCAN, form
            assert rel_err(og["ttau"], oo["ttau"]) < TOL_SCAN and rel_err(og["nlZ"], oo["nlZ"]) < TOL_SCAN, form
            assert og["n_negcav"] == oo["n_negcav"], form
    finally:
        L.check(L.lib().nsagp_site_config(0))
    for form in (2, 1):
        for a, b in zip(res[0], res[form]):
            assert rel_err(a, b) < 1e-9, form


# ------------------------------------------------------------------------------------------ scan tile geometry / element forms
@pytest.mark.parametrize("D,N,k1,k2", [(16, 3, "exp", "matern52"),        # (3, 2, 3): C2
                                        (5, 2, "matern32", "matern52"),    # (4, 4, 3)
                                        (4, 2, "exp", "exp")])             # no specialised pair: padded path
@pytest.mark.parametrize("threads,chunks", [(256, 16), (320, 16), (256, 5), (320, 3)])
def test_scan_tile_geometry_matches_oracle(nsagp, gpu_lib, threads, chunks, D, N, k1, k2):
    """The frozen-site scans of the infinite-horizon path with every tile geometry (nsagp_scan_tile: 256-thread tiles at
    128 registers / 320-thread tiles at 96, full and short tiles) against the oracle."""
    from oracle import ihgp_ep
    T = 1500 if chunks == 16 else 700                # several tiles, a ragged last tile
    pb = make_problem(nsagp, D, N, T, k1, k2, seed=120 + D, kind="precalc", p=9, shift=1.0, gaps=True)
    damping = np.linspace(0.4, 0.2, 3)
    Eo, Vo, _, _, _, oo = ihgp_ep.ihgp_ep_modulator_nmf(*_args(pb, "ref", pb["t"], 0.75, damping, 3))
    L = nsagp._lib
    try:
        L.check(L.lib().nsagp_scan_tile(chunks, threads))
        Eg, Vg, _, _, _, og = nsagp.ihgp_ep_modulator_nmf(*_args(pb, "gpu", pb["t"], 0.75, damping, 3))
    finally:
        L.check(L.lib().nsagp_scan_tile(*SCAN_TILE_DEFAULT))
    assert rel_err(Eg, Eo) < TOL_SCAN and rel_err(Vg, Vo) < TOL_SCAN
    assert rel_err(og["nlZ"], oo["nlZ"]) < TOL_SCAN and rel_err(og["MS"], oo["MS"]) < TOL_SCAN
    assert rel_err(og["ttau"], oo["ttau"]) < TOL_SCAN and rel_err(og["MF"], oo["MF"]) < TOL_SCAN
    assert og["n_negcav"] == oo["n_negcav"]
