"""GPU edge cases through the entry points and the C ABI: shortest signals, signals whose length
straddles the scan chunk / tile sizes, all-missing and leading/trailing-missing data, unsorted and
duplicated inputs with separate test points, the largest supported shapes, argument errors."""
import ctypes as C

import numpy as np
import pytest

from conftest import make_problem, rel_err

pytestmark = pytest.mark.gpu


def _args(pb, which, xt, alpha, damping, itts, y=None, t=None):
    return (pb["w"], pb["t"] if t is None else t, pb["y"] if y is None else y, pb["ss_" + which], pb["mom_" + which], xt,
            pb["kernel1"], pb["kernel2"], 1, pb["D"], pb["N"], alpha, damping, itts)


@pytest.mark.parametrize("T", [1, 2, 3, 31, 32, 33, 511, 512, 513, 1025])
@pytest.mark.parametrize("entry", ["ihgp", "gfep"])
def test_lengths_around_chunk_boundaries(nsagp, gpu_lib, entry, T):
    """scan chunks are 32 steps, CTA tiles 256-512 steps: lengths on both sides of every boundary.
    (9th-order cubature: with the 5th-order rule on these tiny models some EP site updates divide by
    1 + d2lZ*v_cav ~ 1e-9, and then oracle and GPU -- and both GPU forms -- legitimately differ by 1e-5.)"""
    from oracle import gf_ep, ihgp_ep
    pb = make_problem(nsagp, 3, 2, T, "matern32", "matern52", seed=50 + T, kind="power", p=9)
    damping = np.array([0.5, 0.4, 0.3])
    ref = (ihgp_ep.ihgp_ep_modulator_nmf if entry == "ihgp" else gf_ep.gf_ep_modulator_nmf)
    gpu = (nsagp.ihgp_ep_modulator_nmf if entry == "ihgp" else nsagp.gf_ep_modulator_nmf)
    Eo, Vo, _, _, _, oo = ref(*_args(pb, "ref", pb["t"], 0.5, damping, 3))
    Eg, Vg, _, _, _, og = gpu(*_args(pb, "gpu", pb["t"], 0.5, damping, 3))
    assert rel_err(Eg, Eo) < 1e-6 and rel_err(Vg, Vo) < 1e-6
    assert rel_err(og["nlZ"], oo["nlZ"]) < 1e-6
    assert rel_err(og["ttau"], oo["ttau"]) < 1e-6 and rel_err(og["MS"], oo["MS"]) < 1e-6
    eo, _ = ref(*_args(pb, "ref", None, 0.5, damping, 3))
    eg, _ = gpu(*_args(pb, "gpu", None, 0.5, damping, 3))
    assert abs(eg - eo) <= 1e-6 * abs(eo)


@pytest.mark.parametrize("entry", ["ihgp", "gfep"])
@pytest.mark.parametrize("pattern", ["all", "head", "tail", "alternate"])
def test_missing_data_patterns(nsagp, gpu_lib, entry, pattern):
    from oracle import gf_ep, ihgp_ep
    T = 120
    pb = make_problem(nsagp, 4, 2, T, "matern32", "matern52", seed=77, kind="power", p=7)
    y = pb["y"].copy()
    if pattern == "all":
        y[:] = np.nan
    elif pattern == "head":
        y[:40] = np.nan
    elif pattern == "tail":
        y[-40:] = np.nan
    else:
        y[::2] = np.nan
    damping = np.array([0.5, 0.5])
    ref = (ihgp_ep.ihgp_ep_modulator_nmf if entry == "ihgp" else gf_ep.gf_ep_modulator_nmf)
    gpu = (nsagp.ihgp_ep_modulator_nmf if entry == "ihgp" else nsagp.gf_ep_modulator_nmf)
    Eo, Vo, _, _, _, oo = ref(*_args(pb, "ref", pb["t"], 0.5, damping, 2, y=y))
    Eg, Vg, _, _, _, og = gpu(*_args(pb, "gpu", pb["t"], 0.5, damping, 2, y=y))
    assert rel_err(Eg, Eo) < 1e-6 and rel_err(Vg, Vo) < 1e-6
    assert rel_err(og["ttau"], oo["ttau"]) < 1e-6 and rel_err(og["tnu"], oo["tnu"]) < 1e-6
    assert rel_err(og["R"], oo["R"]) < 1e-6
    assert rel_err(og["nlZ"], oo["nlZ"]) < 1e-6


def test_unsorted_duplicated_inputs_and_test_points(nsagp, gpu_lib):
    """x unsorted with a duplicate, xt partly new points: unique(...,'first') + NaN for test-only points
    (gf_ep_modulator_nmf.m:58-66); outputs are returned at the test points only, in xt's order."""
    from oracle import ihgp_ep
    T = 90
    pb = make_problem(nsagp, 3, 2, T, "matern32", "matern52", seed=3, kind="power", p=5)
    rng = np.random.default_rng(0)
    perm = rng.permutation(T)
    x = np.concatenate([pb["t"][perm], pb["t"][perm[:1]]])             # one duplicated input
    y = np.concatenate([pb["y"][perm], [123.0]])                        # its second value must be ignored
    xt = np.array([T + 2.0, 5.0, T + 1.0, 17.0])
    damping = np.array([0.5, 0.5])
    Eo, Vo, _, lbo, ubo, _ = ihgp_ep.ihgp_ep_modulator_nmf(*_args(pb, "ref", xt, 0.5, damping, 2, y=y, t=x))
    Eg, Vg, _, lbg, ubg, _ = nsagp.ihgp_ep_modulator_nmf(*_args(pb, "gpu", xt, 0.5, damping, 2, y=y, t=x))
    assert Eg.shape == (5, 4)
    assert rel_err(Eg, Eo) < 1e-6 and rel_err(Vg, Vo) < 1e-6 and rel_err(lbg, lbo) < 1e-6 and rel_err(ubg, ubo) < 1e-6


def test_largest_shapes(nsagp, gpu_lib):
    """D + N = 32 sites (one warp lane per latent), N = 4 modulators, 8x8 subband blocks (matern72),
    Gauss-Hermite 4^4 = 256 sigma points (several rounds per moment thread)."""
    from oracle import gf_ep, lik as olik
    D, N, T = 28, 4, 40
    pb = make_problem(nsagp, D, N, T, "matern72", "matern32", seed=9, kind="power", p=5)
    mom_gpu = nsagp.likModulatorNMFPower(nsagp.Softplus(0.0), 4, N)     # p not in {3,5,7,9}: tensor Gauss-Hermite
    mom_ref = olik.make_mom("power", olik.softplus_link(0.0), p=4)
    damping = np.array([0.5, 0.5])
    a = list(_args(pb, "ref", pb["t"], 0.5, damping, 2)); a[4] = mom_ref
    Eo, Vo, _, _, _, oo = gf_ep.gf_ep_modulator_nmf(*a)
    a = list(_args(pb, "gpu", pb["t"], 0.5, damping, 2)); a[4] = mom_gpu
    for form in (0, 1):
        Eg, Vg, _, _, _, og = nsagp.gf_ep_modulator_nmf(*a, adf_form=form)
        assert rel_err(Eg, Eo) < 1e-6 and rel_err(Vg, Vo) < 1e-6
        assert rel_err(og["nlZ"], oo["nlZ"]) < 1e-6


def test_argument_errors(nsagp, gpu_lib):
    L = nsagp._lib
    pb = make_problem(nsagp, 3, 2, 20, "matern32", "matern52", seed=1, kind="power", p=5)
    with pytest.raises(ValueError):                                     # ep_damping shorter than ep_itts (indexed by itt+1)
        nsagp.gf_ep_modulator_nmf(*_args(pb, "gpu", pb["t"], 0.5, [0.5], 3))
    with pytest.raises(TypeError):                                      # an arbitrary handle cannot run on the GPU
        a = list(_args(pb, "gpu", pb["t"], 0.5, [0.5], 1)); a[4] = lambda *x: None
        nsagp.gf_ep_modulator_nmf(*a)
    with pytest.raises(ValueError):                                     # wrong parameter-vector length
        a = list(_args(pb, "gpu", pb["t"], 0.5, [0.5], 1)); a[0] = pb["w"][:-1]
        nsagp.gf_ep_modulator_nmf(*a)
    assert gpu_lib.nsagp_ep_full(None, None, None, None, 0, 0, None) == -1          # NSAGP_ERR_INVALID, no crash
    assert b"null" in gpu_lib.nsagp_last_error()
    o = L.Outputs()
    assert gpu_lib.nsagp_plan_fetch(None, 0, C.byref(o)) == -1
