"""GPU sweep over model shapes (no oracle: it would take minutes per case): every combination must launch, stay
finite and give the same answer through the two independent implementations of the sequential pass (adf_form 0:
one CTA per signal, fused update, straight-line math; adf_form 1: one warp per signal, library math, the reference's
literal update order).  Guards the launch geometry (tile sizes, shared memory, registers) of shapes the parity cases
do not cover -- D + N up to the 32-site limit, every kernel pairing, 1..4 modulators."""
import itertools

import numpy as np
import pytest

from conftest import make_problem, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def native_tables(nsagp, gpu_lib):
    """No oracle here, so the library's own Riccati tables are used: SciPy's generic solver is ill-conditioned for the
    8 x 8 matern72 blocks (rcond 2e-16) and its noisy tables flip nearest-neighbour look-ups, which made the two forms of
    the sequential pass differ by 3 % on (matern72, matern72); with the native tables they agree to 3e-12."""
    nsagp.tables.DEFAULT_NATIVE = True
    yield
    nsagp.tables.DEFAULT_NATIVE = False


KERNELS = ["exp", "matern32", "matern52", "matern72"]
SHAPES = [(1, 2), (2, 2), (3, 4), (8, 2), (12, 3), (16, 3), (16, 4), (24, 4), (28, 4), (30, 2)]


def _cases():
    out = []
    for i, (D, N) in enumerate(SHAPES):
        k1, k2 = KERNELS[i % 4], KERNELS[(i // 2 + 1) % 4]
        out.append((D, N, k1, k2))
    for k1, k2 in itertools.product(KERNELS, KERNELS):          # every kernel pairing at one mid-size shape
        out.append((6, 2, k1, k2))
    return out


@pytest.mark.parametrize("D,N,k1,k2", _cases())
def test_every_shape_launches_and_both_forms_agree(nsagp, gpu_lib, D, N, k1, k2):
    T = 150
    pb = make_problem(nsagp, D, N, T, k1, k2, seed=1000 + 31 * D + N, kind="power", p=5, gaps=(D % 3 == 0))
    damping = [0.5, 0.4, 0.3]
    args = (pb["w"], pb["t"], pb["y"], pb["ss_gpu"], pb["mom_gpu"], pb["t"], k1, k2, 1, D, N, 0.5, damping, 3)
    for entry in (nsagp.ihgp_ep_modulator_nmf, nsagp.gf_ep_modulator_nmf):
        res = [entry(*args, adf_form=f) for f in (0, 1)]
        (E0, V0, _, lb0, ub0, o0), (E1, V1, _, lb1, ub1, o1) = res
        assert np.all(np.isfinite(E0)) and np.all(np.isfinite(V0)) and np.all(V0 > 0), entry.__name__
        # (matern72, matern72) through the infinite-horizon path: 8 x 8 blocks with rcond ~1e-16 and sites of 1e7; a
        # 1e-13 difference between the two sequential passes can flip ONE nearest-neighbour table look-up
        # (ihgp_ep_modulator_nmf.m:239), which moves the posterior mean locally by a few per cent while nlZ agrees to
        # 1e-13 -- the reference's own discontinuity, not a kernel property.  Everything else must agree to 1e-6.
        fragile = entry is nsagp.ihgp_ep_modulator_nmf and k1 == k2 == "matern72"
        assert rel_err(E1, E0) < (0.1 if fragile else 1e-6) and rel_err(V1, V0) < 1e-6, entry.__name__
        assert rel_err(o1["nlZ"], o0["nlZ"]) < 1e-6, entry.__name__
    # nlZ mode of both families
    for entry in (nsagp.ihgp_ep_modulator_nmf, nsagp.gf_ep_modulator_nmf):
        a = entry(*(args[:5] + (None,) + args[6:]), adf_form=0)
        b = entry(*(args[:5] + (None,) + args[6:]), adf_form=1)
        assert np.isfinite(a[0]) and abs(a[0] - b[0]) < 1e-6 * abs(b[0]), entry.__name__


@pytest.mark.parametrize("D,N,k1,k2", [(2, 2, "exp", "exp"), (8, 2, "matern32", "matern32"), (24, 4, "exp", "matern52"),
                                        (16, 3, "matern32", "matern72"), (30, 3, "exp", "matern32")])
def test_ekf_shapes_scan_matches_first_generation_kernels(nsagp, gpu_lib, D, N, k1, k2):
    L = nsagp._lib
    T = 130
    pb = make_problem(nsagp, D, N, T, k1, k2, seed=77 + D, kind="power", p=3, gaps=True, w_lik=1e-2)
    order = {"exp": 1, "matern32": 2, "matern52": 3, "matern72": 4}
    n = 2 * order[k1] * D + order[k2] * N
    outs = []
    for form in (1, 2):
        if form == 1 and n > 75:                # the first-generation smoother needs five n x n matrices in shared memory
            continue
        L.check(L.lib().nsagp_giekf_config(form, 11 if form == 2 else 0, 3 if form == 2 else 0))
        try:
            outs.append(nsagp.gf_giekf_modulator_nmf(pb["w"], pb["t"], pb["y"], pb["ss_gpu"], None, pb["t"], k1, k2, 1, D, N, 2, 1))
        finally:
            L.check(L.lib().nsagp_giekf_config(0, 0, 0))
    for E, V, *_ in outs:
        assert np.all(np.isfinite(E)) and np.all(V > 0)
    if len(outs) == 2:
        assert rel_err(outs[1][0], outs[0][0]) < 1e-8 and rel_err(outs[1][1], outs[0][1]) < 1e-8
