"""GPU parity for the analytic-gradient mode of gf_giekf_modulator_nmf (matlab/gf_giekf_modulator_nmf.m:296-437,
GradObj = 'on') -- csrc/ekfgrad.cuh through the C ABI nsagp_giekf_grad:

* against the oracle's dense restatement (oracle/giekf.py giekf_energy_grad) on the same seeded inputs, both as the
  reference computes it (dF / dPinf left unbalanced) and with the reference's commented-out balancing loop;
* against central finite differences of the GPU's own energy (possible only for the balanced variant, which is the
  derivative of the energy) on a signal longer than the oracle handles;
* the NaN contract and the argument checks.

Tolerance: the recursion is sequential FP64 in both implementations; they differ by the operation order of the
dense products (the device works block by block): 1e-8 relative, ENTRY BY ENTRY (the entries span 15 decades in the
unbalanced variant, a max-norm would only test the largest; observed <= 6e-12).
"""
import numpy as np
import pytest

from conftest import make_problem

pytestmark = pytest.mark.gpu
TOL = 1e-8


def _ss_ref_with_derivs():
    from oracle import ssmodel as oss
    return lambda x, p1, p2, k1, k2: oss.ss_modulators_nmf(p1, p2, k1, k2) + oss.ss_modulators_nmf_derivs(p1, p2, k1, k2)


CASES = [
    # D, N, T, k1, k2
    (4, 2, 120, "matern32", "matern52"),
    (6, 3, 150, "exp", "matern52"),            # the kernels of C2 / C4
    (3, 2, 80, "matern72", "exp"),             # 8 x 8 and 1 x 1 blocks
    (5, 2, 100, "matern52", "matern32"),       # 6 x 6 and 2 x 2 blocks
]


@pytest.mark.parametrize("balanced", [False, True])
@pytest.mark.parametrize("D,N,T,k1,k2", CASES)
def test_giekf_gradient_matches_oracle(nsagp, gpu_lib, D, N, T, k1, k2, balanced):
    from oracle import giekf
    pb = make_problem(nsagp, D, N, T, k1, k2, seed=40 + D, w_lik=1e-2)
    eo, go = giekf.gf_giekf_modulator_nmf(pb["w"], pb["t"], pb["y"], _ss_ref_with_derivs(), None, None, k1, k2, 1, D, N, 1, 1,
                                          GradObj="on", balance_derivatives=balanced)
    eg, gg = nsagp.gf_giekf_modulator_nmf(pb["w"], pb["t"], pb["y"], pb["ss_gpu"], None, None, k1, k2, 1, D, N, 1, 1,
                                          GradObj="on", balance_derivatives=balanced)
    assert go.shape == gg.shape == (1 + 3 * D + 2 * N,)
    assert abs(eg - eo) < TOL * abs(eo)
    assert np.all(np.abs(gg - go) <= TOL * np.abs(go)), np.c_[gg, go]
    # the energy is the one the GradObj = 'off' call reports
    e_off, g_off = nsagp.gf_giekf_modulator_nmf(pb["w"], pb["t"], pb["y"], pb["ss_gpu"], None, None, k1, k2, 1, D, N, 1, 1)
    assert abs(e_off - eg) < 1e-10 * abs(eg) and not np.any(g_off)


def test_giekf_gradient_dense_stacks_from_the_closure(nsagp, gpu_lib):
    """A reference-style ss closure hands back dense n x n x P stacks: same result as the block form."""
    D, N, T, k1, k2 = 4, 2, 90, "matern32", "matern52"
    pb = make_problem(nsagp, D, N, T, k1, k2, seed=7, w_lik=1e-2)

    def ss_dense(x, p1, p2, a, b):
        r = nsagp.ss_modulators_nmf(p1, p2, a, b)
        return r[:5] + tuple(np.asarray(s) for s in r[5:])

    e1, g1 = nsagp.gf_giekf_modulator_nmf(pb["w"], pb["t"], pb["y"], pb["ss_gpu"], None, None, k1, k2, 1, D, N, 1, 1, GradObj="on")
    e2, g2 = nsagp.gf_giekf_modulator_nmf(pb["w"], pb["t"], pb["y"], ss_dense, None, None, k1, k2, 1, D, N, 1, 1, GradObj="on")
    assert e1 == e2 and np.array_equal(g1, g2)


def test_giekf_balanced_gradient_is_the_derivative_of_the_energy(nsagp, gpu_lib):
    """C4's kernels at D = 16 (n = 41, 55 parameters), T = 4000: central differences of the GPU energy."""
    D, N, T, k1, k2 = 16, 3, 4000, "exp", "matern52"
    pb = make_problem(nsagp, D, N, T, k1, k2, seed=3, w_lik=1e-2)
    run = lambda w, **kw: nsagp.gf_giekf_modulator_nmf(w, pb["t"], pb["y"], pb["ss_gpu"], None, None, k1, k2, 1, D, N, 1, 1, **kw)
    e, g = run(pb["w"], GradObj="on", balance_derivatives=True)
    assert np.all(np.isfinite(g)) and g.size == 1 + 3 * D + 2 * N
    rng = np.random.default_rng(0)
    for p in np.concatenate([[0], rng.choice(np.arange(1, g.size), 8, replace=False)]):
        h = 1e-6
        wp = pb["w"].copy(); wp[p] += h
        wm = pb["w"].copy(); wm[p] -= h
        fd = (run(wp)[0] - run(wm)[0]) / (2 * h)
        assert abs(fd - g[p]) < 1e-5 * max(abs(fd), 1e-3 * np.max(np.abs(g))), (p, fd, g[p])


def test_giekf_gradient_c4_shape_runs(nsagp, gpu_lib):
    """C4 itself: D = 32 exp subbands, N = 3 matern52 modulators (n = 73, 103 parameters = 103 CTAs)."""
    D, N, T, k1, k2 = 32, 3, 2000, "exp", "matern52"
    pb = make_problem(nsagp, D, N, T, k1, k2, seed=5, w_lik=1e-2)
    e, g = nsagp.gf_giekf_modulator_nmf(pb["w"], pb["t"], pb["y"], pb["ss_gpu"], None, None, k1, k2, 1, D, N, 1, 1, GradObj="on")
    e_off, _ = nsagp.gf_giekf_modulator_nmf(pb["w"], pb["t"], pb["y"], pb["ss_gpu"], None, None, k1, k2, 1, D, N, 1, 1)
    assert g.shape == (103,) and np.all(np.isfinite(g)) and abs(e - e_off) < 1e-10 * abs(e)


def test_giekf_gradient_nan_and_limits(nsagp, gpu_lib):
    D, N, T, k1, k2 = 3, 2, 40, "matern32", "matern52"
    pb = make_problem(nsagp, D, N, T, k1, k2, seed=9, w_lik=1e-2)
    y = pb["y"].copy(); y[7] = np.nan                      # a missing sample makes the reference's energy NaN
    e, g = nsagp.gf_giekf_modulator_nmf(pb["w"], pb["t"], y, pb["ss_gpu"], None, None, k1, k2, 1, D, N, 1, 1, GradObj="on")
    assert np.isnan(e) and np.all(np.isnan(g)) and g.size == 1 + 3 * D + 2 * N


def test_giekf_gradient_c4_matern32_shape_matches_oracle(nsagp, gpu_lib):
    """C4's other shape: D = 32 matern32 subbands, N = 3 matern52 modulators (n = 137, 103 parameters).  2 n^2 doubles
    do not fit one CTA's shared memory, so dP_j lives in an HBM scratch (csrc/ekfgrad.cuh: dP_hbm) -- same code path
    otherwise; checked entry by entry against the oracle's dense recursion on a short signal."""
    from oracle import giekf
    D, N, T, k1, k2 = 32, 3, 25, "matern32", "matern52"
    pb = make_problem(nsagp, D, N, T, k1, k2, seed=9, w_lik=1e-2)
    eo, go = giekf.gf_giekf_modulator_nmf(pb["w"], pb["t"], pb["y"], _ss_ref_with_derivs(), None, None, k1, k2, 1, D, N, 1, 1,
                                          GradObj="on")
    eg, gg = nsagp.gf_giekf_modulator_nmf(pb["w"], pb["t"], pb["y"], pb["ss_gpu"], None, None, k1, k2, 1, D, N, 1, 1,
                                          GradObj="on")
    assert gg.shape == go.shape == (103,)
    assert abs(eg - eo) < TOL * abs(eo)
    assert np.all(np.abs(gg - go) <= TOL * np.abs(go) + 1e-300), np.c_[gg, go]
    # a longer signal runs and agrees with the energy-only call
    pb = make_problem(nsagp, D, N, 1500, k1, k2, seed=10, w_lik=1e-2)
    e, g = nsagp.gf_giekf_modulator_nmf(pb["w"], pb["t"], pb["y"], pb["ss_gpu"], None, None, k1, k2, 1, D, N, 1, 1, GradObj="on")
    e_off, _ = nsagp.gf_giekf_modulator_nmf(pb["w"], pb["t"], pb["y"], pb["ss_gpu"], None, None, k1, k2, 1, D, N, 1, 1)
    assert np.all(np.isfinite(g)) and abs(e - e_off) < 1e-10 * abs(e)
