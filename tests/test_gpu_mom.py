"""GPU parity: moment matching kernels (both device forms) vs the oracle restatement
of matlab/likModulatorNMFPower.m / experiments/likModulatorPreCalcwn.m.

Tolerance: 1e-9 relative to the largest magnitude of each output (north_star asks
1e-8 for same-order arithmetic; the kernels differ from the oracle only in
summation order and in using reciprocals for two divisions)."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _random_cavities(rng, D, N, T, s2z=(1e-4, 1e-2), s2g=(1e-3, 0.5)):
    mu = np.concatenate([rng.normal(0, 0.1, (D, T)), rng.normal(0.5, 1.5, (N, T))])
    s2 = np.concatenate([np.exp(rng.uniform(np.log(s2z[0]), np.log(s2z[1]), (D, T))),
                         np.exp(rng.uniform(np.log(s2g[0]), np.log(s2g[1]), (N, T)))])
    return mu, s2


def _oracle_batch(kind, shift, p, hyp, y, mu, s2, W, alpha, N):
    from oracle import cubature as ocub, lik as olik
    link = olik.softplus_link(shift)
    T = y.size
    M = mu.shape[0]
    lZ = np.empty(T); d1 = np.empty((M, T)); d2 = np.empty((M, T))
    wo, xo = ocub.utp_ws(p, N) if p in (3, 5, 7, 9) else (None, None)
    for k in range(T):
        if kind == "power":
            lZ[k], d1[:, k], d2[:, k] = olik.likModulatorNMFPower(link, hyp, y[k], mu[:, k], s2[:, k], W, p, alpha)
        else:
            lZ[k], d1[:, k], d2[:, k] = olik.likModulatorPreCalcwn(link, hyp, y[k], mu[:, k], s2[:, k], W, alpha, wo, xo)
    return lZ, d1, d2


@pytest.mark.parametrize("D,N,p,kind,shift,alpha", [
    (16, 3, 9, "precalc", 1.0, 0.75),     # BASELINE C2/C3 likelihood
    (16, 3, 9, "precalc", 1.0, 1.0),
    (10, 2, 9, "power", 0.0, 0.5),        # demo_toy_modulators_nmf (C1)
    (10, 2, 9, "power", 0.0, 1.0),
    (5, 2, 7, "power", 0.0, 0.5),
    (12, 3, 7, "precalc", 1.0, 0.1),      # train_model.m settings
    (16, 3, 5, "power", 1.0, 0.75),
    (6, 2, 3, "power", 0.0, 1.0),
    (20, 4, 5, "power", 0.0, 0.5),        # D > 16 path, N = 4
    (4, 2, 6, "power", 0.0, 0.5),         # Gauss-Hermite 6^2 points (mvhermgauss)
])
@pytest.mark.parametrize("warp_form", [False, True])
def test_mom_matches_oracle(nsagp, gpu_lib, D, N, p, kind, shift, alpha, warp_form):
    rng = np.random.default_rng(1234 + D * 7 + N + p)
    T = 300
    W = 0.1 * np.abs((2.0 * rng.random((D, N))) ** 2 - 0.2)
    mu, s2 = _random_cavities(rng, D, N, T)
    a = np.log1p(np.exp(mu[D:] - shift)).T @ W.T
    if kind == "precalc":
        a = np.sqrt(a)
    y = np.sum(a * mu[:D].T, axis=1) + rng.normal(0, 0.02, T)
    y[5] = np.nan                                     # missing sample
    hyp = np.log([1e-4])
    if kind == "power":
        mom = nsagp.likModulatorNMFPower(nsagp.Softplus(shift), p, N)
    else:
        wn, xn = nsagp.utp_ws(p, N)
        mom = nsagp.likModulatorPreCalcwn(nsagp.Softplus(shift), wn, xn)
    lZ, d1, d2 = mom.batch(hyp, y, mu, s2, W, alpha, warp_form=warp_form)
    lZo, d1o, d2o = _oracle_batch(kind, shift, p, hyp, y, mu, s2, W, alpha, N)
    assert rel_err(lZ, lZo) < TOL
    assert rel_err(d1, d1o) < TOL
    assert rel_err(d2, d2o) < TOL


def test_mom_handle_signature(nsagp, gpu_lib):
    """The descriptor is callable like the reference's function handle."""
    from oracle import lik as olik
    rng = np.random.default_rng(7)
    D, N = 10, 2
    W = 0.1 * np.abs((2.0 * rng.random((D, N))) ** 2 - 0.2)
    mu, s2 = _random_cavities(rng, D, N, 1)
    yall = np.array([0.0, 0.03, -0.01])
    mom = nsagp.likModulatorNMFPower(nsagp.Softplus(0.0), 9, N)
    lZ, d1, d2 = mom(np.log([1e-4]), mu[:, 0], s2[:, 0], W, 0.5, yall, 1)
    lZo, d1o, d2o = olik.likModulatorNMFPower(olik.softplus_link(0.0), np.log([1e-4]), yall[1], mu[:, 0], s2[:, 0], W, 9, 0.5)
    assert abs(lZ - lZo) < TOL * max(1.0, abs(lZo))
    assert rel_err(d1, d1o) < TOL and rel_err(d2, d2o) < TOL


def test_mom_jitter_floor(nsagp, gpu_lib):
    """Z is floored at 1e-10 (likModulatorNMFPower.m:28,55): an absurd observation
    gives lZ = log(1e-10) exactly and tiny derivatives."""
    rng = np.random.default_rng(3)
    D, N = 10, 2
    W = 0.1 * np.abs((2.0 * rng.random((D, N))) ** 2 - 0.2)
    mu, s2 = _random_cavities(rng, D, N, 4)
    y = np.full(4, 50.0)
    mom = nsagp.likModulatorNMFPower(nsagp.Softplus(0.0), 9, N)
    for wf in (False, True):
        lZ, d1, d2 = mom.batch(np.log([1e-4]), y, mu, s2, W, 1.0, warp_form=wf)
        assert np.allclose(lZ, np.log(1e-10), rtol=0, atol=1e-12)


# ---- straight-line FP64 routines of the sequential passes (csrc/fastmath.cuh) ----
def _fast(gpu_lib, nsagp, op, x):
    x = np.ascontiguousarray(x, float)
    out = np.empty_like(x)
    nsagp._lib.check(gpu_lib.nsagp_fastmath_eval(op, x.size, nsagp._lib.dptr(x), nsagp._lib.dptr(out)))
    return out


def _ulps(a, b):
    return np.max(np.abs(a - b) / np.spacing(np.abs(b)))


def test_fastmath_accuracy(nsagp, gpu_lib):
    rng = np.random.default_rng(0)
    pos = np.exp(rng.uniform(np.log(1e-250), np.log(1e250), 200000))
    sgn = np.where(rng.random(pos.size) < 0.5, -1.0, 1.0)
    ld = np.longdouble
    ref_rsqrt = (1.0 / np.sqrt(pos.astype(ld))).astype(float)
    u_rcp = [_ulps(_fast(gpu_lib, nsagp, op, pos * sgn), 1.0 / (pos * sgn)) for op in (0, 6)]
    u_rsq = [_ulps(_fast(gpu_lib, nsagp, op, pos), ref_rsqrt) for op in (1, 7)]
    u_sqr = [_ulps(_fast(gpu_lib, nsagp, op, pos), np.sqrt(pos)) for op in (2, 8)]
    print("ulps rcp %s rsqrt %s sqrt %s (full, one Newton step fewer)" % (u_rcp, u_rsq, u_sqr))
    assert u_rcp[0] <= 2 and u_rsq[0] <= 2 and u_sqr[0] <= 2
    assert u_rcp[1] <= 2 and u_rsq[1] <= 2 and u_sqr[1] <= 2        # the variants the moment warps use
    assert _fast(gpu_lib, nsagp, 2, np.zeros(3)).tolist() == [0.0, 0.0, 0.0]
    assert _fast(gpu_lib, nsagp, 8, np.zeros(3)).tolist() == [0.0, 0.0, 0.0]
    xe = np.concatenate([rng.uniform(-700, 700, 200000), rng.uniform(-2, 2, 100000), [0.0, -707.9, 708.9]])
    assert _ulps(_fast(gpu_lib, nsagp, 3, xe), np.exp(xe.astype(ld)).astype(float)) <= 2
    assert np.all(_fast(gpu_lib, nsagp, 3, np.array([-708.0, -745.0, -1e4, -np.inf])) == 0.0)
    big = _fast(gpu_lib, nsagp, 3, np.array([709.5, 1000.0, np.inf, np.nan]))
    assert np.all(np.isfinite(big[:3])) and np.all(big[:3] > 8e307) and np.isnan(big[3])   # saturates, NaN propagates
    u = np.concatenate([1.0 + np.exp(rng.uniform(-40, 40, 200000)), rng.uniform(1, 3, 100000), [1.0, 2.0 ** 0.5, 2.0]])
    lg = _fast(gpu_lib, nsagp, 4, u)
    ref = np.log(u.astype(ld)).astype(float)
    nz = ref != 0
    assert np.max(np.abs(lg[nz] - ref[nz]) / np.abs(ref[nz])) < 1e-15
    assert np.all(lg[~nz] == 0.0)
    assert np.isnan(_fast(gpu_lib, nsagp, 4, np.array([np.nan]))[0])
    # the link literally: log(1 + exp(x)), including the rounding of 1 + exp(x); the reference
    # value is built from the GPU's own exp (1 ulp of exp moves log(1+exp) by up to 1e-13 relative)
    xs = rng.uniform(-60, 60, 200000)
    ex = _fast(gpu_lib, nsagp, 3, xs)
    lit = np.log((1.0 + ex).astype(ld)).astype(float)
    sp = _fast(gpu_lib, nsagp, 5, xs)
    nz = lit != 0
    assert np.max(np.abs(sp[nz] - lit[nz]) / np.abs(lit[nz])) < 1e-15
    assert np.all(sp[~nz] == 0.0)
    # the table-driven logarithm of the moment warps (no division on the chain; op 9) and the link built on it (op 10)
    lt = _fast(gpu_lib, nsagp, 9, u)
    nz = ref != 0
    assert np.max(np.abs(lt[nz] - ref[nz]) / np.abs(ref[nz])) < 1e-15 and _ulps(lt[nz], ref[nz]) <= 2
    assert np.all(lt[~nz] == 0.0)
    near1 = 1.0 + np.exp(rng.uniform(np.log(2.3e-16), np.log(1e-2), 100000))          # results down to one ulp of 1
    assert _ulps(_fast(gpu_lib, nsagp, 9, near1), np.log(near1.astype(ld)).astype(float)) <= 2
    st = _fast(gpu_lib, nsagp, 10, xs)
    nz = lit != 0
    assert np.max(np.abs(st[nz] - lit[nz]) / np.abs(lit[nz])) < 1e-15
    assert np.all(st[~nz] == 0.0)
