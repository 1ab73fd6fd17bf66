"""GPU parity: kernel_ss_kalmanFastFB (stationary filter bank) through the C ABI vs the oracle
restatement of matlab/unifying_prob_tf/kernel_ss_kalmanFastFB.m on the same seeded inputs.  The passes are chunked
scans (csrc/fastfb.cuh): tolerance class 1e-6 (observed ~1e-12)."""
import importlib

import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _signal(D, T, seed, gaps):
    rng = np.random.default_rng(seed)
    om = np.linspace(np.pi / 3, np.pi / 40, D)
    t = np.arange(T)
    y = sum(rng.uniform(0.5, 1.5) * np.cos(om[d] * t + rng.uniform(0, 6)) * np.exp(-((t - T / 2) / (T / 3)) ** 2) for d in range(D))
    y = y + 0.05 * rng.standard_normal(T)
    if gaps:
        for s in rng.integers(10, T - 60, 5):
            y[s:s + int(rng.integers(3, 50))] = np.nan
        y[T - 1] = np.nan                      # a missing last sample
    return om, y


@pytest.mark.parametrize("D,kernel,T,gaps,KF", [
    (4, "exp", 300, False, 0), (8, "matern32", 1500, False, 0), (8, "matern32", 1500, True, 0), (16, "exp", 2000, True, 1),
    (12, "matern52", 700, True, 0), (3, "exp", 1, False, 0), (3, "exp", 257, True, 0),
])
def test_fastfb_matches_oracle(nsagp, gpu_lib, D, kernel, T, gaps, KF):
    from oracle import filterbank as ofb
    fb = importlib.import_module(nsagp.__name__ + ".filterbank")
    om, y = _signal(D, T, 7 + D + T, gaps)
    rng = np.random.default_rng(3)
    lamx = 1.0 / rng.uniform(40, 300, D)
    varx = rng.uniform(0.2, 1.0, D)
    Ao, Qo, Ho, Po, Ko, tau = ofb.get_disc_model(lamx, varx, om, D, kernel)
    Ag, Qg, Hg, Pg, Kg, taug = fb.get_disc_model(lamx, varx, om, D, kernel)
    assert (Ko, tau) == (Kg, taug)
    for a, b in ((Ag, Ao), (Qg, Qo), (Hg, Ho), (Pg, Po)):
        assert rel_err(a, b) < 1e-12
    lo, Xo, Pfo = ofb.kernel_ss_kalmanFastFB(Ao, Qo, Ho, Po, Ko, 0.01, y, 0, KF)
    lg, Xg, Pfg = fb.kernel_ss_kalmanFastFB(Ao, Qo, Ho, Po, Ko, 0.01, y, 0, KF)
    assert Xg.shape == Xo.shape and Pfg.shape == Pfo.shape
    assert abs(lg - lo) < 1e-8 * abs(lo)
    assert rel_err(Xg, Xo) < 1e-6 and rel_err(Pfg, Pfo) < 1e-10
