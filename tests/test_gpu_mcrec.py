"""GPU parity: Monte-Carlo reconstruction (csrc/mcrec.cuh through the C ABI) vs the oracle restatement of
matlab/demo_toy_modulators_nmf.m:119-165 / experiments/missing_data_music.m:138-176 on the same explicit
normal draws (tolerance 1e-10: Welford vs two-pass variance differ by rounding only), and statistical checks
of the device-side generator (Philox4x32-10 + Box-Muller)."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _marginals(rng, D, N, T):
    M = D + N
    Eft = rng.standard_normal((M, T)) * np.r_[0.3 * np.ones(D), np.ones(N)][:, None]
    Varft = rng.uniform(1e-4, 0.3, (M, T))
    W = 0.1 * np.abs((2.0 * rng.random((D, N))) ** 2 - 0.2)
    return Eft, Varft, W


@pytest.mark.parametrize("D,N,T,s,shift,sqrt_model", [
    (10, 2, 333, 25, 0.0, False),       # demo_toy_modulators_nmf shape
    (16, 3, 200, 40, 1.0, True),        # sqrt model, shifted link (experiments)
    (3, 3, 130, 2, 0.0, False),         # s = 2, D = N
    (20, 4, 64, 7, 1.0, True),          # D > 16
    (4, 2, 1, 1, 0.0, False),           # one step, one sample: variance 0
])
def test_reconstruction_matches_oracle(nsagp, gpu_lib, D, N, T, s, shift, sqrt_model):
    from oracle import mcrec
    rng = np.random.default_rng(D * 100 + T)
    Eft, Varft, W = _marginals(rng, D, N, T)
    Z = rng.standard_normal((T, s, D + N))
    Eo, Vo, Emo, Vmo = mcrec.reconstruct(Eft, Varft, W, Z, shift, sqrt_model)
    g = nsagp.reconstruct_signal(Eft, Varft, W, s=s, link_shift=shift, sqrt_model=sqrt_model, Z=Z)
    assert rel_err(g["Esig"], Eo) < 1e-10 and rel_err(g["Eft_mod"], Emo) < 1e-10
    if s > 1:
        assert rel_err(g["Vsig"], Vo) < 1e-9 and rel_err(g["Varft_mod"], Vmo) < 1e-9
    else:
        assert np.all(g["Vsig"] == 0) and np.all(g["Varft_mod"] == 0)


def test_zero_variance_is_deterministic(nsagp, gpu_lib):
    rng = np.random.default_rng(3)
    D, N, T = 6, 2, 50
    Eft, _, W = _marginals(rng, D, N, T)
    g = nsagp.reconstruct_signal(Eft, np.zeros_like(Eft), W, s=16, seed=9)
    expect = np.sum((W @ np.log(1 + np.exp(Eft[D:]))) * Eft[:D], axis=0)
    assert np.allclose(g["Esig"], expect, rtol=1e-13, atol=1e-15) and np.all(np.abs(g["Vsig"]) < 1e-25)


def test_device_generator_statistics(nsagp, gpu_lib):
    """Generated draws: reproducible for a seed, different across seeds, and the sample moments of a LINEAR
    functional (no modulator uncertainty) match their exact values within Monte-Carlo error."""
    rng = np.random.default_rng(5)
    D, N, T, s = 8, 2, 400, 4000
    Eft, Varft, W = _marginals(rng, D, N, T)
    Varft[D:] = 0.0
    a = nsagp.reconstruct_signal(Eft, Varft, W, s=s, seed=123)
    b = nsagp.reconstruct_signal(Eft, Varft, W, s=s, seed=123)
    c = nsagp.reconstruct_signal(Eft, Varft, W, s=s, seed=124)
    assert np.array_equal(a["Esig"], b["Esig"]) and not np.array_equal(a["Esig"], c["Esig"])
    amp = W @ np.log(1 + np.exp(Eft[D:]))
    mean = np.sum(amp * Eft[:D], axis=0); var = np.sum(amp ** 2 * Varft[:D], axis=0)
    zscore = (a["Esig"] - mean) / np.sqrt(var / s)
    assert abs(np.mean(zscore)) < 4 / np.sqrt(T) and 0.85 < np.std(zscore) < 1.15 and np.max(np.abs(zscore)) < 5.5
    assert np.allclose(a["Vsig"], var, rtol=0.15)
    # draws are independent across time steps: lag-1 correlation of the standardised errors ~ N(0, 1/T)
    assert abs(np.corrcoef(zscore[:-1], zscore[1:])[0, 1]) < 5 / np.sqrt(T)
