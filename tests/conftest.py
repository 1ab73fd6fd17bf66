import importlib
import os
import sys

os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")      # before CUDA initialises (see _lib.py)

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG = "nonstationary-audio-gp_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def nsagp():
    return importlib.import_module(PKG)


@pytest.fixture(scope="session")
def gpu_lib(nsagp):
    """The CUDA library, built and loaded; fails loudly (no fallback) if absent."""
    L = nsagp._lib.lib()
    assert L.nsagp_device_count() > 0, "no CUDA device visible"
    # kernel parity is measured on the oracle's own steady-state tables (SciPy's Riccati solver); the library's
    # native table routine is compared with them in test_host_logic.py and end to end in test_gpu_ihgp.py
    nsagp.tables.DEFAULT_NATIVE = False
    return L


def make_problem(nsagp, D, N, T, kernel1, kernel2, seed, kind="power", p=9, shift=0.0, speech=False,
                 gaps=False, w_lik=1e-4):
    """Synthetic problem in the reference's own terms: returns a dict with the
    packed log-parameters, the signal and both closures for oracle and product."""
    from oracle import lik as olik, ssmodel as oss
    rng = np.random.default_rng(seed)
    hyp = (nsagp.synth.speech_hypers if speech else nsagp.synth.demo_hypers)(D, N, rng, w_lik=w_lik)
    y, zf, g = nsagp.synth.sample_signal(hyp, kernel1, kernel2, T, rng, link_shift=shift,
                                         sqrt_model=(kind == "precalc"))
    if gaps:
        y = nsagp.synth.add_gaps(y, rng, n_gaps_per_20k=max(6, int(3 * 20000 / T)), min_len=3, max_len=max(4, T // 20))
    wn, xn = nsagp.utp_ws(p, N)
    if kind == "power":
        mom_gpu = nsagp.likModulatorNMFPower(nsagp.Softplus(shift), p, N)
        mom_ref = olik.make_mom("power", olik.softplus_link(shift), p=p)
    else:
        mom_gpu = nsagp.likModulatorPreCalcwn(nsagp.Softplus(shift), wn, xn)
        from oracle import cubature as ocub
        wo, xo = ocub.utp_ws(p, N)
        mom_ref = olik.make_mom("precalc", olik.softplus_link(shift), wn=wo, xn_unscaled=xo)
    return dict(hyp=hyp, w=hyp.pack_log(), y=y, t=np.arange(1.0, T + 1.0), zf=zf, g=g, D=D, N=N, T=T,
                kernel1=kernel1, kernel2=kernel2, mom_gpu=mom_gpu, mom_ref=mom_ref,
                ss_gpu=lambda x, p1, p2, k1, k2: nsagp.ss_modulators_nmf(p1, p2, k1, k2),
                ss_ref=lambda x, p1, p2, k1, k2: oss.ss_modulators_nmf(p1, p2, k1, k2))


def rel_err(a, b):
    """max |a-b| / max |b| over finite entries; NaN/Inf patterns must coincide."""
    a = np.asarray(a, float); b = np.asarray(b, float)
    assert a.shape == b.shape, (a.shape, b.shape)
    fa, fb = np.isfinite(a), np.isfinite(b)
    assert np.array_equal(fa, fb), "finite masks differ at %d entries" % int(np.sum(fa != fb))
    nf = ~fb
    if nf.any():
        assert np.array_equal(np.isnan(a[nf]), np.isnan(b[nf])) and np.array_equal(np.sign(a[nf][~np.isnan(a[nf])]),
                                                                                np.sign(b[nf][~np.isnan(b[nf])]))
    if not fb.any():
        return 0.0
    scale = max(np.max(np.abs(b[fb])), 1e-300)
    return float(np.max(np.abs(a[fb] - b[fb])) / scale)
