"""CPU: host-side logic of the product (no GPU compute): parameter unpacking, input
merging, block-model extraction, steady-state tables, the C ABI's exported symbols,
loud failure without a device, and the multi-rank sharding (gloo, world_size 2)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, make_problem


def test_abi_exports_every_declared_symbol(nsagp):
    """libnsagp.so loads and exports every function include/nsagp.h declares."""
    hdr = open(os.path.join(ROOT, "include", "nsagp.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(nsagp_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 18
    assert declared == set(nsagp._lib.EXPORTS)
    L = ctypes.CDLL(nsagp._lib.LIB_PATH)
    for sym in declared:
        assert hasattr(L, sym), sym
    assert b"sm_100a" in ctypes.cast(ctypes.CDLL(nsagp._lib.LIB_PATH).nsagp_version, ctypes.CFUNCTYPE(ctypes.c_char_p))()


def test_library_is_sm100a_only(nsagp):
    out = subprocess.run(["cuobjdump", "-lelf", nsagp._lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_device(nsagp):
    """On a box without a GPU the product path raises; it never computes on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    pb = make_problem(nsagp, 4, 2, 50, "matern32", "matern52", seed=1)
    with pytest.raises(nsagp.NsagpError) as ei:
        nsagp.gf_ep_modulator_nmf(pb["w"], pb["t"], pb["y"], pb["ss_gpu"], pb["mom_gpu"], pb["t"], "matern32",
                                  "matern52", 1, 4, 2, 0.5, [0.5], 1)
    assert ei.value.status == -2
    with pytest.raises(nsagp.NsagpError):
        pb["mom_gpu"](np.log([1e-4]), np.zeros(6), np.ones(6), pb["hyp"].W, 1.0, pb["y"], 0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "nonstationary-audio-gp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inc", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "oracle/" not in src, f


def test_mom_must_be_a_descriptor(nsagp):
    pb = make_problem(nsagp, 4, 2, 30, "matern32", "matern52", seed=1)
    with pytest.raises(TypeError):
        nsagp.gf_ep_modulator_nmf(pb["w"], pb["t"], pb["y"], pb["ss_gpu"], lambda *a: None, pb["t"], "matern32",
                                  "matern52", 1, 4, 2, 0.5, [0.5], 1)
    with pytest.raises(TypeError):
        nsagp.likModulatorNMFPower(np.exp, 9, 2)          # arbitrary link handles cannot run on the GPU


def test_merge_inputs_matches_oracle(nsagp):
    from oracle import gf_ep
    rng = np.random.default_rng(0)
    x = rng.permutation(20).astype(float); y = rng.normal(size=20)
    xt = np.array([3.0, 3.0, 7.5, 25.0, 0.0])
    a, ia = nsagp.merge_inputs(x, y, xt)
    b, ib = gf_ep.merge_inputs(x, y, xt)
    assert np.array_equal(np.isnan(a), np.isnan(b)) and np.allclose(a[~np.isnan(a)], b[~np.isnan(b)])
    assert np.array_equal(ia, ib)


def test_unpack_matches_oracle(nsagp):
    from oracle import ssmodel as oss
    entry = nsagp.entry if hasattr(nsagp, "entry") else __import__("importlib").import_module("nonstationary-audio-gp_b200.entry")
    rng = np.random.default_rng(1)
    D, N = 5, 2
    w = rng.normal(size=1 + 3 * D + 2 * N + D * N)
    for a, b in zip(entry._unpack_log(w, 1, D, N), oss.unpack_log(w, 1, D, N)):
        assert np.array_equal(a, b)
    cons = np.array([[0.0, 0.1], [50.0, 1000.0], [0.0, 3.2], [0.0, 20.0], [100.0, 3000.0], [0.0, 1.25]])
    for tune in ([1, 1, 1, 1, 1, 1, 1], [0, 1, 0, 1, 0, 1, 0], [1, 0, 0, 0, 0, 0, 1], [0, 0, 0, 0, 0, 0, 0]):
        sizes = [1, D, D, D, N, N, D * N]
        full = rng.normal(size=sum(sizes))
        idx = np.cumsum([0] + sizes)
        wt = np.concatenate([full[idx[i]:idx[i + 1]] for i in range(7) if tune[i]] + [np.zeros(0)])
        wf = np.concatenate([full[idx[i]:idx[i + 1]] for i in range(7) if not tune[i]] + [np.zeros(0)])
        for a, b in zip(entry._unpack_constrained(wt, 1, D, N, cons, wf, tune),
                        oss.unpack_constraints(wt, 1, D, N, cons, wf, tune)):
            assert np.allclose(a, b, rtol=0, atol=0)
    assert np.allclose(nsagp.inv_sigmoid(nsagp.sigmoid([-2.0, 0.3], (1, 5)), (1, 5)), [-2.0, 0.3])
    with pytest.raises(ValueError):
        nsagp.inv_sigmoid([6.0], (1, 5))
    assert np.allclose(nsagp.lambda_map([2.0], "matern52"), oss.lambda_map([2.0], "matern52"))


@pytest.mark.parametrize("k1,k2", [("exp", "matern52"), ("matern32", "matern52"), ("matern52", "matern32"), ("matern72", "exp")])
def test_model_builder_matches_oracle(nsagp, k1, k2):
    from oracle import ssmodel as oss
    rng = np.random.default_rng(2)
    D, N = 4, 3
    hyp = nsagp.synth.demo_hypers(D, N, rng)
    a = nsagp.ss_modulators_nmf(hyp.w_sub(), hyp.w_mod(), k1, k2)[:5]
    b = oss.ss_modulators_nmf(hyp.w_sub(), hyp.w_mod(), k1, k2)
    for x, yv in zip(a, b):
        assert np.array_equal(x, yv)
    A, Q = nsagp.lti_disc(a[0], a[1], a[2], 1.0)
    Ao, Qo = oss.lti_disc(b[0], b[1], b[2], 1.0)
    assert np.allclose(A, Ao, rtol=1e-13, atol=1e-15) and np.allclose(Q, Qo, rtol=1e-11, atol=1e-300)
    mdl = nsagp.to_block_model(A, Q, a[3], a[4], D, N)
    A2, Q2, H2, P2 = mdl.dense()
    assert np.array_equal(A2, A) and np.array_equal(Q2, Q) and np.array_equal(H2, a[3]) and np.array_equal(P2, a[4])
    assert mdl.n == A.shape[0] and mdl.M == D + N
    Fb, Lb, Hb, Pb = nsagp.ssmodel.balance(a[0], a[1], a[3], a[4])
    Fo, Lo, Ho, Po, _ = oss.balance_ss(b[0], b[1], b[3], b[4])
    assert np.allclose(Fb, Fo) and np.allclose(Hb, Ho) and np.allclose(Pb, Po)
    assert all(np.count_nonzero(Hb[i]) == 1 for i in range(D + N))     # balance is a diagonal scaling here


def test_block_model_rejects_coupled_models(nsagp):
    rng = np.random.default_rng(3)
    hyp = nsagp.synth.demo_hypers(3, 2, rng)
    F, L, Qc, H, Pinf = nsagp.ss_modulators_nmf(hyp.w_sub(), hyp.w_mod(), "matern32", "matern52")[:5]
    A, Q = nsagp.lti_disc(F, L, Qc, 1.0)
    Abad = A.copy(); Abad[0, -1] = 1e-3
    with pytest.raises(ValueError):
        nsagp.to_block_model(Abad, Q, H, Pinf, 3, 2)
    Hbad = H.copy(); Hbad[0, 5] = 0.1
    with pytest.raises(ValueError):
        nsagp.to_block_model(A, Q, Hbad, Pinf, 3, 2)
    with pytest.raises(ValueError):
        nsagp.to_block_model(A, Q, H, Pinf, 2, 2)


@pytest.mark.parametrize("native", [True, False])
def test_tables_match_oracle(nsagp, native):
    """native=True: the library's host routine nsagp_ihgp_tables (doubling algorithm, no GPU needed);
    native=False: the SciPy route.  Both against the oracle's dare restatement."""
    from oracle import ihgp_ep
    rng = np.random.default_rng(4)
    D, N = 3, 2
    hyp = nsagp.synth.demo_hypers(D, N, rng)
    F, L, Qc, H, Pinf = nsagp.ss_modulators_nmf(hyp.w_sub(), hyp.w_mod(), "exp", "matern52")[:5]
    F, L, H, Pinf = nsagp.ssmodel.balance(F, L, H, Pinf)
    A, Q = nsagp.lti_disc(F, L, Qc, 1.0)
    Q = (Q + Q.T) / 2
    mdl = nsagp.to_block_model(A, Q, H, Pinf, D, N)
    tb = nsagp.tables.build_tables(mdl, want_smoother=True, native=native)
    ot = ihgp_ep.ihgp_setup(A, Q, H)
    assert np.array_equal(tb.r, ot["r"])
    for n in range(D + N):
        if native:      # another solver: compare in the max norm (entries that are ~0 differ at rounding level)
            for a, o in ((tb.PP[n], ot["PPlist"][n]), (tb.PG[n], ot["PGlist"][n])):
                assert np.max(np.abs(a - o)) <= 1e-10 * np.max(np.abs(o))
        else:
            assert np.allclose(tb.PP[n], ot["PPlist"][n], rtol=1e-9, atol=1e-300)
            assert np.allclose(tb.PG[n], ot["PGlist"][n], rtol=1e-7, atol=1e-14)
    pp, pg = tb.packed()
    assert pp.size == 200 * (D * mdl.bz ** 2 + N * mdl.bg ** 2) and pg.size == 2 * pp.size


def test_synthetic_signal_statistics(nsagp):
    rng = np.random.default_rng(5)
    hyp = nsagp.synth.demo_hypers(6, 2, rng)
    y, zf, g = nsagp.synth.sample_signal(hyp, "matern32", "matern52", 20000, rng)
    assert y.shape == (20000,) and zf.shape == (20000, 6) and g.shape == (20000, 2)
    assert 0.05 < zf.std() < 0.2                           # var_fast = 0.01
    a = np.log1p(np.exp(g)) @ hyp.W.T
    assert np.allclose(y, np.sum(zf * a, axis=1))
    yg = nsagp.synth.add_gaps(y, rng)
    assert 10 <= np.isnan(yg).sum() <= 6 * 320 + 320
    w = hyp.pack_log()
    assert w.size == 1 + 3 * 6 + 2 * 2 + 12 and np.isclose(np.exp(w[0]), 1e-4)


def test_shard_range_covers_everything(nsagp):
    for B in (0, 1, 7, 256, 1000):
        for world in (1, 2, 3, 8):
            spans = [nsagp.batch.shard_range(B, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


_WORKER = r"""
import importlib, os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
nsagp = importlib.import_module("nonstationary-audio-gp_b200")
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
calls = []
def evaluator(ws, ys):            # stands in for the GPU evaluation: the test is about the sharding
    calls.append(len(ws))
    return [float(np.sum(w) + 10.0 * y[0]) for w, y in zip(ws, ys)]
B = 7
ws = [np.full(3, float(i)) for i in range(B)]
ys = [np.array([float(i)]) for i in range(B)]
out = nsagp.batch.nlz_batch(ws, ys, evaluator)
ref = np.array([3.0 * i + 10.0 * i for i in range(B)])
assert np.array_equal(out, ref), out
lo, hi = nsagp.batch.shard_range(B, 2, dist.get_rank())
assert calls == [hi - lo]
f0, g = nsagp.batch.finite_difference_gradient(np.array([1.0, 2.0, 3.0]), np.array([0.5]), evaluator, h=1e-3)
assert abs(f0 - 11.0) < 1e-12 and np.allclose(g, 1.0, atol=1e-9)
dist.destroy_process_group()
print("ok")
"""


def test_nlz_batch_sharding_gloo_world2(tmp_path):
    """N > 1 path on CPU: two gloo ranks shard 7 units, gather B scalars, agree."""
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0 and o.strip().endswith("ok"), o


def test_row_guess_from_float_bits_is_within_the_lookup_window():
    """The kernels guess the table row of R (or of ttau = 1/R) from the high 32 bits of the double -- a piecewise-linear
    log2 -- and confirm it by counting six thresholds, rows guess-3 .. guess+2 (csrc/ihgp.cuh: count_le_guess,
    csrc/adfcta.cuh: lookup_by_ttau; coefficients: api.cu rg_a / rg_b, make_ttau_guess).  For the reference's grid
    r = logspace(-2, 4, 200) the guess must land within two rows of the nearest-neighbour answer for every R, inside and
    outside the grid, or the kernels fall back to a binary search on every look-up (correct, but slow)."""
    nr = 200
    r = np.logspace(-2, 4, nr)
    scale = (nr - 1) / (np.log10(r[-1]) - np.log10(r[0]))
    l2 = 0.30102999566398120 * scale
    R = np.concatenate([np.logspace(-4, 6, 200001), r, r * (1 + 1e-12), r * (1 - 1e-12), 0.5 * (r[1:] + r[:-1])])
    truth = np.argmin(np.abs(r[None, :] - R[:, None]), axis=1)               # [~, ind] = min(abs(r - R)), first on ties
    hi = (R.view(np.uint64) >> np.uint64(32)).astype(np.float64)
    guess_R = np.clip(np.floor(-np.log10(r[0]) * scale - (1023.0 - 0.043) * l2 + 0.49 + hi * (l2 / 1048576.0)), 0, nr - 1)
    assert np.max(np.abs(guess_R - truth)) <= 2
    tt = 1.0 / R
    hit = (tt.view(np.uint64) >> np.uint64(32)).astype(np.float64)
    guess_t = np.clip(np.floor(-np.log10(r[0]) * scale + (1023.0 - 0.043) * l2 + 0.49 - hit * (l2 / 1048576.0)), 0, nr - 1)
    assert np.max(np.abs(guess_t - truth)) <= 2
    # and it is usually exact or one off: the six-entry window is not wasteful
    assert np.mean(np.abs(guess_R - truth) <= 1) > 0.999 and np.mean(np.abs(guess_t - truth) <= 1) > 0.999
