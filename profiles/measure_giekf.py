#!/usr/bin/env python
"""Device-timed throughput of BASELINE config C4 (gf_giekf_modulator_nmf_constraints: iterated EKF + dense RTS
smoother, D=32 exp subbands x N=3 matern52 modulators, n=73, missing-data gaps) through the C ABI, for the
sequential smoother and the scan smoother (csrc/ekfscan.cuh).  One JSON line per run.
Usage: python profiles/measure_giekf.py [T_scan [T_seq [chunk_len ...]]]
"""
import ctypes as C
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nsagp = importlib.import_module("nonstationary-audio-gp_b200")
L = nsagp._lib
D, N, K1, K2 = 32, 3, "exp", "matern52"


def problem(T, seed=3):
    rng = np.random.default_rng(seed)
    hyp = nsagp.synth.speech_hypers(D, N, rng, w_lik=1e-2)
    y, _, _ = nsagp.synth.sample_signal(hyp, K1, K2, T, rng)
    # six gaps of 10..320 samples per 20k samples (experiments/missing_data_music.m:51,57)
    for s in range(0, T, 20000):
        for j, g in enumerate((10, 20, 40, 80, 160, 320)):
            a = s + 1500 + 3000 * j
            y[a:min(a + g, T)] = np.nan
    F, Lm, Qc, H, Pinf = nsagp.ss_modulators_nmf(hyp.w_sub(), hyp.w_mod(), K1, K2)[:5]
    F, Lm, H, Pinf = nsagp.ssmodel.balance(F, Lm, H, Pinf)
    A, Q = nsagp.lti_disc(F, Lm, Qc, 1.0)
    return hyp, y, nsagp.to_block_model(A, Q, H, Pinf, D, N)


def run(T, form, chunk_len=0, seg_chunks=0, g_iter=1, keep=None):
    hyp, y, mdl = problem(T)
    arrs = [L.as_f64(a) for a in (mdl.A, mdl.Q, mdl.Pinf, mdl.h)]
    cm = L.Model()
    cm.D, cm.N, cm.bz, cm.bg = mdl.D, mdl.N, mdl.bz, mdl.bg
    cm.A, cm.Q, cm.Pinf, cm.h = [L.dptr(a) for a in arrs]
    Wf = np.asfortranarray(hyp.W)
    M, n = mdl.M, mdl.n
    o = L.Outputs()
    bufs = dict(Eft=np.zeros((T, M)), Varft=np.zeros((T, M)), MS=np.zeros((T, n)), maxDiffP=np.zeros(g_iter))
    for k, v in bufs.items():
        setattr(o, k, L.dptr(v))
    L.check(L.lib().nsagp_giekf_config(form, chunk_len, seg_chunks))
    yb = L.as_f64(y)
    best = None
    for rep in range(2):
        L.check(L.lib().nsagp_giekf(C.byref(cm), L.dptr(Wf), float(hyp.w_lik), g_iter, 1, L.dptr(yb), T, L.MODE_PREDICT, C.byref(o)))
        ms = np.zeros(2)
        L.check(L.lib().nsagp_giekf_timings(L.dptr(ms), 2))
        if best is None or ms.sum() < best.sum():
            best = ms.copy()
    L.check(L.lib().nsagp_giekf_config(0, 0, 0))
    rec = dict(config="C4 gf_giekf D=32 N=3 n=%d T=%d g_iter=%d gaps" % (n, T, g_iter),
               smoother="scan (DMMA)" if form != 1 else "first-generation kernels", chunk_len=chunk_len or "auto", seg_chunks=seg_chunks,
               filter_ms=float(best[0]), smoother_ms=float(best[1]),
               filter_us_per_step=float(best[0]) * 1e3 / (T * g_iter), smoother_us_per_step=float(best[1]) * 1e3 / (T * g_iter),
               steps_per_s=T * g_iter / best.sum() * 1e3,
               smoother_dense_equiv_tflops=12.3 * n ** 3 * T * g_iter / (best[1] * 1e-3) / 1e12)
    print(json.dumps(rec), flush=True)
    if keep is not None:
        keep.update(bufs)
    return rec


def main():
    T_scan = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    T_seq = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
    chunk_lens = [int(v) for v in sys.argv[3:]] or [0]
    a, b = {}, {}
    run(T_seq, 1, keep=a)
    run(T_seq, 2, keep=b)
    err = {k: float(np.max(np.abs(a[k] - b[k])) / np.max(np.abs(a[k]))) for k in ("Eft", "Varft", "MS")}
    print(json.dumps(dict(check="scan vs sequential smoother, T=%d" % T_seq, rel_err=err)), flush=True)
    for cl in chunk_lens:
        run(T_scan, 2, chunk_len=cl)


if __name__ == "__main__":
    main()
