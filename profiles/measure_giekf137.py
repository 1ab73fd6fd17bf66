#!/usr/bin/env python
"""C4's second shape (D=32 matern32 subbands, N=3 matern52 modulators, dense n=137) through the C ABI: device times of the
filter and of the large-state scan smoother (csrc/ekfbig.cuh).  Usage: python profiles/measure_giekf137.py [T]"""
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
D, N, K1, K2 = 32, 3, "matern32", "matern52"
T = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
rng = np.random.default_rng(3)
hyp = nsagp.synth.speech_hypers(D, N, rng, w_lik=1e-2)
y, _, _ = nsagp.synth.sample_signal(hyp, K1, K2, T, rng)
for s in range(0, T, 20000):
    for j, g in enumerate((10, 20, 40, 80, 160, 320)):
        a = s + 1500 + 3000 * j
        y[a:min(a + g, T)] = np.nan
F, Lm, Qc, H, Pinf = nsagp.ss_modulators_nmf(hyp.w_sub(), hyp.w_mod(), K1, K2)[:5]
F, Lm, H, Pinf = nsagp.ssmodel.balance(F, Lm, H, Pinf)
A, Q = nsagp.lti_disc(F, Lm, Qc, 1.0)
mdl = nsagp.to_block_model(A, Q, H, Pinf, D, N)
arrs = [L.as_f64(a) for a in (mdl.A, mdl.Q, mdl.Pinf, mdl.h)]
cm = L.Model()
cm.D, cm.N, cm.bz, cm.bg = mdl.D, mdl.N, mdl.bz, mdl.bg
cm.A, cm.Q, cm.Pinf, cm.h = [L.dptr(a) for a in arrs]
Wf = np.asfortranarray(hyp.W)
M, n = mdl.M, mdl.n
o = L.Outputs()
bufs = dict(Eft=np.zeros((T, M)), Varft=np.zeros((T, M)))
for k, v in bufs.items():
    setattr(o, k, L.dptr(v))
yb = L.as_f64(y)
best = None
for rep in range(2):
    L.check(L.lib().nsagp_giekf(C.byref(cm), L.dptr(Wf), float(hyp.w_lik), 1, 1, L.dptr(yb), T, L.MODE_PREDICT, C.byref(o)))
    ms = np.zeros(2)
    L.check(L.lib().nsagp_giekf_timings(L.dptr(ms), 2))
    if best is None or ms.sum() < best.sum():
        best = ms.copy()
print(json.dumps(dict(config="C4-137 gf_giekf D=32 N=3 matern32/matern52 n=%d T=%d g_iter=1 gaps" % (n, T), filter_ms=float(best[0]),
                      smoother_ms=float(best[1]), filter_us_per_step=float(best[0]) * 1e3 / T, smoother_us_per_step=float(best[1]) * 1e3 / T,
                      steps_per_s=T / best.sum() * 1e3, finite=bool(np.all(np.isfinite(bufs["Eft"])) and np.all(bufs["Varft"] > 0)),
                      smoother_dense_equiv_tflops=12.3 * n ** 3 * T / (best[1] * 1e-3) / 1e12)))
