#!/usr/bin/env python
"""One predict-mode run of the C2 (ihgp) or C3 (full) model shape on a long signal with the parallel first pass
switched on, so that the frozen-site passes (scans, site update) are what an ncu capture sees.
Usage: python profiles/one_frozen.py ihgp|full [T] [ep_itts]"""
import importlib
import json
import os
import sys

os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nsagp = importlib.import_module("nonstationary-audio-gp_b200")
import bench_workloads as bw
lm = nsagp._lib
kind = sys.argv[1] if len(sys.argv) > 1 else "ihgp"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 2000000
itts = int(sys.argv[3]) if len(sys.argv) > 3 else 3
rng = np.random.default_rng(7)
hyp = nsagp.synth.speech_hypers(16, 3, rng)
y = bw._tiled_signal(nsagp, hyp, "exp", "matern52", T, 7000)
mdl, tabs = bw._model(nsagp, hyp, "exp", "matern52", 16, 3, kind == "ihgp")
with nsagp.Plan(lm.KIND_IHGP if kind == "ihgp" else lm.KIND_FULL, [mdl], [(bw._mom(nsagp), np.log([hyp.w_lik]), hyp.W)], 0.75,
                np.linspace(0.01, 0.1, itts), itts, y[None, :], lm.MODE_PREDICT, tables=[tabs] if kind == "ihgp" else None) as p:
    p.set_adf_parallel(148, 60000)
    p.run()
    p.run()
    tm = p.timings()
    n, M = mdl.n, mdl.M
    sweeps = itts
    print(json.dumps(dict(kind=kind, T=T, ep_itts=itts, phases_ms=tm,
                          frozen_ms_per_sweep=(tm["fixed_filter"] / max(1, itts - 1), tm["smoother"] / itts, tm["site_update"] / max(1, itts - 1)))))
