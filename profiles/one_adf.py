#!/usr/bin/env python
"""One exact first pass (ADF) of the C2 model at T = argv[1] (default 20000): the workload for an ncu source-level capture
of ihgp_adf_cta_kernel (ep_itts = 1, predict mode: ADF + one smoother scan)."""
import importlib, json, os, sys
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nsagp = importlib.import_module("nonstationary-audio-gp_b200")
import bench_workloads as bw
lm = nsagp._lib
T = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
kind = sys.argv[2] if len(sys.argv) > 2 else "ihgp"
rng = np.random.default_rng(2026)
hyp = nsagp.synth.speech_hypers(16, 3, rng)
y = bw._tiled_signal(nsagp, hyp, "exp", "matern52", T, 7000)
mdl, tabs = bw._model(nsagp, hyp, "exp", "matern52", 16, 3, kind == "ihgp")
with nsagp.Plan(lm.KIND_IHGP if kind == "ihgp" else lm.KIND_FULL, [mdl], [(bw._mom(nsagp), np.log([hyp.w_lik]), hyp.W)], 0.75, [0.01], 1,
                y[None, :], lm.MODE_PREDICT, tables=[tabs] if kind == "ihgp" else None) as p:
    p.run(); p.run()
    tm = p.timings()
    print(json.dumps(dict(T=T, adf_ms=tm["adf"], cycles_per_step=tm["adf"] * 1e-3 * 1.965e9 / T)))
