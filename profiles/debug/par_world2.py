import importlib, os, sys, threading
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
nsagp = importlib.import_module("nonstationary-audio-gp_b200")
import bench_workloads as bw
import test_gpu_chunked as tc
lm = nsagp._lib
kind = sys.argv[1] if len(sys.argv) > 1 else "full"
T, itts = int(os.environ.get("DBG_T", "200000")), 3
rng = np.random.default_rng(7)
hyp = nsagp.synth.speech_hypers(16, 3, rng)
y = bw._tiled_signal(nsagp, hyp, "exp", "matern52", T, 7000)
mdl, tabs = bw._model(nsagp, hyp, "exp", "matern52", 16, 3, kind == "ihgp")
damp = np.linspace(0.01, 0.1, itts)
mk = lambda: nsagp.Plan(lm.KIND_IHGP if kind == "ihgp" else lm.KIND_FULL, [mdl], [(bw._mom(nsagp), np.log([hyp.w_lik]), hyp.W)], 0.75, damp, itts,
                        y[None, :], lm.MODE_PREDICT, tables=[tabs] if kind == "ihgp" else None)
for world in (1, 2):
    for par in (None, (int(os.environ.get("DBG_CH", "16")), int(os.environ.get("DBG_BURN", "20000")))):
        plans = [mk() for _ in range(world)]
        ranges, cms = tc._run_threads(nsagp, plans, par=par)
        print("world", world, "par", par, "nlZ", [repr(float(v)) for v in plans[0].fetch(0, ("nlZ",))["nlZ"]],
              "mismatch", [p.adf_mismatch() for p in plans], "adf ms", [p.timings()["adf"] for p in plans])
        for p in plans: p.close()
        for c in cms: c.close()
