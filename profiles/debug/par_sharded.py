import importlib, os, sys, threading
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest
nsagp = importlib.import_module("nonstationary-audio-gp_b200")
nsagp.tables.DEFAULT_NATIVE = False
import test_gpu_chunked as tc
kind = sys.argv[1] if len(sys.argv) > 1 else "ihgp"
T, itts, world = 24000, 3, 3
damping = np.linspace(0.3, 0.2, itts)
pb = tc._short_memory_problem(nsagp, 6, 3, T, 12, kind)
names = ("Eft", "ttau", "tnu", "R", "MF", "MS", "nlZ", "lZ")
single = tc._plans(nsagp, pb, itts, damping, 1, kind)[0]
single.run()
ref = single.fetch(0, names)
for par in (None, (4, 2500)):
    plans = tc._plans(nsagp, pb, itts, damping, world, kind)
    ranges, cms = tc._run_threads(nsagp, plans, par=par)
    got = tc._assemble(plans, ranges, names)
    print("par", par, "ranges", ranges)
    for k in ("Eft", "ttau", "tnu", "R", "MF", "MS"):
        d = np.abs(got[k] - ref[k]); d[~np.isfinite(d)] = 0
        i = np.unravel_index(np.argmax(d), d.shape)
        bad = np.argwhere(d > 1e-6 * np.nanmax(np.abs(ref[k][np.isfinite(ref[k])])))
        print(k, "max diff", d[i], "at", i, "got", got[k][i], "ref", ref[k][i], "nbad", len(bad), "bad steps", sorted(set(bad[:, 1]))[:12])
    print("nlZ", [g for g in got["nlZ"]], ref["nlZ"])
    print("mismatch", [p.adf_mismatch() for p in plans])
