"""Time of the EKF energy + analytic gradient (csrc/ekfgrad.cuh) at C4's shape, and of the energy-only pass.
    python profiles/debug/giekf_grad_time.py [T]"""
import importlib, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests"))
import numpy as np
nsagp = importlib.import_module("nonstationary-audio-gp_b200")
from conftest import make_problem
T = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
L = nsagp._lib
for D, N, k1, k2 in ((32, 3, "exp", "matern52"), (16, 3, "exp", "matern52"), (16, 3, "matern32", "matern52")):
    pb = make_problem(nsagp, D, N, T, k1, k2, seed=5, w_lik=1e-2)
    for rep in range(2):
        t0 = time.time()
        e, g = nsagp.gf_giekf_modulator_nmf(pb["w"], pb["t"], pb["y"], pb["ss_gpu"], None, None, k1, k2, 1, D, N, 1, 1, GradObj="on")
        wall = time.time() - t0
        ms = (np.zeros(2))
        L.check(L.lib().nsagp_giekf_timings(L.dptr(ms), 2))
    t0 = time.time()
    e0, _ = nsagp.gf_giekf_modulator_nmf(pb["w"], pb["t"], pb["y"], pb["ss_gpu"], None, None, k1, k2, 1, D, N, 1, 1)
    ms0 = np.zeros(2); L.check(L.lib().nsagp_giekf_timings(L.dptr(ms0), 2))
    print("D=%d N=%d %s/%s n=%d nparam=%d T=%d: grad kernel %.1f ms (%.2f us/step, wall %.2f s), energy only %.1f ms; e %.6f vs %.6f, |g|max %.3g"
          % (D, N, k1, k2, len(pb["w"]) and 0 or 0, g.size, T, ms[0], 1e3 * ms[0] / T, wall, ms0[0], e, e0, np.abs(g).max()))
