"""Per-entry error of the GPU gradient against the oracle (debug)."""
import importlib, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests"))
import numpy as np
nsagp = importlib.import_module("nonstationary-audio-gp_b200")
from conftest import make_problem
from oracle import giekf, ssmodel as oss
ss_ref = lambda x, p1, p2, k1, k2: oss.ss_modulators_nmf(p1, p2, k1, k2) + oss.ss_modulators_nmf_derivs(p1, p2, k1, k2)
for D, N, T, k1, k2 in [(4, 2, 120, "matern32", "matern52"), (6, 3, 150, "exp", "matern52"), (3, 2, 80, "matern72", "exp"), (5, 2, 100, "matern52", "matern32"), (32, 3, 60, "exp", "matern52")]:
    pb = make_problem(nsagp, D, N, T, k1, k2, seed=40 + D, w_lik=1e-2)
    for bal in (False, True):
        eo, go = giekf.gf_giekf_modulator_nmf(pb["w"], pb["t"], pb["y"], ss_ref, None, None, k1, k2, 1, D, N, 1, 1, GradObj="on", balance_derivatives=bal)
        eg, gg = nsagp.gf_giekf_modulator_nmf(pb["w"], pb["t"], pb["y"], pb["ss_gpu"], None, None, k1, k2, 1, D, N, 1, 1, GradObj="on", balance_derivatives=bal)
        rel = np.abs(gg - go) / np.maximum(np.abs(go), 1e-300)
        med = np.median(np.abs(go))
        relf = np.abs(gg - go) / np.maximum(np.abs(go), 1e-3 * med)
        print(D, N, k1, k2, bal, "max|g| %.3g med %.3g min %.3g | max rel %.3g (at |g|=%.3g) max floored rel %.3g" % (np.abs(go).max(), med, np.abs(go).min(), rel.max(), np.abs(go)[rel.argmax()], relf.max()))
