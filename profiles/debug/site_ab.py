import importlib, os, sys
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import make_problem, rel_err
nsagp = importlib.import_module("nonstationary-audio-gp_b200")
nsagp.tables.DEFAULT_NATIVE = True
D, N, k1, k2, T = 6, 2, sys.argv[2], sys.argv[3], 150
pb = make_problem(nsagp, D, N, T, k1, k2, seed=1000 + 31 * D + N, kind="power", p=5, gaps=(D % 3 == 0))
damping = [0.5, 0.4, 0.3]
args = (pb["w"], pb["t"], pb["y"], pb["ss_gpu"], pb["mom_gpu"], pb["t"], k1, k2, 1, D, N, 0.5, damping, 3)
out = {}
for name, entry in (("ihgp", nsagp.ihgp_ep_modulator_nmf), ("gfep", nsagp.gf_ep_modulator_nmf)):
    for f in (0, 1):
        E, V, _, _, _, o = entry(*args, adf_form=f)
        out["%s_E%d" % (name, f)] = E; out["%s_V%d" % (name, f)] = V; out["%s_nlZ%d" % (name, f)] = o["nlZ"]; out["%s_tt%d" % (name, f)] = o["ttau"]
        out["%s_neg%d" % (name, f)] = o["n_negcav"]
np.savez(sys.argv[1], **out)
