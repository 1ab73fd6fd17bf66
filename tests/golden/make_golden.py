#!/usr/bin/env python
"""Generate the committed golden fixtures (tests/golden/*.npz) with the oracle.

The reference ships no golden vectors for this path and cannot be executed here
(MATLAB), so these vectors pin the *oracle* (regression guard, and the second
restatement in oracle/c is checked against them) and give the GPU tests a
committed target.  Inputs are seeded NumPy draws (seed stored in each file).

    python tests/golden/make_golden.py
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = {
    # name: (entry, D, N, T, k1, k2, kind, p, shift, alpha, itts, damping, gaps, seed)
    "gfep_demo_small": ("gf_ep", 4, 2, 240, "matern32", "matern52", "power", 9, 0.0, 0.5, 3, (0.5, 0.5, 0.5), False, 100),
    "gfep_c3_small": ("gf_ep", 6, 3, 240, "exp", "matern52", "precalc", 9, 1.0, 0.75, 3, (0.3, 0.2, 0.1), True, 101),
    "ihgp_demo_small": ("ihgp", 4, 2, 240, "matern32", "matern52", "power", 9, 0.0, 0.5, 3, (0.5, 0.5, 0.5), False, 102),
    "ihgp_c2_small": ("ihgp", 6, 3, 240, "exp", "matern52", "precalc", 9, 1.0, 0.75, 4, (0.1, 0.1, 0.1, 0.1), True, 103),
}


def build(name):
    from conftest import make_problem
    from oracle import gf_ep, ihgp_ep
    nsagp = importlib.import_module("nonstationary-audio-gp_b200")
    entry, D, N, T, k1, k2, kind, p, shift, alpha, itts, damping, gaps, seed = CASES[name]
    pb = make_problem(nsagp, D, N, T, k1, k2, seed=seed, kind=kind, p=p, shift=shift, gaps=gaps)
    fn = gf_ep.gf_ep_modulator_nmf if entry == "gf_ep" else ihgp_ep.ihgp_ep_modulator_nmf
    args = (pb["w"], pb["t"], pb["y"], pb["ss_ref"], pb["mom_ref"], pb["t"], k1, k2, 1, D, N, alpha, np.array(damping), itts)
    Eft, Varft, _, lb, ub, out = fn(*args)
    nlz, _ = fn(*(args[:5] + (None,) + args[6:]))
    return dict(entry=entry, D=D, N=N, T=T, kernel1=k1, kernel2=k2, kind=kind, p=p, shift=shift, alpha=alpha,
                itts=itts, damping=np.array(damping), seed=seed, w=pb["w"], y=pb["y"], t=pb["t"],
                Eft=Eft, Varft=Varft, lb=lb, ub=ub, ttau=out["ttau"], tnu=out["tnu"], R=out["R"], MS=out["MS"],
                nlZ=out["nlZ"], nlz_mode=nlz, n_negcav=out["n_negcav"])


if __name__ == "__main__":
    for name in CASES:
        d = build(name)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, "nlZ", d["nlZ"], "nlz_mode", d["nlz_mode"])
