#!/usr/bin/env python
"""One gf_ep_modulator_nmf predict run of BASELINE config C3's shape (D=16 exp subbands, N=3 matern52 modulators,
n=41, sqrt-model likelihood p=9, ep_itts=3) at T = argv[1] (default 100000): the workload for ncu captures of the
full-state sequential pass.  Prints the phase timings."""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nsagp = importlib.import_module("nonstationary-audio-gp_b200")
L = nsagp._lib

T = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
rng = np.random.default_rng(7)
D, N = 16, 3
hyp = nsagp.synth.speech_hypers(D, N, rng)
y, _, _ = nsagp.synth.sample_signal(hyp, "exp", "matern52", T, rng, link_shift=1.0, sqrt_model=True)
F, Lm, Qc, H, Pinf = nsagp.ss_modulators_nmf(hyp.w_sub(), hyp.w_mod(), "exp", "matern52")[:5]
A, Q = nsagp.lti_disc(F, Lm, Qc, 1.0)
mdl = nsagp.to_block_model(A, Q, H, Pinf, D, N)
wn, xn = nsagp.utp_ws(9, N)
mom = nsagp.likModulatorPreCalcwn(nsagp.Softplus(1.0), wn, xn)
with nsagp.Plan(L.KIND_FULL, [mdl], [(mom, np.log([hyp.w_lik]), hyp.W)], 0.75, np.linspace(0.05, 0.1, 3), 3, y[None, :],
                L.MODE_PREDICT) as p:
    p.run()
    p.run()
    print(json.dumps(dict(T=T, phases_ms=p.timings())))
