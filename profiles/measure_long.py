#!/usr/bin/env python
"""IHGP predict run of the bench model (C2) on a LONG signal (default T = 2 000 000, ep_itts = 3): per-phase device
times and the achieved bandwidth of the frozen-site passes against their algorithmic bytes (8 + 24 n + 88 M per step
and sweep) -- one 100 000-step signal exposes only M*T/32 = 59 k scan threads and leaves those passes
latency-bound; a long signal shows what they do when the machine is full."""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

nsagp = importlib.import_module(bench.PKG)
L = nsagp._lib
T = int(sys.argv[1]) if len(sys.argv) > 1 else 2000000
itts = 3
hyp, y = bench.make_signal(nsagp, 0, T)
mdl, tabs = bench.host_setup(nsagp, hyp)
wn, xn = nsagp.utp_ws(bench.P_CUB, bench.N)
mom = nsagp.likModulatorPreCalcwn(nsagp.Softplus(bench.SHIFT), wn, xn)
with nsagp.Plan(L.KIND_IHGP, [mdl], [(mom, np.log([hyp.w_lik]), hyp.W)], bench.ALPHA, bench.damping(itts), itts, y[None, :],
                L.MODE_PREDICT, tables=[tabs]) as plan:
    plan.run()
    plan.run()
    ph = plan.timings()
n, M = mdl.n, mdl.M
bytes_per_step_sweep = 8 + 24 * n + 88 * M
frozen_ms = ph["fixed_filter"] + ph["smoother"]
sweeps = (itts - 1) + itts                      # frozen filter passes + smoother passes
gbps = bytes_per_step_sweep * T * sweeps / 2 / (frozen_ms * 1e-3) / 1e9     # each pass is half a filter+smoother sweep
print(json.dumps(dict(T=T, ep_itts=itts, phases_ms=ph, frozen_pass_GBps=gbps, frozen_pass_frac_of_6455=gbps / 6454.9,
                      adf_steps_per_s=T / (ph["adf"] * 1e-3), site_update_steps_per_s=T * (itts - 1) / (ph["site_update"] * 1e-3))))
