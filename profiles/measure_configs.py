#!/usr/bin/env python
"""Device-timed throughput of the other BASELINE.json configurations (they are parity-test
cases, not bench.py lines): C1 demo, C3 full-state EP at T=500k, C5 batched nlZ over 256 clips.
Prints one JSON line per configuration.  Run on the GPU box:  python profiles/measure_configs.py
"""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nsagp = importlib.import_module("nonstationary-audio-gp_b200")
L = nsagp._lib


def model(hyp, k1, k2, D, N, balance, ihgp, smoother):
    F, Lm, Qc, H, Pinf = nsagp.ss_modulators_nmf(hyp.w_sub(), hyp.w_mod(), k1, k2)[:5]
    if balance:
        F, Lm, H, Pinf = nsagp.ssmodel.balance(F, Lm, H, Pinf)
    A, Q = nsagp.lti_disc(F, Lm, Qc, 1.0)
    if ihgp:
        Q = (Q + Q.T) / 2
    mdl = nsagp.to_block_model(A, Q, H, Pinf, D, N)
    tabs = nsagp.tables.build_tables(mdl, want_smoother=smoother) if ihgp else None
    return mdl, tabs


def timed(plan, reps=3):
    plan.run()
    ms = []
    for _ in range(reps):
        plan.run()
        ms.append(plan.timings()["total"])
    return float(np.median(ms)), plan.timings()


def main():
    out = []
    # ---- C1: demo_toy_modulators_nmf, gf_ep predict, D=10 N=2 matern32/matern52, T=5000, ep_itts=3
    rng = np.random.default_rng(100)
    D, N, T = 10, 2, 5000
    hyp = nsagp.synth.demo_hypers(D, N, rng)
    y, _, _ = nsagp.synth.sample_signal(hyp, "matern32", "matern52", T, rng)
    mdl, _ = model(hyp, "matern32", "matern52", D, N, False, False, False)
    mom = nsagp.likModulatorNMFPower(nsagp.Softplus(0.0), 9, N)
    with nsagp.Plan(L.KIND_FULL, [mdl], [(mom, np.log([hyp.w_lik]), hyp.W)], 0.5, [0.5] * 3, 3, y[None, :], L.MODE_PREDICT) as p:
        ms, ph = timed(p)
    out.append(dict(config="C1 gf_ep_modulator_nmf predict D=10 N=2 n=46 T=5000 ep_itts=3", ms=ms,
                    steps_per_s=T * 3 / ms * 1e3, phases_ms=ph))
    # ---- C3: gf_ep predict, D=16 N=3 exp/matern52 (n=41), T=500000, ep_itts=3
    rng = np.random.default_rng(7)
    D, N, T = 16, 3, 500000
    hyp = nsagp.synth.speech_hypers(D, N, rng)
    y, _, _ = nsagp.synth.sample_signal(hyp, "exp", "matern52", T, rng, link_shift=1.0, sqrt_model=True)
    mdl, _ = model(hyp, "exp", "matern52", D, N, False, False, False)
    wn, xn = nsagp.utp_ws(9, N)
    mom = nsagp.likModulatorPreCalcwn(nsagp.Softplus(1.0), wn, xn)
    damp = np.linspace(0.05, 0.1, 3)
    with nsagp.Plan(L.KIND_FULL, [mdl], [(mom, np.log([hyp.w_lik]), hyp.W)], 0.75, damp, 3, y[None, :], L.MODE_PREDICT) as p:
        ms, ph = timed(p, reps=2)
    out.append(dict(config="C3 gf_ep_modulator_nmf predict D=16 N=3 n=41 T=500000 ep_itts=3", ms=ms,
                    steps_per_s=T * 3 / ms * 1e3, phases_ms=ph))
    # ---- C5: 256 clips x nlZ mode (what fminunc evaluates), T=39062 each (10 M steps), ihgp and gf_ep
    B, T = 256, 39062
    rng = np.random.default_rng(11)
    mdls, liks, tabs, ys = [], [], [], []
    t0 = time.perf_counter()
    base = nsagp.synth.speech_hypers(D, N, rng)
    mdl_i, tab_i = model(base, "exp", "matern52", D, N, True, True, False)
    mdl_f, _ = model(base, "exp", "matern52", D, N, False, False, False)
    for b in range(B):
        yb, _, _ = nsagp.synth.sample_signal(base, "exp", "matern52", T, np.random.default_rng(1000 + b), link_shift=1.0,
                                             sqrt_model=True)
        ys.append(yb)
    ys = np.stack(ys)
    gen_s = time.perf_counter() - t0
    lik = (mom, np.log([base.w_lik]), base.W)
    for form in (0, 1):
        with nsagp.Plan(L.KIND_IHGP, [mdl_i] * B, [lik] * B, 0.75, [0.1], 1, ys, L.MODE_NLZ, tables=[tab_i] * B) as p:
            p.set_adf_form(form)
            ms, ph = timed(p, reps=2)
        out.append(dict(config="C5 ihgp nlZ mode, 256 clips x T=39062 (10 M steps), adf_form=%d" % form, ms=ms,
                        steps_per_s=B * T / ms * 1e3, phases_ms=ph, signal_generation_s=gen_s))
    for form in (0, 1):
        with nsagp.Plan(L.KIND_FULL, [mdl_f] * B, [lik] * B, 0.75, [0.1], 1, ys, L.MODE_NLZ) as p:
            p.set_adf_form(form)
            ms, ph = timed(p, reps=2)
        out.append(dict(config="C5 gf_ep nlZ mode ep_itts=1, 256 clips x T=39062 (10 M steps), adf_form=%d" % form, ms=ms,
                        steps_per_s=B * T / ms * 1e3, phases_ms=ph))
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
