#!/usr/bin/env python
"""A/B of the round-2 (third part) changes on one box, same library: frozen-site scans with the two latent families in one
CTA tile (nsagp_scan_merge 1) against one launch per family (0), padded single-size scans against family-specialised ones
(nsagp_scan_config), and the site update with the sigma points two at a time (nsagp_site_config 0) against one at a time
(2).  IHGP model of the bench (C2) at T = 1e5 and 2e6, the full-state path (C3) at 5e5; the first pass runs in its
parallel form on the long signals so that the run is short (the phases compared here do not depend on it)."""
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
lib = L.lib()
itts = 4
wn, xn = nsagp.utp_ws(bench.P_CUB, bench.N)
mom = nsagp.likModulatorPreCalcwn(nsagp.Softplus(bench.SHIFT), wn, xn)
CONFIGS = [  # name, merge, family_min_steps, site form, L2 prefetch, chunks per tile, tile threads, (unused)
    ("r2i: launch per family >= 400k, site 1pt", 0, 400000, 2, 0, 16, 320, 0),
    ("one tile, site 2pt, L2 prefetch, 320 threads", 1, 0, 0, 1, 16, 320, 0),
    ("  same, 256 threads (128 registers): default", 1, 0, 0, 1, 16, 256, 0),
    ("  same, 256 threads, 8 chunks", 1, 0, 0, 1, 8, 256, 0),
]
for kind, T in ((L.KIND_IHGP, 100000), (L.KIND_IHGP, 2000000), (L.KIND_IHGP, 10000000), (L.KIND_FULL, 500000)):
    hyp, y = bench.make_signal(nsagp, 0, T)
    mdl, tabs = bench.host_setup(nsagp, hyp)
    for name, merge, fam, form, pf, ch, th, el in CONFIGS:
        L.check(lib.nsagp_scan_merge(merge)); L.check(lib.nsagp_scan_config(fam)); L.check(lib.nsagp_site_config(form))
        L.check(lib.nsagp_scan_prefetch(pf)); L.check(lib.nsagp_scan_tile(ch, th))
        with nsagp.Plan(kind, [mdl], [(mom, np.log([hyp.w_lik]), hyp.W)], bench.ALPHA, bench.damping(itts), itts, y[None, :],
                        L.MODE_PREDICT, tables=[tabs] if kind == L.KIND_IHGP else None) as plan:
            if T > 100000:
                plan.set_adf_parallel(148, 60000)
            plan.run()
            best = None
            for _ in range(3):
                plan.run()
                ph = plan.timings()
                if best is None or ph["total"] - ph["adf"] < best["total"] - best["adf"]:
                    best = ph
            nlZ = float(plan.fetch(names=("nlZ",))["nlZ"][-1])
        per = lambda k, cnt: best[k] / cnt * 1e6 / T           # ms per 1e6 steps and pass
        print(json.dumps(dict(kind="ihgp" if kind == L.KIND_IHGP else "full", T=T, config=name,
                              filter_ms_per_1e6=per("fixed_filter", itts - 1), smoother_ms_per_1e6=per("smoother", itts),
                              site_ms_per_1e6=per("site_update", itts - 1), nlZ=nlZ)), flush=True)
L.check(lib.nsagp_scan_merge(1)); L.check(lib.nsagp_scan_config(0)); L.check(lib.nsagp_site_config(0)); L.check(lib.nsagp_scan_prefetch(1)); L.check(lib.nsagp_scan_tile(16, 256))
