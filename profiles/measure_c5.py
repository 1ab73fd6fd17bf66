#!/usr/bin/env python
"""C5 (256 problems x 39 062 steps, nlZ mode) on one GPU with the sequential pass as one full-width CTA per problem
(form 0: 148 at a time, two waves) and as half-width CTAs, two per SM (form 2: one wave); and a 512-problem batch."""
import importlib
import json
import os
import sys

os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench_workloads as bw  # noqa: E402

nsagp = importlib.import_module("nonstationary-audio-gp_b200")
ctx = bw.Ctx(nsagp, torch, None, 0, 1, 0)
for B in (256, 512, 148):
    for form in (0, 2):
        r = bw.c5_batch(ctx, B=B, adf_form=form)
        print(json.dumps(dict(B=B, adf_form=form, ihgp=r["ihgp_nlZ_ep_itts1"], gf_ep=r["gf_ep_nlZ_ep_itts3"])), flush=True)
