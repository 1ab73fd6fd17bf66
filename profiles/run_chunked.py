#!/usr/bin/env python
"""One C2-shaped signal time-chunked over the GPUs of a box (torchrun, NCCL): checks the result
against a single-GPU run on rank 0 and reports wall times.
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/run_chunked.py [T] [ep_itts]
"""
import importlib
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nsagp = importlib.import_module("nonstationary-audio-gp_b200")
L = nsagp._lib


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    itts = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    os.environ.setdefault("NCCL_DEBUG", "ERROR")
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    L.check(L.lib().nsagp_set_device(lr))
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    comm = nsagp.chunked.TorchComm()
    D, N = 16, 3
    rng = np.random.default_rng(2026)
    hyp = nsagp.synth.speech_hypers(D, N, rng)
    y, _, _ = nsagp.synth.sample_signal(hyp, "exp", "matern52", T, rng, link_shift=1.0, sqrt_model=True)
    F, Lm, Qc, H, Pinf = nsagp.ss_modulators_nmf(hyp.w_sub(), hyp.w_mod(), "exp", "matern52")[:5]
    F, Lm, H, Pinf = nsagp.ssmodel.balance(F, Lm, H, Pinf)
    A, Q = nsagp.lti_disc(F, Lm, Qc, 1.0)
    Q = (Q + Q.T) / 2
    mdl = nsagp.to_block_model(A, Q, H, Pinf, D, N)
    tabs = nsagp.tables.build_tables(mdl, want_smoother=True)
    wn, xn = nsagp.utp_ws(9, N)
    mom = nsagp.likModulatorPreCalcwn(nsagp.Softplus(1.0), wn, xn)
    damp = np.linspace(0.01, 0.1, itts)
    mk = lambda: nsagp.Plan(L.KIND_IHGP, [mdl], [(mom, np.log([hyp.w_lik]), hyp.W)], 0.75, damp, itts, y[None, :],
                            L.MODE_PREDICT, tables=[tabs])
    plan = mk()
    names = ("Eft", "ttau", "MS", "nlZ")
    times = []
    for rep in range(3):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        ranges = nsagp.chunked.run_ihgp_chunked(plan, comm, damp)
        torch.cuda.synchronize(); dist.barrier()
        times.append(time.perf_counter() - t0)
    got = nsagp.chunked.gather_outputs(plan, comm, ranges, names)
    if comm.rank == 0:
        ref_plan = mk()
        ref_plan.run()
        t0 = time.perf_counter(); ref_plan.run(); t_single = time.perf_counter() - t0
        ref = ref_plan.fetch(0, names)
        err = {k: float(np.max(np.abs(got[k] - ref[k])) / np.max(np.abs(ref[k]))) for k in names}
        print(json.dumps({"world": comm.world, "T": T, "ep_itts": itts, "chunked_s": min(times), "single_gpu_s": t_single,
                          "rel_err_vs_single": err, "steps_per_s_chunked": T * itts / min(times),
                          "note": "the ADF pass (first filter pass) is replicated on every rank; only the frozen-site passes shard"}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
