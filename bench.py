#!/usr/bin/env python
"""Benchmark of the EP-in-Kalman hot path (BASELINE.json metric: EP filter+smoother
time-steps/sec, FP64).

Workload (BASELINE.json configs[1], "C2"): ihgp_ep_modulator_nmf, predict mode,
D = 16 subbands (exp kernel) x N = 3 NMF modulators (matern52), state dim 41,
likModulatorPreCalcwn with the 9th-order symmetric rule (77 sigma points), link
softplus(g-1), alpha = 0.75, ep_itts = 20, damping linspace(0.01, 0.1, 20),
T = 100 000 synthetic "speech-shaped" samples, one signal per GPU.

A bench "step" is one complete EP run over the signal: ep_itts filter+smoother
sweeps.  time-steps/sec = T * ep_itts * steps / time  (sweeps =
ep_itts in predict mode).  Host-side setup the reference also does once per call
with MATLAB built-ins (ss, balance, lti_disc, 1216 DAREs) is outside the timed
region and reported as setup_s.

  value : device-timed, model/tables/signal already resident in HBM
  e2e   : the same run through the C ABI call nsagp_ep_ihgp with HOST buffers
          (H2D of signal+model+tables, D2H of Eft/Varft/lb/ub inside the timed region)
  --impl reference : the reference's MATLAB loop restated in plain C (oracle/c, same
          dense operations; the reference itself cannot run: no MATLAB/Octave) on all
          host cores, one independent clip per core, bounded sample.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")      # before torch initialises CUDA (see _lib.py)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "nonstationary-audio-gp_b200"

D, N, T_FULL = 16, 3, 100000
K1, K2 = "exp", "matern52"
ALPHA, SHIFT, P_CUB = 0.75, 1.0, 9
EP_ITTS = 20
# DRAM bytes per time step of the ADF kernel from the committed ncu --set full capture (profiles/); None = not captured
TRAFFIC_ADF_BYTES_PER_STEP = (32.034560e6 + 24.303104e6) / 100000     # profiles/r2x_adf_full.md (T = 100000; part of the writes was still in L2 when the kernel ended)
WORKLOAD = ("C2: ihgp_ep_modulator_nmf predict mode, D=16 exp subbands x N=3 matern52 modulators (n=41), "
            "likModulatorPreCalcwn p=9 (S=77), alpha=0.75, ep_itts=20, T=100000, 1 signal per GPU")


def damping(itts):
    return np.linspace(0.01, 0.1, itts)


def config_dict(T, itts):
    """The `config` object of the JSON line -- identical for the GPU arm and the reference arm."""
    return {"workload": WORKLOAD if (T == T_FULL and itts == EP_ITTS) else "C2 shape with T=%d ep_itts=%d (non-default)" % (T, itts),
            "T": T, "ep_itts": itts, "signals_per_gpu": 1, "l2": "flushed between timed runs (256 MiB write)",
            "step": "one full EP run = ep_itts filter+smoother sweeps"}


def make_signal(nsagp, seed, T):
    rng = np.random.default_rng(seed)
    hyp = nsagp.synth.speech_hypers(D, N, rng)
    y, _, _ = nsagp.synth.sample_signal(hyp, K1, K2, T, rng, link_shift=SHIFT, sqrt_model=True)
    return hyp, y


def host_setup(nsagp, hyp):
    """What the .m wrapper keeps in MATLAB: ss, balance, lti_disc, DARE tables."""
    F, L, Qc, H, Pinf = nsagp.ss_modulators_nmf(hyp.w_sub(), hyp.w_mod(), K1, K2)[:5]
    F, L, H, Pinf = nsagp.ssmodel.balance(F, L, H, Pinf)
    A, Q = nsagp.lti_disc(F, L, Qc, 1.0)
    Q = (Q + Q.T) / 2
    mdl = nsagp.to_block_model(A, Q, H, Pinf, D, N)
    tabs = nsagp.tables.build_tables(mdl, want_smoother=True)
    return mdl, tabs


def prefix_check(nsagp, seed=2026, Tp=2000, itts=EP_ITTS, hyp=None, y=None):
    """The metric's second half ("lZ rel. error"): the product path against the C restatement of the reference
    (oracle/c/nsagp_oracle.c, the checker -- never the thing measured) on the first Tp samples of the benched signal,
    same model, same EP schedule.  Runs OUTSIDE every timed region.  Kernel parity is measured on the oracle's own
    steady-state tables (SciPy Riccati solver); `*_native_tables` additionally swaps in the library's own table
    routine (what the timed runs use), whose solutions differ from SciPy's by ~1e-10 and can flip a
    nearest-neighbour table look-up."""
    from oracle import c_oracle, cubature as ocub, ihgp_ep, ssmodel as oss
    lib_mod = nsagp._lib
    if hyp is None:
        hyp, y = make_signal(nsagp, seed, T_FULL)
    yp = np.ascontiguousarray(y[:Tp])
    damp = damping(itts)
    lik_param, p1, p2, W = oss.unpack_log(hyp.pack_log(), 1, D, N)
    ss = lambda x, a, b, k1, k2: oss.ss_modulators_nmf(a, b, k1, k2)
    A, Q, H, Pinf = ihgp_ep._model(lik_param, p1, p2, ss, np.arange(1.0, Tp + 1), K1, K2)
    wo, xo = ocub.utp_ws(P_CUB, N)
    ref = c_oracle.IhgpProblem(A, H, Pinf, ihgp_ep.ihgp_setup(A, Q, H), 1, lik_param, SHIFT, W, wo, xo, ALPHA, damp,
                               itts).predict(yp)
    F, L_, Qc, Hh, Pi = nsagp.ss_modulators_nmf(hyp.w_sub(), hyp.w_mod(), K1, K2)[:5]
    F, L_, Hh, Pi = nsagp.ssmodel.balance(F, L_, Hh, Pi)
    Ad, Qd = nsagp.lti_disc(F, L_, Qc, 1.0)
    mdl = nsagp.to_block_model(Ad, (Qd + Qd.T) / 2, Hh, Pi, D, N)
    wn, xn = nsagp.utp_ws(P_CUB, N)
    mom = nsagp.likModulatorPreCalcwn(nsagp.Softplus(SHIFT), wn, xn)
    rel = lambda a, b: float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / np.max(np.abs(b)))
    out = {"prefix_steps": Tp, "ep_itts": itts, "oracle": "oracle/c/nsagp_oracle.c (plain-C restatement; parity unpinned: "
           "the reference is MATLAB and ships no golden vectors)"}
    for native in (False, True):
        tabs = nsagp.tables.build_tables(mdl, want_smoother=True, native=native)
        with nsagp.Plan(lib_mod.KIND_IHGP, [mdl], [(mom, np.log([hyp.w_lik]), hyp.W)], ALPHA, damp, itts, yp[None, :],
                        lib_mod.MODE_PREDICT, tables=[tabs]) as plan:
            plan.run()
            got = plan.fetch(0, ("Eft", "nlZ", "ttau", "MS"))
        sfx = "_native_tables" if native else ""
        out["lZ_rel_err" + sfx] = float(np.max(np.abs(got["nlZ"] - ref["nlZ"]) / np.abs(ref["nlZ"])))
        out["Eft_rel_err" + sfx] = rel(got["Eft"], ref["Eft"])
        if not native:
            out["lZ_rel_err_first_sweep"] = float(abs(got["nlZ"][0] - ref["nlZ"][0]) / abs(ref["nlZ"][0]))
            out["ttau_rel_err"] = rel(got["ttau"], ref["ttau"])
            out["MS_rel_err"] = rel(got["MS"], ref["MS"])
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------- CPU arms
def _oracle_clip(args):
    """One independent clip through the CPU port (runs in a worker process).  The port is
    oracle/c/nsagp_oracle.c: the reference's MATLAB loop restated in plain C with the same dense
    n-by-n operations (the reference itself cannot run: no MATLAB/Octave here or on the GPU box).
    Model construction and the DARE tables are excluded from the timed region, as on the GPU side."""
    seed, T, itts = args
    nsagp = importlib.import_module(PKG)
    from oracle import c_oracle, cubature as ocub, ihgp_ep, ssmodel as oss
    hyp, y = make_signal(nsagp, seed, T)
    wo, xo = ocub.utp_ws(P_CUB, N)
    ss = lambda x, p1, p2, k1, k2: oss.ss_modulators_nmf(p1, p2, k1, k2)
    t = np.arange(1.0, T + 1.0)
    lik_param, param1, param2, Wnmf = oss.unpack_log(hyp.pack_log(), 1, D, N)
    A, Q, H, Pinf = ihgp_ep._model(lik_param, param1, param2, ss, t, K1, K2)
    tabs = ihgp_ep.ihgp_setup(A, Q, H)
    prob = c_oracle.IhgpProblem(A, H, Pinf, tabs, 1, lik_param, SHIFT, Wnmf, wo, xo, ALPHA, damping(itts), itts)
    t0 = time.perf_counter()
    prob.predict(y)
    return time.perf_counter() - t0


def cpu_arm(steps, warmup, T_sample, itts, cores):
    """CPU port on `cores` host cores, one independent clip per core per step.  Returns
    (time-steps/s over the timed region of the slowest core, seconds per step)."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    times = []
    with ctx.Pool(cores) as pool:
        for it in range(warmup + steps):
            per_clip = pool.map(_oracle_clip, [(1000 + 17 * it + c, T_sample, itts) for c in range(cores)])
            if it >= warmup:
                times.append(max(per_clip))
    per_step = float(np.mean(times))
    return cores * T_sample * itts / per_step, per_step


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # The reference's loop is serial in time (no parfor; its n = 41 matrices are too small for BLAS
    # threads), so ONE signal can use one host core.  "All the host threads it can use" is therefore one independent
    # signal per core; the value is the box's aggregate throughput over those signals, per-core figure in `sample`.
    cores = max(1, os.cpu_count() or 1)
    T_sample, itts = args.ref_T, args.ep_itts
    value, per_step = cpu_arm(args.steps, args.warmup, T_sample, itts, cores)
    sample = ("bounded sample of the workload: %d independent signals (one per host core; the reference loop is serial per "
              "signal), the first T=%d of the workload's %d samples each, the workload's full EP schedule (ep_itts=%d) per step; "
              "plain-C port of matlab/ihgp_ep_modulator_nmf.m (oracle/c/nsagp_oracle.c, same dense operations; the reference is "
              "MATLAB and neither MATLAB nor Octave exists here or on the GPU box); per-step cost is independent of T; "
              "%.0f time-steps/s per core; model/table setup untimed as on the GPU side" % (cores, T_sample, args.T, itts, value / cores))
    line = {"impl": "reference", "metric": "EP filter+smoother time-steps/sec (FP64)", "value": value,
            "unit": "time-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config_dict(args.T, itts),
            "cpu_baseline": {"value": value, "unit": "time-steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "time-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import ctypes as C

    import torch
    nsagp = importlib.import_module(PKG)
    lib_mod = nsagp._lib
    L = lib_mod.lib()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    lib_mod.check(L.nsagp_set_device(local_rank))
    if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
        os.environ["NCCL_DEBUG"] = "ERROR"          # keep stdout to the one JSON line (NCCL prints its version there)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.current_stream()
    lib_mod.check(L.nsagp_set_stream(C.c_void_p(stream.cuda_stream)))

    T, itts = args.T, args.ep_itts
    t0 = time.perf_counter()
    hyp, y = make_signal(nsagp, 2026 + rank, T)            # one independent clip per rank (weak scaling)
    gen_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    mdl, tabs = host_setup(nsagp, hyp)
    setup_s = time.perf_counter() - t0
    wn, xn = nsagp.utp_ws(P_CUB, N)
    mom = nsagp.likModulatorPreCalcwn(nsagp.Softplus(SHIFT), wn, xn)
    lik_param = np.log([hyp.w_lik])
    damp = damping(itts)

    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")   # 256 MiB > 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm -------------------------------------------------
    plan = nsagp.Plan(lib_mod.KIND_IHGP, [mdl], [(mom, lik_param, hyp.W)], ALPHA, damp, itts, y[None, :],
                      lib_mod.MODE_PREDICT, tables=[tabs])
    for _ in range(args.warmup):
        plan.run()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    L.nsagp_launch_count(1)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    phase = {"adf": 0.0, "fixed_filter": 0.0, "smoother": 0.0, "site_update": 0.0, "total": 0.0}
    wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()                                       # evict the previous run's arrays from L2
        ev[i][0].record(stream)
        plan.run()
        ev[i][1].record(stream)
        tm = plan.timings()
        for k in phase:
            phase[k] += tm[k]
    barrier()
    wall = time.perf_counter() - wall0
    launches = int(L.nsagp_launch_count(0))
    clocks = sampler.stop()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    chk = plan.fetch(0, ("nlZ", "n_negcav"))
    plan.close()

    # ---- end-to-end arm: C ABI call with host buffers ------------------------
    Tm = (T, D + N)
    pin = lambda shape: torch.empty(shape, dtype=torch.float64).pin_memory().numpy()
    y_pin = pin((T,)); y_pin[:] = y
    outs = {k: pin(Tm) for k in ("Eft", "Varft", "lb", "ub")}
    keep = []
    cm = lib_mod.Model()
    arrs = [lib_mod.as_f64(a) for a in (mdl.A, mdl.Q, mdl.Pinf, mdl.h)]
    cm.D, cm.N, cm.bz, cm.bg = mdl.D, mdl.N, mdl.bz, mdl.bg
    cm.A, cm.Q, cm.Pinf, cm.h = [lib_mod.dptr(a) for a in arrs]
    cl = mom.c_lik(lik_param, hyp.W, keep)
    pp, pg = tabs.packed()
    r = lib_mod.as_f64(tabs.r); pp = lib_mod.as_f64(pp); pg = lib_mod.as_f64(pg)
    ct = lib_mod.Tables(r.size, lib_mod.dptr(r), lib_mod.dptr(pp), lib_mod.dptr(pg))
    dmp = lib_mod.as_f64(damp)
    ep = lib_mod.Ep(ALPHA, lib_mod.dptr(dmp), itts)
    co = lib_mod.Outputs()
    for k, a in outs.items():
        setattr(co, k, lib_mod.dptr(a))
    h2d = y_pin.nbytes + sum(a.nbytes for a in arrs) + r.nbytes + pp.nbytes + pg.nbytes + hyp.W.nbytes + wn.nbytes + xn.nbytes
    d2h = sum(a.nbytes for a in outs.values())

    def e2e_call():
        lib_mod.check(L.nsagp_ep_ihgp(C.byref(cm), C.byref(cl), C.byref(ep), C.byref(ct), lib_mod.dptr(y_pin), T,
                                      lib_mod.MODE_PREDICT, C.byref(co)))
    for _ in range(min(args.warmup, 2)):
        e2e_call()
    barrier()
    e0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_call()
    barrier()
    e2e_s = time.perf_counter() - e0

    # ---- reductions over ranks (max time) ------------------------------------
    vals = torch.tensor([dev_ms, e2e_s, wall], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    dev_ms, e2e_s, wall = [float(v) for v in vals.cpu()]
    units = world * T * itts * args.steps
    value = units / (dev_ms * 1e-3)
    e2e_value = units / e2e_s

    line = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
        n, M = mdl.n, mdl.M
        # Dominant kernel: ihgp_adf_cta_kernel, the sequential ADF filter pass (one launch per run).
        # Algorithmic bytes per time step of that pass: read y (8) and the old
        # sites ttau,tnu (16M); write ttau,tnu,R (24M), the filtered mean (8n) and lZ (8).
        adf_bytes = T * (8 + 16 * M + 24 * M + 8 * n + 8)
        adf_ms = phase["adf"] / args.steps
        ach = adf_bytes / (adf_ms * 1e-3) / 1e9 if adf_ms > 0 else 0.0
        # whole-run figure with SURVEY 8d's per-step-per-sweep bytes (8 + 24n + 88M)
        sweep_bytes = (8 + 24 * n + 88 * M) * T * itts
        sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
        line = {
            "metric": "EP filter+smoother time-steps/sec (FP64)", "value": value, "unit": "time-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(T, itts),
            "e2e": {"value": e2e_value, "unit": "time-steps/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "call": "nsagp_ep_ihgp (C ABI, host buffers)"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "ihgp_adf_cta_kernel (sequential ADF filter pass, one CTA per signal)",
                         "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                         "traffic": TRAFFIC_ADF_BYTES_PER_STEP * T / 1e9 if TRAFFIC_ADF_BYTES_PER_STEP else None,
                         "traffic_unit": "GB per launch (ncu dram__bytes_read+write, profiles/)",
                         "peak_source": peak_src,
                         "note": "the pass is a nonlinear recurrence in time (step k needs the posterior of step k-1), "
                                 "so it is bound by the latency of one step on one SM, not by HBM: see cycles_per_time_step",
                         "cycles_per_time_step": adf_ms * 1e-3 * sm_hz / T if adf_ms > 0 else None,
                         "whole_run_GBps": sweep_bytes / (dev_ms / args.steps * 1e-3) / 1e9},
            "phases_ms_per_step": {k: v / args.steps for k, v in phase.items()},
            "ep_itts_1": {"note": "what the reference's training runs (experiments/train_model.m:59-60): ONE sweep = the "
                                  "sequential first filter pass; time-steps/s of that pass alone on one signal",
                          "steps_per_s": world * T / (adf_ms * 1e-3) if adf_ms > 0 else None},
            "adf_only_steps_per_s": T / (adf_ms * 1e-3) if adf_ms > 0 else None,
            "setup_s": {"host_model_and_dare_tables": setup_s, "signal_generation": gen_s},
            "check": {"nlZ_first": float(chk["nlZ"][0]), "nlZ_last": float(chk["nlZ"][-1]), "n_negcav": chk["n_negcav"]},
            "wall_s": wall,
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = 1
            Tc, ic = 8000, EP_ITTS
            v, per = cpu_arm(1, 0, Tc, ic, cores)
            line["cpu_baseline"] = {"value": v, "unit": "time-steps/s", "cores": cores, "kind": "port",
                                    "sample": "plain-C port of matlab/ihgp_ep_modulator_nmf.m (oracle/c/nsagp_oracle.c, same dense "
                                              "operations), same model, T=%d ep_itts=%d, 1 core (%.1f s)" % (Tc, ic, per)}
    # ---- the other BASELINE configurations (strong-scaling workloads; all ranks take part) ----------------------
    extras = {}
    if not args.no_extras:
        import bench_workloads as bw
        ctx = bw.Ctx(nsagp, torch, dist, rank, world, local_rank)
        t_ex = time.perf_counter()
        for name, fn in (("c3_one_signal_500k", lambda: bw.one_signal_chunked(ctx, "full", 500000, 20, 7, label="C3")),
                         ("ihgp_one_signal_10M", lambda: bw.one_signal_chunked(ctx, "ihgp", args.long_T, 20, 9, reps=1, warm=1,
                                                                               exact=(world == 1), burnin=100000,
                                                                               label="north_star 10M", ep_itts_1=True)),
                         ("c5_batch_256_clips", lambda: bw.c5_batch(ctx)),
                         ("c4_giekf", lambda: bw.c4_giekf(ctx, cpu_baseline=(world == 1)))):
            try:
                extras[name] = fn()
            except Exception as e:                       # an extra workload must never take the headline line down
                extras[name] = {"error": "%s: %s" % (type(e).__name__, e)}
            L.nsagp_release_cache()
            if dist is not None:
                # a failure is symmetric (a rank that loses its peer times out in the exchange kernel after 10 s and
                # raises too), so every rank arrives here; meet again before the next workload
                try:
                    dist.barrier()
                except Exception as e:
                    extras.setdefault(name, {})["barrier_error"] = "%s: %s" % (type(e).__name__, e)
        extras["wall_s"] = time.perf_counter() - t_ex
    if rank == 0:
        if extras:
            line["extras"] = extras
        if not args.no_parity:
            try:
                line["lZ_rel_err"] = prefix_check(nsagp, hyp=hyp, y=y, Tp=min(2000, T), itts=itts)
            except Exception as e:
                line["lZ_rel_err"] = {"error": "%s: %s" % (type(e).__name__, e)}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--T", type=int, default=T_FULL)
    ap.add_argument("--ep-itts", dest="ep_itts", type=int, default=EP_ITTS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the other BASELINE configurations (extras object)")
    ap.add_argument("--no-parity", action="store_true", help="skip the lZ rel. error check against the oracle")
    ap.add_argument("--long-T", dest="long_T", type=int, default=10000000)
    ap.add_argument("--ref-T", dest="ref_T", type=int, default=2000, help="reference arm: samples per signal per step")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
