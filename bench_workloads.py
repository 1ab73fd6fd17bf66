"""The other BASELINE.json configurations as measured workloads of bench.py (the C2 headline stays in bench.py):

  c3      ONE gf_ep_modulator_nmf signal, T = 500 000, D=16 x N=3 (n = 41), ep_itts = 20, time-chunked over the ranks
          with the device-side carry exchange (csrc/comm.cuh) -- strong scaling; exact first pass (replicated) and
          the opt-in parallel-in-time first pass (burn-in overlap, measured boundary mismatch reported)
  ihgp10m ONE ihgp_ep_modulator_nmf signal of 10 M samples (north_star: "near-linear 8-GPU scaling on a 10M-sample
          signal"), same two forms
  c5      256 clips x hyper-parameter grid, nlZ mode, 10 M time steps in total, sharded over the ranks (strong
          scaling; the only communication is the gather of B scalars)
  c4      gf_giekf_modulator_nmf (iterated EKF, dense n = 73, missing-data gaps), one signal per GPU

Every function returns a dict for the bench line; times are CUDA-event device times, max over ranks.
"""
import ctypes as C
import time

import numpy as np

D, N = 16, 3
K1, K2 = "exp", "matern52"
ALPHA, SHIFT, P_CUB = 0.75, 1.0, 9


class Ctx:
    def __init__(self, nsagp, torch, dist, rank, world, local_rank):
        self.nsagp, self.torch, self.dist, self.rank, self.world, self.local_rank = nsagp, torch, dist, rank, world, local_rank
        self.lm = nsagp._lib
        self.L = nsagp._lib.lib()

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, vals):
        t = self.torch.tensor([float(v) for v in vals], dtype=self.torch.float64, device="cuda")
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t.cpu()]


def _tiled_signal(nsagp, hyp, k1, k2, T, seed, piece=1000000, distinct=2):
    """A T-sample synthetic signal.  Beyond `piece` samples, `distinct` independent draws of the same model are
    concatenated in turn (generation is host-side NumPy, 3 us per sample, and would otherwise dominate the bench's wall
    clock; throughput does not depend on the sample values)."""
    cache, out, s = {}, [], 0
    while s < T:
        m = min(piece, T - s)
        j = (s // piece) % distinct
        if j not in cache or cache[j].size < m:
            cache[j] = nsagp.synth.sample_signal(hyp, k1, k2, m, np.random.default_rng(seed + j), link_shift=SHIFT, sqrt_model=True)[0]
        out.append(cache[j][:m])
        s += m
    return np.concatenate(out)


def _model(nsagp, hyp, k1, k2, D_, N_, ihgp, smoother=True):
    F, Lm, Qc, H, Pinf = nsagp.ss_modulators_nmf(hyp.w_sub(), hyp.w_mod(), k1, k2)[:5]
    if ihgp:
        F, Lm, H, Pinf = nsagp.ssmodel.balance(F, Lm, H, Pinf)
    A, Q = nsagp.lti_disc(F, Lm, Qc, 1.0)
    if ihgp:
        Q = (Q + Q.T) / 2
    mdl = nsagp.to_block_model(A, Q, H, Pinf, D_, N_)
    return mdl, (nsagp.tables.build_tables(mdl, want_smoother=smoother) if ihgp else None)


def _hbm_peak():
    try:
        import json, os
        return float(json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0                      # the profiling guide's fallback


def _mom(nsagp):
    wn, xn = nsagp.utp_ws(P_CUB, N)
    return nsagp.likModulatorPreCalcwn(nsagp.Softplus(SHIFT), wn, xn)


def _timed_chunked(ctx, plan, dcomm, reps, warm):
    ms, phases = [], []
    for i in range(warm + reps):
        ctx.barrier()
        ctx.nsagp.chunked.run_chunked_device(plan, dcomm)
        if i >= warm:
            tm = plan.timings()
            ms.append(tm["total"]); phases.append(tm)
    t = float(np.mean(ctx.max_over_ranks([np.mean(ms)])))
    ph = {k: float(np.mean([p[k] for p in phases])) for k in phases[0]}
    return t, ph


def one_signal_chunked(ctx, kind, T, itts, seed, reps=2, warm=1, chunks_per_gpu=148, burnin=60000, exact=True, label="",
                       ep_itts_1=False):
    """ONE signal of T samples time-chunked over the ranks (strong scaling)."""
    nsagp, lm = ctx.nsagp, ctx.lm
    ihgp = kind == "ihgp"
    t0 = time.perf_counter()
    rng = np.random.default_rng(seed)
    hyp = nsagp.synth.speech_hypers(D, N, rng)
    y = _tiled_signal(nsagp, hyp, K1, K2, T, seed * 1000)
    gen_s = time.perf_counter() - t0
    mdl, tabs = _model(nsagp, hyp, K1, K2, D, N, ihgp)
    damp = np.linspace(0.01, 0.1, itts)
    plan = nsagp.Plan(lm.KIND_IHGP if ihgp else lm.KIND_FULL, [mdl], [(_mom(nsagp), np.log([hyp.w_lik]), hyp.W)], ALPHA, damp, itts,
                      y[None, :], lm.MODE_PREDICT, tables=[tabs] if ihgp else None)
    dcomm = nsagp.chunked.DeviceComm.connect_torch(plan) if ctx.dist is not None else nsagp.chunked.DeviceComm.connect_threads([plan])[0]
    res = {"workload": "%s: ONE %s signal, D=%d x N=%d (n=%d), T=%d, ep_itts=%d, time-chunked over %d GPU(s), device-side carry "
                       "exchange over NVLink peer memory" % (label, "ihgp_ep_modulator_nmf" if ihgp else "gf_ep_modulator_nmf", D, N, mdl.n,
                                                              T, itts, ctx.world),
           "scaling": "strong", "T": T, "ep_itts": itts, "signal_generation_s": gen_s,
           "collective": "comm_exchange_kernel: P2P stores into peer mailboxes + release/acquire flags, %d exchanges per run "
                         "(+1 with the parallel first pass)" % (3 * itts - 1)}
    units = T * itts
    if exact:
        ms, ph = _timed_chunked(ctx, plan, dcomm, reps, warm)
        res["exact"] = {"first_pass": "sequential, replicated on every rank (exact)", "ms": ms, "steps_per_s": units / ms * 1e3, "phases_ms": ph}
        nlz_exact = plan.fetch(0, ("nlZ",))["nlZ"]
    plan.set_adf_parallel(chunks_per_gpu, burnin)
    ms, ph = _timed_chunked(ctx, plan, dcomm, reps, warm)
    mis, scale = plan.adf_mismatch()
    mis, scale = ctx.max_over_ranks([mis, scale])
    res["parallel_first_pass"] = {"first_pass": "burn-in overlap: %d chunks per GPU, %d steps of burn-in each (opt-in, approximate)" % (chunks_per_gpu, burnin),
                                  "ms": ms, "steps_per_s": units / ms * 1e3, "phases_ms": ph,
                                  "boundary_mismatch_rel": mis / scale if scale > 0 else None}
    if ihgp:
        # the frozen-site passes against the HBM roof (SURVEY 8d / BASELINE.md 4: 8 + 24 n + 88 M algorithmic bytes per time
        # step and filter+smoother sweep); a run has itts - 1 frozen filter passes and itts smoother passes
        bps = 8 + 24 * mdl.n + 88 * mdl.M
        sweeps = (2 * itts - 1) / 2.0
        frozen_ms = ph["fixed_filter"] + ph["smoother"]
        gbs = bps * (T / ctx.world) * sweeps / (frozen_ms * 1e-3) / 1e9
        res["frozen_sweep_roofline"] = {"bound": "hbm", "kernels": "scan_reduce2 / scan_apply2_kernel<FilterElem | SmootherElem> + carry",
                                        "algorithmic_bytes_per_step_and_sweep": bps, "achieved": gbs, "unit": "GB/s per GPU",
                                        "peak": _hbm_peak(), "frac": gbs / _hbm_peak(),
                                        "ms_per_sweep": frozen_ms / sweeps}
    if exact:
        nlz_par = plan.fetch(0, ("nlZ",))["nlZ"]
        res["parallel_first_pass"]["nlZ_rel_dev_vs_exact"] = float(np.max(np.abs(nlz_par - nlz_exact) / np.abs(nlz_exact)))
        res["parallel_first_pass"]["speedup_vs_exact_same_gpus"] = res["exact"]["ms"] / ms
    plan.close()
    if ep_itts_1:
        # ONE sweep (ep_itts = 1: first filter pass + smoother), what the reference's training runs
        # (experiments/train_model.m:59-60), on the same signal with the parallel first pass
        plan1 = nsagp.Plan(lm.KIND_IHGP if ihgp else lm.KIND_FULL, [mdl], [(_mom(nsagp), np.log([hyp.w_lik]), hyp.W)], ALPHA, damp[:1], 1,
                           y[None, :], lm.MODE_PREDICT, tables=[tabs] if ihgp else None)
        plan1.set_adf_parallel(chunks_per_gpu, burnin)
        ms1, ph1 = _timed_chunked(ctx, plan1, dcomm, reps, warm)
        mis1, scale1 = ctx.max_over_ranks(plan1.adf_mismatch())
        res["ep_itts_1_parallel_first_pass"] = {"ms": ms1, "steps_per_s": T / ms1 * 1e3, "phases_ms": ph1,
                                                "boundary_mismatch_rel": mis1 / scale1 if scale1 > 0 else None}
        if exact:
            res["ep_itts_1_exact_first_pass_steps_per_s"] = T / (res["exact"]["phases_ms"]["adf"] + ph1["smoother"]) * 1e3
        plan1.close()
    dcomm.close()
    return res


def c5_batch(ctx, reps=2, warm=1, B=256, T=39062, n_clips=32, adf_form=None):
    """256 (clip, hyper-parameter point) problems in nlZ mode, 10 M time steps in total, contiguous shards per rank."""
    nsagp, lm = ctx.nsagp, ctx.lm
    t0 = time.perf_counter()
    rng = np.random.default_rng(11)
    base = nsagp.synth.speech_hypers(D, N, rng)
    clips = [nsagp.synth.sample_signal(base, K1, K2, T, np.random.default_rng(1000 + c), link_shift=SHIFT, sqrt_model=True)[0]
             for c in range(n_clips)]
    gen_s = time.perf_counter() - t0
    grid = [(ls, s2) for ls in (0.5, 0.8, 1.25, 2.0) for s2 in (0.5, 2.0)]          # len_slow scale x sigma^2 scale: 8 points
    lo = ctx.rank * B // ctx.world
    hi = (ctx.rank + 1) * B // ctx.world
    mom = _mom(nsagp)
    t0 = time.perf_counter()
    cache = {}
    probs = []
    for b in range(lo, hi):
        g = b % len(grid)
        if g not in cache:
            import copy
            h = copy.deepcopy(base)
            h.len_slow = base.len_slow * grid[g][0]
            h.w_lik = base.w_lik * grid[g][1]
            cache[g] = (h, _model(nsagp, h, K1, K2, D, N, True, smoother=False), _model(nsagp, h, K1, K2, D, N, False)[0])
        probs.append((cache[g], clips[(b // len(grid)) % n_clips]))
    setup_s = time.perf_counter() - t0
    ys = np.stack([p[1] for p in probs])
    out = {"workload": "C5: %d clips x %d-point hyper-parameter grid = %d problems x T=%d (%.1f M steps), nlZ mode, sharded over %d GPU(s)"
                       % (n_clips, len(grid), B, T, B * T / 1e6, ctx.world), "scaling": "strong", "signal_generation_s": gen_s,
           "host_setup_s": setup_s, "collective": "none on the data path (gather of B scalars)"}
    for name, kind, itts in (("ihgp_nlZ_ep_itts1", lm.KIND_IHGP, 1), ("gf_ep_nlZ_ep_itts3", lm.KIND_FULL, 3)):
        models = [p[0][1][0] if kind == lm.KIND_IHGP else p[0][2] for p in probs]
        tabs = [p[0][1][1] for p in probs] if kind == lm.KIND_IHGP else None
        liks = [(mom, np.log([p[0][0].w_lik]), p[0][0].W) for p in probs]
        with nsagp.Plan(kind, models, liks, ALPHA, np.linspace(0.1, 0.1, itts), itts, ys, lm.MODE_NLZ, tables=tabs) as plan:
            if adf_form is not None:
                plan.set_adf_form(adf_form)
            ms = []
            for i in range(warm + reps):
                ctx.barrier()
                plan.run()
                if i >= warm:
                    ms.append(plan.timings()["total"])
            t = ctx.max_over_ranks([np.mean(ms)])[0]
            sweeps = max(1, itts - 1)
            out[name] = {"ms": t, "steps_per_s": B * T * sweeps / t * 1e3, "sweeps_counted": sweeps,
                         "edata_first": plan.fetch(0, ("edata",))["edata"],
                         "edata_last": plan.fetch(len(probs) - 1, ("edata",))["edata"]}
    return out


def c4_giekf(ctx, T=100000, reps=2, cpu_baseline=True):
    """gf_giekf_modulator_nmf predict, D=32 exp subbands x N=3 matern52 modulators (dense n = 73), six gaps per 20k
    samples, g_iter = 1, through the host-buffer C entry point; one signal per GPU."""
    nsagp, lm, L = ctx.nsagp, ctx.lm, ctx.L
    Dk, Nk = 32, 3
    rng = np.random.default_rng(3 + ctx.rank)
    hyp = nsagp.synth.speech_hypers(Dk, Nk, rng, w_lik=1e-2)
    y, _, _ = nsagp.synth.sample_signal(hyp, K1, K2, T, rng)
    for s in range(0, T, 20000):
        for j, g in enumerate((10, 20, 40, 80, 160, 320)):
            a = s + 1500 + 3000 * j
            y[a:min(a + g, T)] = np.nan
    F, Lm, Qc, H, Pinf = nsagp.ss_modulators_nmf(hyp.w_sub(), hyp.w_mod(), K1, K2)[:5]
    F, Lm, H, Pinf = nsagp.ssmodel.balance(F, Lm, H, Pinf)
    A, Q = nsagp.lti_disc(F, Lm, Qc, 1.0)
    mdl = nsagp.to_block_model(A, Q, H, Pinf, Dk, Nk)
    arrs = [lm.as_f64(a) for a in (mdl.A, mdl.Q, mdl.Pinf, mdl.h)]
    cm = lm.Model()
    cm.D, cm.N, cm.bz, cm.bg = mdl.D, mdl.N, mdl.bz, mdl.bg
    cm.A, cm.Q, cm.Pinf, cm.h = [lm.dptr(a) for a in arrs]
    Wf = np.asfortranarray(hyp.W)
    M, n = mdl.M, mdl.n
    o = lm.Outputs()
    bufs = dict(Eft=np.zeros((T, M)), Varft=np.zeros((T, M)))
    for k, v in bufs.items():
        setattr(o, k, lm.dptr(v))
    yb = lm.as_f64(y)
    best, wall = None, None
    for rep in range(reps + 1):
        ctx.barrier()
        t0 = time.perf_counter()
        lm.check(L.nsagp_giekf(C.byref(cm), lm.dptr(Wf), float(hyp.w_lik), 1, 1, lm.dptr(yb), T, lm.MODE_PREDICT, C.byref(o)))
        w = time.perf_counter() - t0
        ms = np.zeros(2)
        lm.check(L.nsagp_giekf_timings(lm.dptr(ms), 2))
        if rep > 0 and (best is None or ms.sum() < best.sum()):
            best, wall = ms.copy(), w
    f_ms, s_ms, wall = ctx.max_over_ranks([best[0], best[1], wall])
    sm_hz = 1.965e9
    cpu = None
    if ctx.rank == 0 and cpu_baseline:
        # the oracle (NumPy restatement of matlab/gf_giekf_modulator_nmf.m, dense n x n arithmetic as in the reference)
        # on a bounded prefix: the checker timed as the CPU baseline, never the thing measured
        from oracle import giekf as ogk, ssmodel as oss
        Tc = 400
        ss_ref = lambda x, p1, p2, k1, k2: oss.ss_modulators_nmf(p1, p2, k1, k2)
        t = np.arange(1.0, Tc + 1.0)
        t0 = time.perf_counter()
        ogk.gf_giekf_modulator_nmf(hyp.pack_log(), t, y[:Tc], ss_ref, None, t, K1, K2, 1, Dk, Nk, 1, 1)
        dt = time.perf_counter() - t0
        cpu = {"value": Tc / dt, "unit": "time-steps/s", "cores": 1, "kind": "port",
               "sample": "oracle/giekf.py (NumPy, dense n=%d), first %d samples of the same signal, g_iter=1 (%.1f s)" % (n, Tc, dt)}
    grad = None
    if ctx.rank == 0:
        # the analytic-gradient mode of the same file (gf_giekf_modulator_nmf.m:296-437, GradObj = 'on'): energy and
        # 1 + 3D + 2N = 103 sensitivity recursions on a gap-free signal of the same model (a missing sample makes
        # the reference's energy NaN), through the public entry point with host buffers
        Tg = min(T, 50000)
        yg, _, _ = nsagp.synth.sample_signal(hyp, K1, K2, Tg, np.random.default_rng(11))
        ss_gpu = lambda x, p1, p2, k1, k2: nsagp.ss_modulators_nmf(p1, p2, k1, k2)
        tg = np.arange(1.0, Tg + 1.0)
        g_ms, g_wall = None, None
        for rep in range(2):
            t0 = time.perf_counter()
            e, g = nsagp.gf_giekf_modulator_nmf(hyp.pack_log(), tg, yg, ss_gpu, None, None, K1, K2, 1, Dk, Nk, 1, 1, GradObj="on")
            g_wall = time.perf_counter() - t0
            ms = np.zeros(2)
            lm.check(L.nsagp_giekf_timings(lm.dptr(ms), 2))
            g_ms = ms[0]
        grad = {"workload": "gf_giekf_modulator_nmf nlZ + analytic gradient (GradObj='on'), same model, T=%d, %d parameters = %d CTAs" % (Tg, g.size, g.size),
                "kernel_ms": g_ms, "steps_per_s": Tg / g_ms * 1e3, "e2e_steps_per_s": Tg / g_wall,
                "parameter_steps_per_s": g.size * Tg / g_ms * 1e3, "cycles_per_step": g_ms * 1e-3 * sm_hz / Tg,
                "finite": bool(np.isfinite(e) and np.all(np.isfinite(g)))}
        if cpu_baseline:
            from oracle import giekf as ogk, ssmodel as oss
            Tc = 100
            ss_ref = lambda x, p1, p2, k1, k2: oss.ss_modulators_nmf(p1, p2, k1, k2) + oss.ss_modulators_nmf_derivs(p1, p2, k1, k2)
            t0 = time.perf_counter()
            eo, go = ogk.gf_giekf_modulator_nmf(hyp.pack_log(), tg[:Tc], yg[:Tc], ss_ref, None, None, K1, K2, 1, Dk, Nk, 1, 1, GradObj="on")
            dt = time.perf_counter() - t0
            ep, gp = nsagp.gf_giekf_modulator_nmf(hyp.pack_log(), tg[:Tc], yg[:Tc], ss_gpu, None, None, K1, K2, 1, Dk, Nk, 1, 1, GradObj="on")
            grad["cpu_baseline"] = {"value": Tc / dt, "unit": "time-steps/s", "cores": 1, "kind": "port",
                                    "sample": "oracle/giekf.py giekf_energy_grad (NumPy, dense n=%d, %d parameters incl. the 2n x 2n expm stack), first %d samples (%.1f s)" % (n, go.size, Tc, dt)}
            grad["gradient_rel_err_vs_oracle_prefix"] = float(np.max(np.abs(gp - go) / np.abs(go)))      # worst ENTRY
    return {"cpu_baseline": cpu, "gradient_mode": grad, "workload": "C4: gf_giekf_modulator_nmf predict, D=32 x N=3 (dense n=%d), T=%d, g_iter=1, missing-data gaps, one signal per GPU" % (n, T),
            "scaling": "weak", "filter_ms": f_ms, "smoother_ms": s_ms, "steps_per_s": ctx.world * T / (f_ms + s_ms) * 1e3,
            "e2e_steps_per_s": ctx.world * T / wall, "filter_cycles_per_step": f_ms * 1e-3 * sm_hz / T,
            "smoother_dense_equiv_tflops": 12.3 * n ** 3 * T / (s_ms * 1e-3) / 1e12,
            "smoother_fp64_tensor_frac_of_37.2_tflops": 12.3 * n ** 3 * T / (s_ms * 1e-3) / 1e12 / 37.2}
