"""Time-chunked execution of one signal (nonstationary-audio-gp_b200/chunked.py): the EP schedule
with its frozen-site passes sharded over several ranks must reproduce the single-plan run.  On one
GPU the ranks are emulated by threads (ThreadComm), each driving its own plan on its own stream;
the only difference to a multi-GPU run is the transport of the few-KB carries."""
import threading

import numpy as np
import pytest

from conftest import make_problem, rel_err

pytestmark = pytest.mark.gpu


def _plans(nsagp, pb, itts, damping, count, kind="ihgp"):
    L = nsagp._lib
    hyp = pb["hyp"]
    F, Lm, Qc, H, Pinf = nsagp.ss_modulators_nmf(hyp.w_sub(), hyp.w_mod(), pb["kernel1"], pb["kernel2"])[:5]
    if kind == "ihgp":
        F, Lm, H, Pinf = nsagp.ssmodel.balance(F, Lm, H, Pinf)
    A, Q = nsagp.lti_disc(F, Lm, Qc, 1.0)
    if kind == "ihgp":
        Q = (Q + Q.T) / 2
    mdl = nsagp.to_block_model(A, Q, H, Pinf, pb["D"], pb["N"])
    tabs = [nsagp.tables.build_tables(mdl, want_smoother=True)] if kind == "ihgp" else None
    mk = lambda: nsagp.Plan(L.KIND_IHGP if kind == "ihgp" else L.KIND_FULL, [mdl],
                            [(pb["mom_gpu"], np.log([hyp.w_lik]), hyp.W)], 0.75, damping, itts,
                            pb["y"][None, :], L.MODE_PREDICT, tables=tabs)
    return [mk() for _ in range(count)]


@pytest.mark.parametrize("kind", ["ihgp", "full"])
@pytest.mark.parametrize("world,T,gaps", [(2, 1500, False), (3, 2000, True), (4, 1111, False)])
def test_chunked_matches_single_plan(nsagp, gpu_lib, world, T, gaps, kind):
    itts = 4
    damping = np.linspace(0.3, 0.1, itts)
    pb = make_problem(nsagp, 6, 3, T, "exp", "matern52", seed=5 + world, kind="precalc", p=9, shift=1.0, gaps=gaps)
    names = ("Eft", "Varft", "lb", "ub", "ttau", "tnu", "R", "MS", "MF", "nlZ", "maxDiffM", "n_negcav")
    single = _plans(nsagp, pb, itts, damping, 1, kind)[0]
    single.run()
    ref = single.fetch(0, names)
    plans = _plans(nsagp, pb, itts, damping, world, kind)
    run = nsagp.chunked.run_ihgp_chunked if kind == "ihgp" else nsagp.chunked.run_full_chunked
    comms = nsagp.chunked.ThreadComm.make(world)
    results, errors = [None] * world, []

    def work(r):
        try:
            ranges = run(plans[r], comms[r], damping)
            results[r] = nsagp.chunked.gather_outputs(plans[r], comms[r], ranges, names)
        except Exception as e:                                   # pragma: no cover - surfaced below
            errors.append(e)
            try:
                comms[r].s.barrier.abort()
            except Exception:
                pass

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors, errors
    for r in range(world):
        got = results[r]
        assert rel_err(got["nlZ"], ref["nlZ"]) < 1e-8
        for k in ("Eft", "Varft", "lb", "ub", "ttau", "tnu", "R", "MS", "MF"):
            assert rel_err(got[k], ref[k]) < 1e-6, (k, r)
        assert rel_err(got["maxDiffM"], ref["maxDiffM"]) < 1e-6
        assert got["n_negcav"] == ref["n_negcav"]
    for p in plans + [single]:
        p.close()
