"""Batched nlZ evaluation over independent clips / hyper-parameter points, sharded
over the GPUs of one box (BASELINE config 5).

This is what ``fminunc(...,'GradObj','off')`` does around the reference's nlZ mode
(matlab/demo_toy_modulators_nmf.m:100-104, experiments/train_model.m:226-241):
many independent evaluations of ``-sum(lZ)``, one per parameter vector.  The
units are independent, so ranks take contiguous shards, run them as one batched
plan on their own GPU, and the only communication is the final gather of B
scalars (torch.distributed all_gather: NCCL on GPUs, gloo in CPU tests).
"""
import numpy as np

from . import _lib, entry, ssmodel, tables as tables_mod


def shard_range(B, world, rank):
    """Contiguous, balanced split of B units: first (B mod world) ranks get one more."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(B, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _gather(values, B, world, rank, group):
    """all_gather of ragged shards; every rank returns the full length-B vector."""
    if world == 1:
        return np.asarray(values, float)
    import torch
    import torch.distributed as dist
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    width = -(-B // world)
    buf = torch.full((width,), float("nan"), dtype=torch.float64, device=dev)
    buf[:len(values)] = torch.as_tensor(np.asarray(values, float), device=dev)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    out = np.empty(B)
    for r, p in enumerate(parts):
        lo, hi = shard_range(B, world, r)
        out[lo:hi] = p[:hi - lo].cpu().numpy()
    return out


def gpu_evaluator(kind, ss, mom, kernel1, kernel2, num_lik_params, D, N, ep_fraction, ep_damping, ep_itts,
                  running_sites=False, max_batch=4096):
    """Returns f(ws, ys) -> nlZ values for a list of parameter vectors and clips,
    evaluated as batched plans on the current GPU (nlZ mode of the entry points)."""
    def evaluate(ws, ys):
        out = []
        for s in range(0, len(ws), max_batch):
            models, liks, tabs = [], [], []
            for w in ws[s:s + max_batch]:
                lik_param, p1, p2, W = entry._unpack_log(w, num_lik_params, D, N)
                mdl, _ = entry._discrete_model(ss, None, p1, p2, kernel1, kernel2, D, N,
                                               balance=(kind == _lib.KIND_IHGP), symmetrise_Q=(kind == _lib.KIND_IHGP))
                models.append(mdl); liks.append((mom, lik_param, W))
                if kind == _lib.KIND_IHGP:
                    tabs.append(tables_mod.build_tables(mdl, want_smoother=False))
            mode = _lib.MODE_NLZ_RUNNING if running_sites else _lib.MODE_NLZ
            y = np.stack([np.asarray(v, float).ravel() for v in ys[s:s + max_batch]])
            with entry.Plan(kind, models, liks, ep_fraction, ep_damping, ep_itts, y, mode,
                            tables=tabs if kind == _lib.KIND_IHGP else None) as plan:
                plan.run()
                out.extend(plan.fetch(b, ("edata",))["edata"] for b in range(len(models)))
        return out
    return evaluate


def nlz_batch(ws, ys, evaluator, group=None):
    """Evaluate nlZ for B (parameter vector, clip) pairs, sharded over the ranks of
    ``group`` (or run locally when torch.distributed is not initialised).  ``ws``
    and ``ys`` are length-B sequences; ``ys`` may also be a single clip shared by
    all parameter vectors.  Every rank returns the full length-B result."""
    B = len(ws)
    if not isinstance(ys, (list, tuple)) and np.ndim(ys) == 1:
        ys = [ys] * B
    if len(ys) != B:
        raise ValueError("need one clip per parameter vector (or a single shared clip)")
    world, rank = 1, 0
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            world, rank = dist.get_world_size(group), dist.get_rank(group)
    except ImportError:
        pass
    lo, hi = shard_range(B, world, rank)
    vals = evaluator(list(ws[lo:hi]), list(ys[lo:hi])) if hi > lo else []
    return _gather(vals, B, world, rank, group)


def finite_difference_gradient(w, y, evaluator, h=1e-6, group=None):
    """Forward-difference gradient of nlZ at w as ONE batched call: nlZ(w) and
    nlZ(w + h e_i) for all i (what fminunc does one call at a time because the
    reference's analytic gradient is identically zero, gf_ep_modulator_nmf.m:363)."""
    w = np.asarray(w, float).ravel()
    pts = [w] + [w + h * np.eye(w.size)[i] for i in range(w.size)]
    vals = nlz_batch(pts, y, evaluator, group=group)
    return vals[0], (vals[1:] - vals[0]) / h
