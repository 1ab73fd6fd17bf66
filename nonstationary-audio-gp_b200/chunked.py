"""One long signal, time-chunked over the GPUs of a box (north_star: "partitioned by
time-chunking the scan, using NCCL only for the O(state^2) per-chunk carry exchange and the lZ
reductions").

Infinite-horizon predict mode (ihgp_ep_modulator_nmf.m:195-526).  Every rank builds the same plan
over the whole signal.  The first EP iteration's filter pass (ADF) is a nonlinear recurrence in
time and runs replicated on every rank -- no communication, no speed-up.  Every later pass is
linear in the state once the sites are frozen, so rank r executes it only over its own range of
time steps and the ranks exchange, per pass,

  * one scan aggregate each (M affine maps: M*(b*b+b) doubles, ~1.8 KB at D=16, N=3),
  * the sites of the step left of each range (the filter looks one step back, :239),
  * the mean at the ends (m carried into the next filter pass, :198; the smoother's start),
  * the lZ partial sums and max-diff diagnostics (a few scalars).

``Comm`` is the exchange: ``TorchComm`` over torch.distributed (NCCL on GPUs, gloo in CPU tests
of the host logic), ``ThreadComm`` for several emulated ranks inside one process (one GPU).
"""
import ctypes as C
import threading

import numpy as np

from . import _lib

(ST_RESET, ST_ADF, ST_SUM_LZ, ST_FILTER_REDUCE, ST_FILTER_APPLY, ST_LAST_STEP, ST_COPY_MF, ST_SMOOTHER_REDUCE,
 ST_SMOOTHER_APPLY, ST_SITE_UPDATE, ST_CARRY_MEAN, ST_GET_SITES, ST_SET_SITES, ST_GET_MEAN, ST_SET_MEAN, ST_GET_MCARRY,
 ST_SET_MCARRY, ST_GET_DIAG, ST_RESET_DIAG, ST_GET_VM0, ST_SET_VM0, ST_SET_TRACE, ST_FINISH, ST_GET_COV,
 ST_SET_COV) = range(25)


def split_ranges(T, world):
    """Contiguous, balanced time ranges [t0, t1) covering [0, T); every rank gets >= 2 steps."""
    if world < 1 or T < 2 * world:
        raise ValueError("need at least two time steps per rank")
    base, extra = divmod(T, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


class TorchComm:
    """Exchange over a torch.distributed process group (payloads are a few KB: latency bound)."""

    def __init__(self, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.dev = (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl"
                    else torch.device("cpu"))

    def allgather(self, a):
        t = self.torch.as_tensor(np.ascontiguousarray(a, float), device=self.dev)
        parts = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(parts, t, group=self.group)
        return [p.cpu().numpy() for p in parts]

    def allreduce(self, a, op):
        t = self.torch.as_tensor(np.atleast_1d(np.asarray(a, float)).copy(), device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM if op == "sum" else self.dist.ReduceOp.MAX, group=self.group)
        return t.cpu().numpy()


class ThreadComm:
    """Emulated ranks: one thread per rank inside one process (each thread drives its own plan on its
    own CUDA stream), meeting at a barrier.  Lets the chunked schedule be tested on a single GPU."""

    class _Shared:
        def __init__(self, world):
            self.world = world
            self.slots = [None] * world
            self.barrier = threading.Barrier(world)

    def __init__(self, shared, rank):
        self.s, self.rank, self.world = shared, rank, shared.world

    @classmethod
    def make(cls, world):
        shared = cls._Shared(world)
        return [cls(shared, r) for r in range(world)]

    def allgather(self, a):
        self.s.slots[self.rank] = np.array(a, float, copy=True)
        self.s.barrier.wait()
        out = [x.copy() for x in self.s.slots]
        self.s.barrier.wait()
        return out

    def allreduce(self, a, op):
        parts = np.stack(self.allgather(np.atleast_1d(np.asarray(a, float))))
        return parts.sum(axis=0) if op == "sum" else parts.max(axis=0)


def _stage(plan, op, x=0.0, k=0, inp=None, n_out=0):
    inp_a = None if inp is None else np.ascontiguousarray(inp, float)
    out = np.empty(n_out) if n_out else None
    _lib.check(_lib.lib().nsagp_plan_stage(plan._h, op, float(x), int(k),
                                           None if inp_a is None else _lib.dptr(inp_a), 0 if inp_a is None else inp_a.size,
                                           None if out is None else _lib.dptr(out), n_out))
    return out


def run_ihgp_chunked(plan, comm, ep_damping):
    """Run the whole EP schedule of a KIND_IHGP / MODE_PREDICT plan (B = 1) with the frozen-site
    passes sharded over ``comm``'s ranks.  Afterwards ``gather_outputs`` assembles the result."""
    mdl = plan.models[0]
    T, M, n, itts = plan.T, mdl.M, mdl.n, plan.ep_itts
    BM = {1: 2, 2: 2, 3: 3, 4: 4, 5: 6, 6: 6, 7: 8, 8: 8}[max(mdl.bz, mdl.bg)]
    W = BM * BM + BM
    rank, world = comm.rank, comm.world
    ranges = split_ranges(T, world)
    t0, t1 = ranges[rank]
    _lib.check(_lib.lib().nsagp_plan_set_range(plan._h, t0, t1))
    damping = np.atleast_1d(np.asarray(ep_damping, float))
    last = world - 1
    _stage(plan, ST_RESET)
    damp = damping[0]
    for itt in range(1, itts + 1):
        _stage(plan, ST_RESET_DIAG)
        if itt == 1:
            _stage(plan, ST_ADF, x=damp)                       # replicated: identical on every rank
            _stage(plan, ST_SET_TRACE, x=-_stage(plan, ST_SUM_LZ, k=1, n_out=1)[0], k=0)
        else:
            # sites left of each range (the filter element of step k looks at R(:,k-1), :239)
            edge = comm.allgather(_stage(plan, ST_GET_SITES, k=t1 - 1, n_out=3 * M))
            if rank > 0:
                _stage(plan, ST_SET_SITES, k=t0 - 1, inp=edge[rank - 1])
            aggs = comm.allgather(_stage(plan, ST_FILTER_REDUCE, n_out=M * W))
            _stage(plan, ST_FILTER_APPLY, inp=np.concatenate(aggs[:rank]) if rank else np.zeros(0))
            if rank == last:
                _stage(plan, ST_LAST_STEP, x=damp)
        if itt == itts:
            _stage(plan, ST_COPY_MF)
        if itt < itts:
            damp = damping[itt]
        # the smoother starts from the filtered mean of the last step (held by the last rank)
        mT = comm.allgather(_stage(plan, ST_GET_MEAN, k=T - 1, n_out=n))[last]
        _stage(plan, ST_SET_MEAN, k=T - 1, inp=mT)
        aggs = comm.allgather(_stage(plan, ST_SMOOTHER_REDUCE, n_out=M * W))
        right = aggs[rank + 1:][::-1]                          # processing order: from the end of the signal
        _stage(plan, ST_SMOOTHER_APPLY, inp=np.concatenate(right) if right else np.zeros(0))
        if itt < itts:
            _stage(plan, ST_SITE_UPDATE, x=damp, k=int(itt > 1))
            lz = comm.allreduce(_stage(plan, ST_SUM_LZ, k=0, n_out=1), "sum")[0]
            _stage(plan, ST_SET_TRACE, x=-lz, k=itt)
            # m carries from the smoother (k = 0, rank 0) into the next filter pass (:198)
            _stage(plan, ST_CARRY_MEAN)
            _stage(plan, ST_SET_MCARRY, inp=comm.allgather(_stage(plan, ST_GET_MCARRY, n_out=M * BM))[0])
        bits = _stage(plan, ST_GET_DIAG, n_out=2)              # bit patterns of non-negative doubles: max == max of values
        _stage(plan, ST_SET_TRACE, k=-1 - (itt - 1), inp=comm.allreduce(bits, "max"))
    _stage(plan, ST_SET_VM0, inp=comm.allgather(_stage(plan, ST_GET_VM0, n_out=M))[0])
    _stage(plan, ST_FINISH)
    return ranges


def run_full_chunked(plan, comm, ep_damping):
    """The same for a KIND_FULL / MODE_PREDICT plan (gf_ep_modulator_nmf.m:113-283): the first filter
    pass replicated, the parallel Kalman-filter scan, the RTS scan and the site updates sharded.  The
    full-state filter element of step k uses only the sites of step k, so no site halo is needed; the
    carries are (A, b, C, eta, J) resp. (E, g, L) per latent block."""
    mdl = plan.models[0]
    T, M, n, itts = plan.T, mdl.M, mdl.n, plan.ep_itts
    BM = {1: 2, 2: 2, 3: 3, 4: 4, 5: 6, 6: 6, 7: 8, 8: 8}[max(mdl.bz, mdl.bg)]
    Wf, Ws, PB = 3 * BM * BM + 2 * BM, 2 * BM * BM + BM, M * BM * BM
    rank, world = comm.rank, comm.world
    ranges = split_ranges(T, world)
    t0, t1 = ranges[rank]
    _lib.check(_lib.lib().nsagp_plan_set_range(plan._h, t0, t1))
    damping = np.atleast_1d(np.asarray(ep_damping, float))
    last = world - 1
    _stage(plan, ST_RESET)
    damp = damping[0]
    for itt in range(1, itts + 1):
        _stage(plan, ST_RESET_DIAG)
        if itt == 1:
            _stage(plan, ST_ADF, x=damp)
            _stage(plan, ST_SET_TRACE, x=-_stage(plan, ST_SUM_LZ, k=1, n_out=1)[0], k=0)      # nlZ(1) (:186-188)
        else:
            aggs = comm.allgather(_stage(plan, ST_FILTER_REDUCE, n_out=M * Wf))
            _stage(plan, ST_FILTER_APPLY, inp=np.concatenate(aggs[:rank]) if rank else np.zeros(0))
            if rank == last:
                _stage(plan, ST_LAST_STEP, x=damp)
        if itt == itts:
            _stage(plan, ST_COPY_MF)
        if itt < itts:
            damp = damping[itt]
        state = comm.allgather(np.concatenate([_stage(plan, ST_GET_MEAN, k=T - 1, n_out=n),
                                               _stage(plan, ST_GET_COV, k=T - 1, n_out=PB)]))[last]
        _stage(plan, ST_SET_MEAN, k=T - 1, inp=state[:n])
        _stage(plan, ST_SET_COV, k=T - 1, inp=state[n:])
        aggs = comm.allgather(_stage(plan, ST_SMOOTHER_REDUCE, n_out=M * Ws))
        right = aggs[rank + 1:][::-1]
        _stage(plan, ST_SMOOTHER_APPLY, inp=np.concatenate(right) if right else np.zeros(0))
        if itt < itts:
            _stage(plan, ST_SITE_UPDATE, x=damp)
            lz = comm.allreduce(_stage(plan, ST_SUM_LZ, k=0, n_out=1), "sum")[0]
            _stage(plan, ST_SET_TRACE, x=-lz, k=itt)                                          # nlZ(itt+1) (:276-278)
        bits = _stage(plan, ST_GET_DIAG, n_out=2)
        _stage(plan, ST_SET_TRACE, k=-1 - (itt - 1), inp=comm.allreduce(bits, "max"))
    _stage(plan, ST_FINISH)
    return ranges


# ------------------------------------------------------------------- device-side carry exchange
class DeviceComm:
    """Mailbox exchange over NVLink peer access (csrc/comm.cuh; C ABI nsagp_comm_*): after a one-time set-up the
    ranks never meet on the host again during an EP run.  ``connect_torch`` exchanges the CUDA IPC handles of the
    mailboxes once through torch.distributed (one process per GPU); ``connect_threads`` wires emulated ranks that
    are threads of one process (one GPU) by device address."""

    def __init__(self, rank, world, slot_doubles):
        self.rank, self.world = int(rank), int(world)
        self._h = C.c_void_p()
        _lib.check(_lib.lib().nsagp_comm_create(C.byref(self._h), self.rank, self.world, int(slot_doubles)))

    def export(self):
        handle = (C.c_ubyte * 64)()
        ptr = C.c_uint64()
        _lib.check(_lib.lib().nsagp_comm_export(self._h, handle, C.byref(ptr)))
        return bytes(handle), int(ptr.value)

    def connect_handles(self, handles):
        buf = (C.c_ubyte * (64 * self.world)).from_buffer_copy(b"".join(handles))
        _lib.check(_lib.lib().nsagp_comm_connect(self._h, buf, None))
        return self

    def connect_ptrs(self, ptrs):
        arr = (C.c_uint64 * self.world)(*[int(p) for p in ptrs])
        _lib.check(_lib.lib().nsagp_comm_connect(self._h, None, arr))
        return self

    @classmethod
    def connect_torch(cls, plan, group=None):
        """One process per GPU: all ranks call this with their (identically shaped) plan."""
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        cm = cls(rank, world, _lib.lib().nsagp_plan_comm_slot_doubles(plan._h))
        handle, _ = cm.export()
        handles = [None] * world
        dist.all_gather_object(handles, handle, group=group)
        if world > 1:
            cm.connect_handles(handles)
        return cm

    @classmethod
    def connect_threads(cls, plans):
        """Emulated ranks inside one process: one comm per plan, wired by device address."""
        world = len(plans)
        cms = [cls(r, world, _lib.lib().nsagp_plan_comm_slot_doubles(pl._h)) for r, pl in enumerate(plans)]
        ptrs = [cm.export()[1] for cm in cms]
        if world > 1:
            for cm in cms:
                cm.connect_ptrs(ptrs)
        return cms

    def close(self):
        if self._h:
            _lib.lib().nsagp_comm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def run_chunked_device(plan, dcomm):
    """The whole EP schedule of a predict-mode plan (either kind, B = 1), frozen-site passes sharded over the ranks
    of ``dcomm``, carries exchanged on the device.  One C call, no host synchronisation inside.  With
    ``plan.set_adf_parallel(chunks, burnin)`` the first filter pass is sharded too (approximate, error reported by
    ``plan.adf_mismatch()``); otherwise it runs replicated and exact on every rank."""
    ranges = split_ranges(plan.T, dcomm.world)
    t0, t1 = ranges[dcomm.rank]
    _lib.check(_lib.lib().nsagp_plan_set_range(plan._h, t0, t1))
    _lib.check(_lib.lib().nsagp_plan_run_chunked(plan._h, dcomm._h))
    return ranges


def gather_outputs(plan, comm, ranges, names=("Eft", "Varft", "lb", "ub")):
    """Assemble the time-indexed outputs from every rank's own range; every rank gets the result."""
    res = plan.fetch(0, tuple(names))
    out = {}
    for nm, a in res.items():
        if isinstance(a, np.ndarray) and a.ndim == 2 and a.shape[1] == plan.T and not (nm == "Varft" and plan.kind == _lib.KIND_IHGP):
            parts = comm.allgather(np.ascontiguousarray(a))
            full = np.empty_like(a)
            for r, (lo, hi) in enumerate(ranges):
                full[:, lo:hi] = parts[r][:, lo:hi]
            out[nm] = full
        elif isinstance(a, np.ndarray) and a.ndim == 1 and a.size == plan.T:
            parts = comm.allgather(a)
            full = np.empty_like(a)
            for r, (lo, hi) in enumerate(ranges):
                full[lo:hi] = parts[r][lo:hi]
            out[nm] = full
        else:
            out[nm] = a
    if "n_negcav" in res:
        out["n_negcav"] = int(comm.allreduce(float(res["n_negcav"]), "sum")[0])
    return out
