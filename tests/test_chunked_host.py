"""Host logic of the time-chunked runs on CPU: range splitting and the exchange primitives over a
world_size-2 gloo group (the GPU stages themselves are covered by tests/test_gpu_chunked.py)."""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_split_ranges(nsagp):
    sr = nsagp.chunked.split_ranges
    assert sr(10, 3) == [(0, 4), (4, 7), (7, 10)]
    for T, w in [(100000, 8), (17, 8), (1000003, 7)]:
        r = sr(T, w)
        assert r[0][0] == 0 and r[-1][1] == T and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1
    with pytest.raises(ValueError):
        sr(7, 4)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    nsagp = importlib.import_module("nonstationary-audio-gp_b200")
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    comm = nsagp.chunked.TorchComm()
    parts = comm.allgather(np.arange(4.0) + 10 * rank)
    s = comm.allreduce(np.array([rank + 1.0, 2.0]), "sum")
    m = comm.allreduce(np.array([rank + 1.0, -rank]), "max")
    q.put((rank, [p.tolist() for p in parts], s.tolist(), m.tolist()))
    dist.destroy_process_group()


def test_torch_comm_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29640 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    out = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(60) for p in procs]
    for rank, parts, s, m in out:
        assert parts == [[0.0, 1.0, 2.0, 3.0], [10.0, 11.0, 12.0, 13.0]]
        assert s == [3.0, 4.0] and m == [2.0, 0.0]
