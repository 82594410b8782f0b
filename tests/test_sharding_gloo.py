"""Host-side logic of the N > 1 path on CPU: shard arithmetic and the rendezvous helper (rank 0 creates
the id, torch.distributed broadcasts it), world_size 2 over gloo.  The device side of the sharded
resampler is covered by tests/test_gpu_sharded.py on a multi-GPU box."""
import os
import sys

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import wsb200 as ws
    made = []

    def fake_id():
        made.append(1)
        return bytes(range(128))

    captured = {}

    class FakeState:
        def __init__(self, n, **kw):
            captured.update(kw, n=n)

    orig = ws.core.SMCState
    ws.core.SMCState = FakeState
    try:
        ws.core.sharded_state(1001, make_id=fake_id, ess_perc_min=0.7)
    finally:
        ws.core.SMCState = orig
    lo, hi = ws.shard_bounds(1001, rank, world)
    q.put((rank, len(made), captured["nccl_id"], captured["rank"], captured["nranks"], captured["n"], lo, hi))
    dist.destroy_process_group()


def test_sharded_state_rendezvous_and_bounds():
    world, port = 2, 29731
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, made0, id0, rk0, nr0, n0, lo0, hi0), (r1, made1, id1, rk1, nr1, n1, lo1, hi1) = out
    assert made0 == 1 and made1 == 0                 # only rank 0 creates the id
    assert id0 == id1 == bytes(range(128))
    assert (rk0, nr0, rk1, nr1) == (0, 2, 1, 2) and n0 == n1 == 1001
    assert (lo0, hi0, lo1, hi1) == (0, 500, 500, 1001)  # contiguous, covers everything


def test_shard_bounds_partition():
    import wsb200 as ws
    for n in (1, 7, 1000, 10 ** 8 + 3):
        for R in (1, 2, 3, 8):
            b = [ws.shard_bounds(n, r, R) for r in range(R)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(R - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1
