"""Host-side logic of the N > 1 path on CPU: shard arithmetic and the rendezvous helper (rank 0 creates
the id, torch.distributed broadcasts it), world_size 2 over gloo.  The device side of the sharded
resampler is covered by tests/test_gpu_sharded.py on a multi-GPU box."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))


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


def _plan(L, bnd, R, r, n):
    import ctypes as C
    out = np.zeros(7 * R + 6, dtype=np.int64)
    b = np.ascontiguousarray(bnd, dtype=np.int32)
    L.hh_exchange_plan.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p]
    L.hh_exchange_plan(b.ctypes.data_as(C.c_void_p), R, r, n, out.ctypes.data_as(C.c_void_p))
    per = out[:7 * R].reshape(R, 7)
    t = out[7 * R:]
    return per, dict(fs=t[0], fe=t[1], remote_send=t[2], remote_recv=t[3], total_remote=t[4], fits=bool(t[5]))


@pytest.mark.parametrize("R", [2, 3, 4, 8])
def test_exchange_plan_is_consistent_between_senders_and_receivers(R):
    """csrc/ws_exchange.h, the host arithmetic of the sharded exchange: for random partitions of the global
    slots into per-rank produced ranges, what every sender derives (piece sizes, where a piece lands in the receiver's
    spare rows or back buffer: the direct NVLink exchange writes there without asking) must be what every receiver
    derives for itself — no overlap, no gap, the same `fits` / migrant totals on every rank."""
    import hostlib
    L = hostlib.lib()
    rng = np.random.default_rng(R)
    for trial in range(300):
        n = int(rng.integers(R * 5000, R * 400_000))
        kind = trial % 4
        if kind == 0:      # near-uniform weights: bounds close to the shard boundaries
            cuts = np.array([n * q // R for q in range(1, R)]) + rng.integers(-50, 50, R - 1)
        elif kind == 1:    # arbitrary masses
            cuts = np.sort(rng.integers(0, n + 1, R - 1))
        elif kind == 2:    # one rank holds (almost) everything
            cuts = np.sort(np.concatenate([rng.integers(0, 20, R - 1 - (R - 1) // 2), n - rng.integers(0, 20, (R - 1) // 2)]))
        else:              # some ranks produce nothing
            cuts = np.sort(rng.choice([0, n // 3, n // 2, n], R - 1))
        cuts = np.clip(np.sort(cuts), 0, n)
        edges = np.concatenate([[0], cuts, [n]]).astype(np.int64)
        bnd = np.stack([edges[:-1], edges[1:]], axis=1).reshape(-1)
        lo = [n * d // R for d in range(R + 1)]
        plans = [_plan(L, bnd, R, r, n) for r in range(R)]
        fits0, tot0 = plans[0][1]["fits"], plans[0][1]["total_remote"]
        for r, (per, t) in enumerate(plans):
            assert t["fits"] == fits0 and t["total_remote"] == tot0            # every rank decides alike
            assert per[:, 1].sum() == t["fe"] - t["fs"]                        # every produced slot goes somewhere
            assert per[:, 3].sum() == lo[r + 1] - lo[r]                        # every one of my slots comes from somewhere
            assert t["remote_send"] == per[:, 1].sum() - per[r, 1] and t["remote_recv"] == per[:, 3].sum() - per[r, 3]
        assert tot0 == sum(t["remote_recv"] for _, t in plans) == sum(t["remote_send"] for _, t in plans)
        for d in range(R):                                                      # receiver d
            per_d, t_d = plans[d]
            n_d = lo[d + 1] - lo[d]
            rows_lazy, rows_eager = [], []
            for q in range(R):                                                  # sender q
                per_q, _ = plans[q]
                assert per_q[d, 1] == per_d[q, 3]                               # piece size: sender == receiver
                if q == d or per_q[d, 1] == 0:
                    continue
                assert per_q[d, 6] == per_d[q, 2]                               # eager: lands at the receiver's recv_off
                assert per_q[d, 5] == n_d + per_d[q, 4]                         # lazy: lands at the receiver's spare_pos
                rows_lazy.append((per_q[d, 5], per_q[d, 5] + per_q[d, 1]))
                rows_eager.append((per_q[d, 6], per_q[d, 6] + per_q[d, 1]))
            if per_d[d, 3] > 0:
                rows_eager.append((per_d[d, 2], per_d[d, 2] + per_d[d, 3]))       # the offspring that stay
            rows_eager.sort()
            assert rows_eager[0][0] == 0 and rows_eager[-1][1] == n_d
            assert all(a[1] == b[0] for a, b in zip(rows_eager, rows_eager[1:]))  # eager pieces tile [0, n_d)
            rows_lazy.sort()
            assert all(a[1] <= b[0] for a, b in zip(rows_lazy, rows_lazy[1:]))    # spare-row pieces do not overlap
            if rows_lazy:
                assert rows_lazy[0][0] == n_d
                if fits0:
                    assert rows_lazy[-1][1] <= n_d + max(4096, n_d // 32)         # ... and stay inside the spare rows


def test_spare_ring_never_hands_out_a_row_twice():
    """The spare-row ring of the sharded genealogy (csrc/ws_exchange.h) against a brute-force occupancy map: random
    allocations (one per resampling event) and releases of the oldest events; a region is contiguous, inside the
    capacity, disjoint from every live region, and an allocation only fails when the rows behind the newest region
    and in front of the oldest one are both too short."""
    import ctypes as C
    import hostlib
    L = hostlib.lib()
    L.hh_spare_ring.argtypes = [C.c_int64, C.c_void_p, C.c_int, C.c_void_p]
    rng = np.random.default_rng(11)
    for trial in range(40):
        cap = int(rng.integers(8, 200))
        ops, ev, oldest = [], 0, 0
        for _ in range(120):
            if rng.random() < 0.7 or oldest > ev:
                ev += 1
                ops.append((0, int(rng.integers(0, max(2, cap // 3))), ev))
            else:
                oldest = int(rng.integers(oldest, ev + 1))
                ops.append((1, oldest, 0))
        a = np.asarray(ops, dtype=np.int64)
        out = np.zeros(len(ops), dtype=np.int64)
        L.hh_spare_ring(cap, a.ctypes.data_as(C.c_void_p), len(ops), out.ctypes.data_as(C.c_void_p))
        live = {}                      # event -> (start, cnt)
        failed = 0
        for (kind, x, y), o in zip(ops, out):
            if kind == 0:
                if o < 0:
                    failed += 1
                    occ = np.zeros(cap, dtype=bool)
                    for s, c in live.values():
                        occ[s:s + c] = True
                    # no contiguous free run of x rows may exist at the two places the ring looks at; in particular
                    # an empty ring never refuses a request that fits the capacity
                    assert x > cap or occ.any()
                    continue
                assert 0 <= o and o + x <= cap
                for s, c in live.values():
                    assert o + x <= s or s + c <= o or x == 0 or c == 0, (trial, o, x, s, c)
                live[y] = (int(o), x)
            else:
                live = {e: v for e, v in live.items() if e > x}
                assert o == len(live)
        assert failed < len(ops)


def test_mailbox_protocol_double_buffering_holds_under_skew():
    """The exchange protocol of csrc/ws_mailbox.cuh (sequence-tagged 8-byte words, mailboxes double-buffered on the parity
    of the sequence number) restated over host threads (tests/host/harness.cpp: hh_mailbox_protocol): R ranks run
    hundreds of all-to-all exchanges back to back with random pauses, one of them reading slowly — every payload must
    be the one its sender wrote for THAT exchange and nobody may wait for ever.  Negative control: with one buffer the
    fast ranks overwrite words the slow reader has not consumed yet, and it waits for a sequence number that is gone."""
    import ctypes as C
    import hostlib
    L = hostlib.lib()
    L.hh_mailbox_protocol.restype = C.c_int64
    L.hh_mailbox_protocol.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int64, C.c_int]
    for R, n_words, slow in ((2, 3, -1), (2, 31, 0), (4, 3, 1), (8, 1, 3), (8, 5, -1)):
        for seed in (1, 2, 3):
            assert L.hh_mailbox_protocol(R, 300, n_words, 2, seed, 50_000_000, slow) == 0, (R, n_words, slow, seed)
    # one buffer: the slow reader is overtaken (detected as a wait beyond the limit, or — never on an atomic word — a torn payload)
    assert any(L.hh_mailbox_protocol(4, 300, 4, 1, seed, 200_000, 0) != 0 for seed in (1, 2, 3))
