"""Sharded (multi-GPU) global resampling: R ranks must reproduce the single-GPU run.  Needs >= 2 GPUs
(run with `gpurun --gpus 2`); skipped on a 1-GPU box."""
import ctypes
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import wsb200
    n = ctypes.c_int()
    wsb200.load().ws_device_count(ctypes.byref(n))
    return n.value


MODEL = '''
@model function ssm(obs)
    I2 = [1.0 0.0; 0.0 1.0]
    x .= [0.0, 0.0]
    v .= [1.0, 0.0]
    for o in obs
        x .= x + v
        dv ~ MvNormal([0.0, 0.0], 0.1 * I2)
        v .= v + dv
        o => MvNormal(x, 0.5 * I2)
    end
end
'''


def _obs(T):
    rng = np.random.default_rng(3)
    return [np.array([t, 0.0]) + 0.7 * rng.standard_normal(2) for t in range(T)]


def _worker(rank, world, port, n, T, ess, q, resampler="stratified", mailbox="1"):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WSB200_MAILBOX=mailbox)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import wsb200 as ws
    st = ws.sharded_state(n, device=rank, seed=77, ess_perc_min=ess, resampler=resampler)
    ws.run(ws.model(MODEL)(_obs(T)), st)
    le = ws.log_evidence(st)
    mx = ws.E(lambda x: x[0], st)
    mig, mbx = ctypes.c_int64(), ctypes.c_int64()
    st.store._call("ws_get_migrated", ctypes.byref(mig))
    st.store._call("ws_get_mailbox_exchanges", ctypes.byref(mbx))
    q.put((rank, st["x"], st["v"], st.weights, le, mx, st.stats()["resamples_done"], mig.value, mbx.value))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,ess,resampler,mailbox", [(2, 1.0, "stratified", "1"), (2, 0.5, "stratified", "1"), (4, 1.0, "stratified", "1"),
                                                         (2, 1.0, "systematic", "1"), (2, 1.0, "multinomial", "1"), (4, 0.5, "multinomial", "1"),
                                                         (2, 0.5, "stratified", "0"), (4, 1.0, "systematic", "0")])
def test_sharded_equals_single_gpu(world, ess, resampler, mailbox):
    """mailbox = "1": the step's small exchanges are made by the kernels themselves through peer-mapped mailboxes
    (csrc/ws_mailbox.cuh); "0": the NCCL collectives they replace.  Both must reproduce the single-GPU run."""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    import wsb200 as ws
    n, T = 200_003, 12
    single = ws.SMCState(n, device=0, seed=77, ess_perc_min=ess, resampler=resampler)
    ws.run(ws.model(MODEL)(_obs(T)), single)
    x1, v1, w1 = single["x"], single["v"], single.weights
    le1, mx1 = ws.log_evidence(single), ws.E(lambda x: x[0], single)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29741 + int(ess * 10) + 20 * ["stratified", "systematic", "multinomial"].index(resampler) + world + 100 * (mailbox == "0")
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, T, ess, q, resampler, mailbox)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted((q.get(timeout=300) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    xs = np.concatenate([o[1] for o in out])
    vs = np.concatenate([o[2] for o in out])
    wsum = np.concatenate([o[3] for o in out])
    assert xs.shape == x1.shape
    # same Philox counters (global indices) and an order-independent fixed-point CDF: the sharded run is
    # the single-GPU run, up to the rounding of the (m, S) reduction
    bad = (np.abs(xs - x1) > 1e-9 * (1 + np.abs(x1))).any(axis=1) | (np.abs(vs - v1) > 1e-9 * (1 + np.abs(v1))).any(axis=1)
    print(f"world={world} ess={ess} {resampler} mailbox={mailbox}: {int(bad.sum())} of {n} particles differ; migrated {[o[7] for o in out]}; "
          f"mailbox exchanges {[o[8] for o in out]}")
    # every Resample statement: one exchange for the decision; a step that fires: two more (+ the barrier behind pushed offspring)
    assert all((o[8] >= T + 2 * o[6]) if mailbox == "1" else (o[8] == 0) for o in out)
    assert bad.sum() <= 5
    np.testing.assert_allclose(wsum, w1, rtol=1e-9, atol=1e-9)
    for o in out:
        assert abs(o[4] - le1) <= 1e-10 * abs(le1)
        assert abs(o[5] - mx1) <= 1e-9 * (1 + abs(mx1))
        assert o[6] == single.stats()["resamples_done"] and o[6] >= 3
    assert sum(o[7] for o in out) > 0, "some offspring must have crossed the shard boundary"


def _worker_schools(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import models
    import wsb200 as ws
    st = ws.sharded_state(n, device=rank, seed=5, ess_perc_min=0.5)
    ws.run(ws.model(models.SCHOOLS)(8, models.SCHOOLS_Y, models.SCHOOLS_SIGMA), st)
    div = ws.marginal_diversity(st.store, ["μ"])
    out = (rank, ws.log_evidence(st), ws.E(lambda μ: μ, st), ws.E(lambda τ: τ, st), div, st.stats()["moves_run"],
           st.stats()["resamples_done"], st["μ"], st["θ"])
    q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_eight_schools_with_diversity_gated_moves():
    """BASELINE configs[3] shape: diversity-gated autoRW moves on a sharded state (exact cross-rank distinct
    count, all-reduced autoRW moments) reproduce the single-GPU run."""
    world = 2
    if _ngpu() < world:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import models
    import wsb200 as ws
    n = 150_001
    single = ws.SMCState(n, device=0, seed=5, ess_perc_min=0.5)
    ws.run(ws.model(models.SCHOOLS)(8, models.SCHOOLS_Y, models.SCHOOLS_SIGMA), single)
    le1, mu1, tau1 = ws.log_evidence(single), ws.E(lambda μ: μ, single), ws.E(lambda τ: τ, single)
    div1 = ws.marginal_diversity(single.store, ["μ"])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_schools, args=(r, world, 29761, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted((q.get(timeout=300) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    mus = np.concatenate([o[7] for o in out])
    th = np.concatenate([o[8] for o in out])
    differ = int((np.abs(mus - single["μ"]) > 1e-9 * (1 + np.abs(mus))).sum())
    print(f"eight schools sharded: {differ} of {n} particles differ in mu; diversity {out[0][4]} vs {div1}; "
          f"moves {out[0][5]} vs {single.stats()['moves_run']}")
    assert th.shape == (n, 8)
    for o in out:
        assert abs(o[1] - le1) < 1e-9 * abs(le1)
        assert abs(o[2] - mu1) < 1e-3 and abs(o[3] - tau1) < 1e-3
        assert o[4] == out[0][4]                       # every rank sees the same global diversity
        assert o[5] == single.stats()["moves_run"] and o[6] == single.stats()["resamples_done"]
    assert abs(out[0][4] - div1) < 5e-4
    assert differ < 0.002 * n


def _linreg_data():
    rng = np.random.default_rng(8)
    xs = rng.uniform(0, 10, 40)
    return list(xs), list(1 - 0.5 * xs + rng.standard_normal(40))


def _worker_linreg(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import models
    import wsb200 as ws
    st = ws.sharded_state(n, device=rank, seed=9, ess_perc_min=0.5)
    ws.run(ws.model(models.LINREG)(*_linreg_data()), st)
    q.put((rank, ws.log_evidence(st), st.stats()["moves_run"], st.stats()["resamples_done"], st["α"], st["β"]))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_linear_regression_with_moves_after_resampling():
    """BASELINE configs[2] shape (examples/linear_regression.jl: `y => Normal(alpha + beta x, 1)`; `if resampled` autoRW
    moves) on a sharded state.  The single-GPU run takes its observations in speculative blocks (ws_exec_spec); a sharded
    state refuses them and runs element by element — both must give the same particles."""
    world = 2
    if _ngpu() < world:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import models
    import wsb200 as ws
    n = 120_001
    single = ws.SMCState(n, device=0, seed=9, ess_perc_min=0.5)
    ws.run(ws.model(models.LINREG)(*_linreg_data()), single)
    le1 = ws.log_evidence(single)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_linreg, args=(r, world, 29771, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted((q.get(timeout=300) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    al = np.concatenate([o[4] for o in out])
    be = np.concatenate([o[5] for o in out])
    differ = int(((np.abs(al - single["α"]) > 1e-9 * (1 + np.abs(al))) | (np.abs(be - single["β"]) > 1e-9 * (1 + np.abs(be)))).sum())
    print(f"linear regression sharded: {differ} of {n} particles differ; moves {out[0][2]} vs {single.stats()['moves_run']}; "
          f"resamples {out[0][3]} vs {single.stats()['resamples_done']}")
    for o in out:
        assert abs(o[1] - le1) < 1e-9 * abs(le1)
        assert o[2] == single.stats()["moves_run"] > 0 and o[3] == single.stats()["resamples_done"] > 0
    assert differ < 0.002 * n      # (all-reduced autoRW moments: the proposal scale agrees to rounding, near-tie accepts may flip)


def _worker_describe(rank, world, port, n, T, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import wsb200 as ws
    st = ws.sharded_state(n, device=rank, seed=77, ess_perc_min=0.5)
    ws.run(ws.model(MODEL)(_obs(T)), st)
    d = ws.describe(st)
    q.put((rank, d.to_dict("list")))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_describe_equals_single_gpu():
    """describe (src/utils.jl:183-289) on a sharded state: shard partials merged across ranks, the median's radix select
    over all-reduced fixed-point histograms.  Every rank must report the single-GPU numbers (the particles are the
    same, bit for bit; sums differ by their order only, the median and the extrema not at all)."""
    world = 2
    if _ngpu() < world:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import wsb200 as ws
    n, T = 100_003, 9          # ends on a weighted state (ESS gate 0.5): the weights matter
    single = ws.SMCState(n, device=0, seed=77, ess_perc_min=0.5)
    ws.run(ws.model(MODEL)(_obs(T)), single)
    d1 = ws.describe(single).to_dict("list")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_describe, args=(r, world, 29771, n, T, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted((q.get(timeout=300) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert out[0][1]["variable"] == d1["variable"]
    for _, d in out:
        for f in ("mean", "std", "ess"):
            for a, b in zip(d[f], d1[f]):
                np.testing.assert_allclose(np.asarray(a, dtype=float), np.asarray(b, dtype=float), rtol=1e-10, atol=1e-12)
        for f in ("median", "min", "max"):
            for a, b in zip(d[f], d1[f]):
                assert np.array_equal(np.asarray(a, dtype=float), np.asarray(b, dtype=float)), (f, a, b)
        assert d["hist"] == d1["hist"]
    # ... and the ranks agree with each other to the last bit
    assert all(np.array_equal(np.asarray(a, dtype=float), np.asarray(b, dtype=float))
               for f in ("mean", "median", "std", "min", "max") for a, b in zip(out[0][1][f], out[1][1][f]))


def _skew_run(ws, st, n, lo, hi, lam):
    """p ~ N(0,1); g = global index / n (uploaded); logw = -lam g; Resample; q .= p + g (reads through the ancestors)."""
    ws.Sample("p", "Normal", (0.0, 1.0)).apply(st)
    st.store.setcol("g", np.arange(lo, hi, dtype=np.float64) / n)
    ws.Weight(None, (ws.col("g") * (-lam),)).apply(st)
    r = ws.Resample()
    r.apply(st)
    ws.Assign("q", ws.col("p") + ws.col("g")).apply(st)
    return r.last.ess_perc


def _worker_skew(rank, world, port, n, lam, push_min, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WSB200_PUSH_MIN=str(push_min))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import wsb200 as ws
    st = ws.sharded_state(n, device=rank, seed=77, ess_perc_min=float("inf"))
    lo, hi = ws.shard_bounds(n, rank, world)
    ess = _skew_run(ws, st, n, lo, hi, lam)
    pushed, mig = ctypes.c_int64(), ctypes.c_int64()
    st.store._call("ws_get_pushed", ctypes.byref(pushed))
    st.store._call("ws_get_migrated", ctypes.byref(mig))
    q.put((rank, st["p"], st["g"], st["q"], ess, pushed.value, mig.value))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("lam,push_min,mode", [(10.0, 1000, "eager"), (0.1, 1000, "lazy"), (10.0, 1 << 40, "nccl")])
def test_sharded_direct_exchange_equals_single_gpu(lam, push_min, mode):
    """Migration under load: log-weights fall with the global index, so rank 0 produces offspring for rank 1.  The direct
    exchange (gather kernels writing into the peer's planes over NVLink, cudaIpc mappings) must give the single-GPU
    result particle for particle — with heavy migration (eager: final slots of the peer's back planes), with light
    migration (lazy: the peer's spare rows, read through the ancestors by the next pass) — and so must the
    stage + ncclSend/ncclRecv path it replaces."""
    world = 2
    if _ngpu() < world:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import wsb200 as ws
    n = 400_001
    single = ws.SMCState(n, device=0, seed=77, ess_perc_min=float("inf"))
    ess1 = _skew_run(ws, single, n, 0, n, lam)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_skew, args=(r, world, 29781, n, lam, push_min, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted((q.get(timeout=300) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for k, name in ((1, "p"), (2, "g"), (3, "q")):
        got = np.concatenate([o[k] for o in out])
        assert np.array_equal(got, single[name]), (mode, name, int((got != single[name]).sum()))
    assert all(abs(o[4] - ess1) < 1e-12 for o in out)
    migrated = sum(o[6] for o in out)
    pushed = sum(o[5] for o in out)
    print(f"{mode}: {migrated} of {n} particles migrated, {pushed} written directly into the peer")
    if mode == "nccl":
        assert pushed == 0 and migrated > 100_000
    else:
        assert pushed == migrated > 1000
        assert (migrated > n // 64) == (mode == "eager")


def _worker_history(rank, world, port, n, T, genealogy, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WSB200_GENEALOGY=genealogy)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import models
    import wsb200 as ws
    rng = np.random.default_rng(7)
    obs = list(np.cumsum(0.3 * rng.standard_normal(T)) + rng.standard_normal(T))
    st = ws.sharded_state(n, device=rank, seed=31, ess_perc_min=1.0)
    st.store._call("ws_set_timing", 1)
    ws.run(ws.model(models.SSM1D)(obs), st)
    kt = st.kernel_times()
    g = st.genealogy()
    traced, mig = ctypes.c_int64(), ctypes.c_int64()
    st.store._call("ws_get_traced_pushes", ctypes.byref(traced))
    st.store._call("ws_get_migrated", ctypes.byref(mig))
    names = st.store.colnames()
    # history columns read newest first, oldest first and in between
    order = names[::-1][:5] + names[:5] + names[5:-5]
    cols = {name: st[name] for name in order}
    q.put((rank, cols, st.weights, ws.log_evidence(st), g, traced.value, mig.value, kt["gather"]["launches"], st.stats()["resamples_done"]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n,T", [(2, 200_003, 40), (2, 60_001, 90), (4, 100_003, 40)])
def test_sharded_genealogy_keeps_history_columns_in_place(world, n, T):
    """examples/1D_ssm.jl shape on a sharded state: x{t} is written once and not read again, so it must not be gathered
    at every later resampling event (src/stores.jl:105-121 does exactly that).  Planes that are behind keep their order
    over the events of a sharded run too: offspring that change GPU are traced through the retained ancestor vectors
    by their sender, land in spare rows handed out event by event, and a read composes the vectors as on one GPU.
    Every column must equal the single-GPU run (itself checked against the oracle's eager resample! in
    test_gpu_genealogy.py).  The second case has 4 096 spare rows per rank and ~90 events: the rings fill up and the
    oldest planes are brought up to date to release them."""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import models
    import wsb200 as ws
    rng = np.random.default_rng(7)
    obs = list(np.cumsum(0.3 * rng.standard_normal(T)) + rng.standard_normal(T))
    single = ws.SMCState(n, device=0, seed=31, ess_perc_min=1.0)
    ws.run(ws.model(models.SSM1D)(obs), single)
    names = single.store.colnames()
    ref_cols = {name: single[name] for name in names}
    res = {}
    for genealogy in ("1", "0"):
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = 29801 + world + 10 * (genealogy == "0") + (T % 7)
        procs = [ctx.Process(target=_worker_history, args=(r, world, port, n, T, genealogy, q)) for r in range(world)]
        for p in procs:
            p.start()
        out = sorted((q.get(timeout=600) for _ in range(world)), key=lambda t: t[0])
        for p in procs:
            p.join(timeout=120)
            assert p.exitcode == 0
        res[genealogy] = out
        differing = 0
        for name in names:
            got = np.concatenate([o[1][name] for o in out])
            bad = np.abs(got - ref_cols[name]) > 1e-9 * (1 + np.abs(ref_cols[name]))
            differing = max(differing, int(bad.sum()))
        print(f"world={world} n={n} T={T} genealogy={genealogy}: worst column has {differing} of {n} particles differing; "
              f"vectors kept {[o[4]['vectors'] for o in out]}, traced values {[o[5] for o in out]}, migrated {[o[6] for o in out]}, "
              f"gather launches {[o[7] for o in out]}")
        assert differing <= 5
        np.testing.assert_allclose(np.concatenate([o[2] for o in out]), single.weights, rtol=1e-9, atol=1e-12)
        for o in out:
            assert abs(o[3] - ws.log_evidence(single)) <= 1e-10 * abs(ws.log_evidence(single))
            assert o[8] == single.stats()["resamples_done"]
    on, off = res["1"], res["0"]
    assert all(o[5] > 0 for o in on) and all(o[5] == 0 for o in off)          # offspring were traced for planes that were behind
    assert all(o[4]["vectors"] > 3 for o in on)                                # ... whose ancestor vectors were kept
    # without the genealogy every stale plane is gathered before every event: O(T^2) plane gathers against O(T)
    assert sum(o[7] for o in on) * 2 < sum(o[7] for o in off)
