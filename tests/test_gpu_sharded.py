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


def _worker(rank, world, port, n, T, ess, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import wsb200 as ws
    st = ws.sharded_state(n, device=rank, seed=77, ess_perc_min=ess)
    ws.run(ws.model(MODEL)(_obs(T)), st)
    le = ws.log_evidence(st)
    mx = ws.E(lambda x: x[0], st)
    mig = ctypes.c_int64()
    st.store._call("ws_get_migrated", ctypes.byref(mig))
    q.put((rank, st["x"], st["v"], st.weights, le, mx, st.stats()["resamples_done"], mig.value))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,ess", [(2, 1.0), (2, 0.5)])
def test_sharded_equals_single_gpu(world, ess):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    import wsb200 as ws
    n, T = 200_003, 12
    single = ws.SMCState(n, device=0, seed=77, ess_perc_min=ess)
    ws.run(ws.model(MODEL)(_obs(T)), single)
    x1, v1, w1 = single["x"], single["v"], single.weights
    le1, mx1 = ws.log_evidence(single), ws.E(lambda x: x[0], single)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, 29741 + int(ess * 10), n, T, ess, q)) for r in range(world)]
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
    print(f"world={world} ess={ess}: {int(bad.sum())} of {n} particles differ; migrated {[o[7] for o in out]}")
    assert bad.sum() <= 5
    np.testing.assert_allclose(wsum, w1, rtol=1e-9, atol=1e-9)
    for o in out:
        assert abs(o[4] - le1) <= 1e-10 * abs(le1)
        assert abs(o[5] - mx1) <= 1e-9 * (1 + abs(mx1))
        assert o[6] == single.stats()["resamples_done"] and o[6] >= 3
    assert sum(o[7] for o in out) > 0, "some offspring must have crossed the shard boundary"
