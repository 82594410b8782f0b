"""Sharded configs of BASELINE.json (one rank per GPU, launched with torch.distributed.run):
  C4  examples/eight_schools.jl, N = 1e7 particles in total, sharded over the ranks (BASELINE configs[3])
  C5  resampling microbenchmark, 1e8 particles PER GPU, one exact global Resample (BASELINE configs[4])
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node R --master-addr 127.0.0.1 benchmarks/run_sharded.py [c4] [c5]
Rank 0 prints one JSON line per configuration; times are device-side, max over ranks."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist

import models
import wsb200 as ws

HBM = 6542.1  # GB/s, MEASURED_PEAKS.json


def setup():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def make_state(n_global, world, local, **kw):
    if world > 1:
        return ws.sharded_state(n_global, device=local, **kw)
    return ws.SMCState(n_global, device=local, **kw)


def max_over_ranks(x, world, local):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=torch.device("cuda", local))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x, world, local):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=torch.device("cuda", local))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def emit(rank, **kw):
    if rank == 0:
        print(json.dumps(kw), flush=True)


def c4(rank, world, local):
    n = 10_000_000
    for J, y, sig, tag in ((8, models.SCHOOLS_Y, models.SCHOOLS_SIGMA, "literal data, J=8"),):
        make = ws.model(models.SCHOOLS)
        st = make_state(n, world, local, ess_perc_min=0.5, seed=1)
        ws.run(make(J, y, sig), st)      # warm-up (NCCL channels, allocations)
        del st
        st = make_state(n, world, local, ess_perc_min=0.5, seed=2)
        st.sync()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        ws.run(make(J, y, sig), st)
        mu, tau = ws.E(lambda μ: μ, st), ws.E(lambda τ: τ, st)
        st.sync()
        dt = max_over_ranks(time.perf_counter() - t0, world, local)
        s = st.stats()
        emit(rank, config=f"C4 examples/eight_schools.jl N={n} over {world} GPU(s), {tag}, diversity-gated autoRW moves",
             seconds=dt, particle_updates_per_sec=n * J / dt, mu=mu, tau=tau, log_evidence=ws.log_evidence(st),
             resamples=s["resamples_done"], moves=s["moves_run"], n_gpus=world)
        del st
    # synthetic J = 512, 1e6 particles in total
    J = 512
    rng = np.random.default_rng(1)
    sig = rng.uniform(9, 18, J)
    th = 4.0 + 3.0 * rng.standard_normal(J)
    y = th + sig * rng.standard_normal(J)
    n2 = 1_000_000
    st = make_state(n2, world, local, ess_perc_min=0.5, seed=3)
    st.sync()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    ws.run(ws.model(models.SCHOOLS)(J, list(y), list(sig)), st)
    mu = ws.E(lambda μ: μ, st)
    st.sync()
    dt = max_over_ranks(time.perf_counter() - t0, world, local)
    s = st.stats()
    emit(rank, config=f"C4 synthetic J={J} N={n2} over {world} GPU(s)", seconds=dt, particle_updates_per_sec=n2 * J / dt, mu=mu,
         log_evidence=ws.log_evidence(st), resamples=s["resamples_done"], moves=s["moves_run"], n_gpus=world)


def c5(rank, world, local):
    """logw = s z, payload P planes, ONE exact global Resample (allgather of the weight mass, global CDF offsets, search on
    the global slot grid, NCCL migration of the offspring that cross a shard boundary, eager gather)."""
    import ctypes as C
    per_gpu = 100_000_000
    cases = [(P, label, s, "stratified") for P in (1, 6) for label, s in (("s=0.5", 0.5), ("s=2", 2.0), ("s=4", 4.0))]
    cases += [(6, "s=2", 2.0, "systematic"), (6, "s=2", 2.0, "multinomial"), (16, "s=2", 2.0, "stratified")]
    if True:
        for P, label, s, scheme in cases:
            st = make_state(per_gpu * world, world, local, ess_perc_min=float("inf"), seed=0x5EED, resampler=scheme)
            store = st.store
            store._call("ws_set_lazy_gather", 0)
            for p in range(P):
                ws.Sample(f"p{p}", "Normal", (0.0, 1.0)).apply(st)
            ws.Sample("z", "Normal", (0.0, 1.0)).apply(st)
            store._call("ws_set_timing", 1)
            times, mig = [], []
            for rep in range(4):
                ws.Weight(None, (ws.col("z") * s,)).apply(st)
                st.sync()
                m0 = C.c_int64()
                store._call("ws_get_migrated", C.byref(m0))
                if world > 1:
                    dist.barrier()
                t0 = time.perf_counter()
                r = ws.Resample()
                r.apply(st)
                st.sync()
                dt = max_over_ranks(time.perf_counter() - t0, world, local)
                m1 = C.c_int64()
                store._call("ws_get_migrated", C.byref(m1))
                if rep > 0:
                    times.append(dt * 1e3)
                    mig.append(m1.value - m0.value)
                ess = r.last.ess_perc
                ws.Sample("z", "Normal", (0.0, 1.0)).apply(st)
            ms = float(np.median(times))
            planes = P + 1
            alg = 32 + 16 * planes
            n = per_gpu * world
            emit(rank, config=f"C5 sharded resample N={per_gpu} per GPU x {world}, payload={P}+1 planes {label} {scheme}",
                 ms=ms, ess_perc=ess, particles_per_sec=n / (ms * 1e-3), alg_bytes_per_particle=alg,
                 achieved_gbs_per_gpu=alg * per_gpu / (ms * 1e-3) / 1e9, hbm_frac_per_gpu=alg * per_gpu / (ms * 1e-3) / 1e9 / HBM,
                 migrated_particles_rank0=float(np.mean(mig)), migrated_bytes_rank0=float(np.mean(mig)) * 8 * planes,
                 n_gpus=world, timing="host wall clock around Resample (sync on both sides), max over ranks")
            del st


def c5skew(rank, world, local):
    """The migration path under load: log-weights fall with the GLOBAL particle index (logw = -lambda i / N), so the
    low ranks hold almost all of the mass and most offspring must cross shard boundaries over NVLink (SURVEY 8e:
    worst-case egress of one heavy rank ~ N 8P (R-1)/R bytes).  Reported against the measured NVLink peer bandwidth
    (770 GB/s per direction, B200_PROFILING.md)."""
    import ctypes as C
    if world < 2:
        return
    NVLINK = 770.0
    per_gpu, P = 100_000_000, 6
    n = per_gpu * world
    for lam in (2.0, 10.0):
        st = make_state(n, world, local, ess_perc_min=float("inf"), seed=0x5EED)
        store = st.store
        for p in range(P):
            ws.Sample(f"p{p}", "Normal", (0.0, 1.0)).apply(st)
        lo, hi = ws.shard_bounds(n, rank, world)
        g = (np.arange(lo, hi, dtype=np.float64)) / float(n)
        store._call("ws_set_timing", 1)
        times, mig, sent_direct = [], [], []
        for rep in range(3):
            store.setcol("g", g)
            ws.Weight(None, (ws.col("g") * (-lam),)).apply(st)
            st.sync()
            m0, s0 = C.c_int64(), C.c_int64()
            store._call("ws_get_migrated", C.byref(m0))
            store._call("ws_get_pushed", C.byref(s0))
            dist.barrier()
            t0 = time.perf_counter()
            r = ws.Resample()
            r.apply(st)
            st.sync()
            dt = max_over_ranks(time.perf_counter() - t0, world, local)
            m1, s1 = C.c_int64(), C.c_int64()
            store._call("ws_get_migrated", C.byref(m1))
            store._call("ws_get_pushed", C.byref(s1))
            if rep > 0:
                times.append(dt * 1e3)
                mig.append(m1.value - m0.value)
                sent_direct.append(s1.value - s0.value)
        ms = float(np.median(times))
        planes = P + 1
        sent = float(np.mean(mig))
        sent_max = max_over_ranks(sent, world, local)     # busiest receiver (ingress)
        sent_sum = sum_over_ranks(sent, world, local)     # all migrants
        egress_max = max_over_ranks(float(np.mean(sent_direct)), world, local)  # busiest sender (direct exchange only; 0 on the NCCL path)
        emit(rank, config=f"C5 rank-skewed resample (logw = -{lam} i/N), N={per_gpu} per GPU x {world}, payload={P}+1 planes",
             ms=ms, ess_perc=r.last.ess_perc, particles_per_sec=n / (ms * 1e-3),
             received_particles_max_rank=sent_max, received_bytes_max_rank=sent_max * 8 * planes,
             migrated_particles_total=sent_sum, migrated_bytes_total=sent_sum * 8 * planes,
             nvlink_ingress_gbs_max_rank=sent_max * 8 * planes / (ms * 1e-3) / 1e9,
             sent_particles_max_rank=egress_max, nvlink_egress_gbs_max_rank=egress_max * 8 * planes / (ms * 1e-3) / 1e9,
             nvlink_total_gbs=sent_sum * 8 * planes / (ms * 1e-3) / 1e9,
             nvlink_frac_of_770=max(sent_max, egress_max) * 8 * planes / (ms * 1e-3) / 1e9 / NVLINK, n_gpus=world,
             timing="host wall clock around Resample incl. migration and gather (sync on both sides), max over ranks; "
                    "the NVLink figure divides the busiest rank's migrated bytes by the WHOLE step time")
        del st


def c2hist(rank, world, local):
    """examples/2D_ssm.jl VERBATIM (x{t} history kept, 2 (T + 3) planes) as ONE sharded filter: with the sharded genealogy a
    history column is never gathered again (migrating offspring are traced by their sender); without it every stale plane
    is gathered before every event, as the reference's resample! does (src/stores.jl:105-121)."""
    import ctypes as C
    rng = np.random.default_rng(42)
    T = 400
    obs = [rng.standard_normal(2) * 0.5 + np.array([t, 0.0]) for t in range(T)]
    quick = os.environ.get("WSB200_C2HIST_QUICK") == "1"   # one configuration (profiling runs)
    for n, Tn_off in ((1_000_000, 400), (10_000_000, 100)):
        for on in ((True,) if quick else (True, False)):
            Tn = T if on else Tn_off   # the eager variant is quadratic in T: shorter run at the large size, per-step figure quoted
            st = make_state(n, world, local, ess_perc_min=1.0, seed=1)
            st.set_genealogy(on)
            root = ws.model(models.SSM2D)(obs[:Tn])
            st.sync()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            ws.run(root, st)
            le = ws.log_evidence(st)
            st.sync()
            dt = max_over_ranks(time.perf_counter() - t0, world, local)
            t1 = time.perf_counter()
            x_mid = st[f"x_{Tn // 2}"]          # a column ~T/2 events behind
            dt_read = max_over_ranks(time.perf_counter() - t1, world, local)
            traced, mig = C.c_int64(), C.c_int64()
            st.store._call("ws_get_traced_pushes", C.byref(traced))
            st.store._call("ws_get_migrated", C.byref(mig))
            emit(rank, config=f"C2-history examples/2D_ssm.jl verbatim N={n} T={Tn} over {world} GPU(s), genealogy={'on' if on else 'off'}",
                 seconds=dt, particle_updates_per_sec=n * Tn / dt, ms_per_step=1e3 * dt / Tn, log_evidence=le, genealogy=st.genealogy(),
                 read_mid_column_seconds=dt_read, mid_column_mean_rank0=float(x_mid[:, 0].mean()), columns=len(st.store.colnames()),
                 traced_values_rank0=traced.value, migrated_rank0=mig.value, n_gpus=world)
            del st


if __name__ == "__main__":
    rank, world, local = setup()
    which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["c4", "c5", "c5skew"]
    for w in which:
        {"c4": c4, "c5": c5, "c5skew": c5skew, "c2hist": c2hist}[w](rank, world, local)
    if world > 1:
        dist.destroy_process_group()
