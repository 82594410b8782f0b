#!/usr/bin/env python
"""Throughput of the BASELINE.json configs other than the headline one (C1, LGSSM-1D, C3, C4, C5), one JSON
line each.  `bench.py` is the contract benchmark (C2); this script gives the context numbers quoted in the
README and keeps the CPU restatement (oracle/ws_oracle.c) beside the LGSSM-1D number, the model the
reference's own published numbers are for (benchmarks/ssm/results/grid_results.csv).

    python benchmarks/run_configs.py [c1] [lgssm] [c3] [c4] [c5] [--quick]
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import wsb200 as ws  # noqa: E402
import models  # noqa: E402

HBM = 6542.1
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    HBM = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
QUICK = "--quick" in sys.argv


def timed_run(build_args, src, n, ess, seed=1, record_tape=True, **state_kw):
    state = ws.SMCState(n, ess_perc_min=ess, seed=seed, **state_kw)
    state.record_tape = record_tape
    root = ws.model(src)(*build_args)
    state.sync()
    t0 = time.perf_counter()
    ws.run(root, state)
    le = ws.log_evidence(state)
    state.sync()
    return state, time.perf_counter() - t0, le


def emit(**kw):
    print(json.dumps(kw), flush=True)


def c1():
    rng = np.random.default_rng(7)
    x, v, obs = 0.0, 0.0, []
    for _ in range(50):
        obs.append(x + rng.standard_normal())
        x, v = x + v, v + 0.1 * rng.standard_normal()
    timed_run((obs,), models.SSM1D, 1000, 0.5)  # warm-up (library load, allocations)
    st, dt, le = timed_run((obs,), models.SSM1D, 1000, 0.5)
    emit(config="C1 examples/1D_ssm.jl N=1000 T=50 (history kept, 53 columns)", seconds=dt, particle_updates_per_sec=1000 * 50 / dt,
         log_evidence=le, resamples=st.stats()["resamples_done"], launches=st.stats()["kernel_launches"],
         note="launch-latency bound: 50 steps x (1 fused pass + reduce + resample kernels + host sync)")


def lgssm():
    from oracle import cref, ref
    a, q, r = 0.9, 1.0, 0.5
    rng = np.random.default_rng(42)
    T = 200 if QUICK else 1000
    x, ys = rng.standard_normal(), []
    for _ in range(T):
        x = a * x + q * rng.standard_normal()
        ys.append(x + r * rng.standard_normal())
    mean_exact, le_exact = ref.kalman_filter_evidence(ys, a, q, r)
    for n in (1_000, 1_000_000, 100_000_000 if not QUICK else 10_000_000):
        timed_run((ys[:5], a, q, r, 1.0), models.LGSSM1D, n, 1.0, record_tape=False)
        st, dt, le = timed_run((ys, a, q, r, 1.0), models.LGSSM1D, n, 1.0, record_tape=False)
        mean = ws.E(lambda x: x, st)
        emit(config=f"LGSSM-1D benchmarks/ssm/WeightedSampling/lgssm1d.jl T={T} N={n} ess_perc_min=1.0", seconds=dt,
             particle_updates_per_sec=n * T / dt, post_mean=mean, kalman_mean=mean_exact, log_evidence=le,
             kalman_log_evidence=le_exact, hbm_frac_80B=80.0 * n * T / dt / 1e9 / HBM,
             reference_published_seconds={1_000: 0.092459, 1_000_000: 22.170888}.get(n),
             note="reference numbers: Julia 1.12, 1 thread, author's unstated hardware (grid_results.csv:2-3,14-15), T=1000")
        del st
    n_cpu = 1_000_000
    t0 = time.perf_counter()
    le_c, mean_c, nres = cref.lgssm1d_run(n_cpu, ys, a, q, r, 1.0, seed=3, ess_perc_min=1.0)
    dt = time.perf_counter() - t0
    emit(config=f"LGSSM-1D CPU restatement (oracle/ws_oracle.c, 1 core) T={T} N={n_cpu}", seconds=dt,
         particle_updates_per_sec=n_cpu * T / dt, post_mean=mean_c, log_evidence=le_c)


def c2hist():
    """examples/2D_ssm.jl VERBATIM (x{t} history kept): the reference gathers all 2(t+3) planes at step t;
    the genealogy keeps one 4-byte ancestor vector per event instead."""
    rng = np.random.default_rng(42)
    T = 100 if QUICK else 400
    obs = [rng.standard_normal(2) * 0.5 + np.array([t, 0.0]) for t in range(T)]
    for n in ((1_000_000,) if QUICK else (1_000_000, 10_000_000)):
        for on in (True, False):
            if not on and n * T > 5e8:
                Tn = T // 4   # the eager variant is quadratic in T: shorter run, per-step figure quoted
            else:
                Tn = T
            st = ws.SMCState(n, ess_perc_min=1.0, seed=1)
            st.set_genealogy(on)
            root = ws.model(models.SSM2D)(obs[:Tn])
            st.sync()
            t0 = time.perf_counter()
            ws.run(root, st)
            le = ws.log_evidence(st)
            st.sync()
            dt = time.perf_counter() - t0
            t1 = time.perf_counter()
            x_mid = st[f"x_{Tn // 2}"]          # a column ~T/2 events behind
            dt_read = time.perf_counter() - t1
            emit(config=f"C2-history examples/2D_ssm.jl verbatim N={n} T={Tn} genealogy={'on' if on else 'off'}", seconds=dt,
                 particle_updates_per_sec=n * Tn / dt, ms_per_step=1e3 * dt / Tn, log_evidence=le,
                 genealogy=st.genealogy(), read_mid_column_seconds=dt_read, mid_column_mean=float(x_mid[:, 0].mean()),
                 columns=len(st.store.colnames()))
            del st


def hier():
    """benchmarks/multilevel (hierarchical regression, the accuracy-matched comparison protocol of
    run_benchmark.py:90-98): elapsed time, rmse of the posterior-mean alpha_j, particle ESS."""
    configs = ((20, 10),) if QUICK else ((20, 10), (100, 10), (100, 50))
    for J, n_obs in configs:
        groups, a_true = models.simulate_hier(J, n_obs)
        timed_run((1, [[(0.0, 0.0)]]), models.HIER, 50, 0.5)      # warm-up, as run_ws.jl:50-54
        for n in ((100_000,) if QUICK else (100_000, 1_000_000)):
            st, dt, le = timed_run((J, groups), models.HIER, n, 0.5)
            w = ws.exp_norm(st)
            a_est = np.array([float(np.sum(w * st[f"alpha_{j + 1}"])) for j in range(J)])
            s = st.stats()
            emit(config=f"multilevel hierarchical_regression J={J} n_obs={n_obs} N={n}", seconds=dt,
                 rmse_alpha=float(np.sqrt(np.mean((a_est - a_true) ** 2))), ess=float(1.0 / np.sum(w * w)),
                 mu_alpha=float(np.sum(w * st["mu_alpha"])), tau_alpha=float(np.sum(w * st["tau_alpha"])),
                 beta=float(np.sum(w * st["beta"])), sigma=float(np.sum(w * st["sigma"])), truth=[5.0, 2.0, 3.0, 1.0],
                 log_evidence=le, resamples=s["resamples_done"], moves=s["moves_run"], launches=s["kernel_launches"])
            del st


def c3():
    rng = np.random.default_rng(42)
    npts = 1000 if QUICK else 10_000
    n = 1_000_000 if QUICK else 10_000_000
    xs = rng.uniform(0, 10, npts)
    ys = 1.0 - 0.5 * xs + 0.5 * rng.standard_normal(npts)
    timed_run((xs[:20], ys[:20]), models.LINREG, 10_000, 0.5)
    st, dt, le = timed_run((xs, ys), models.LINREG, n, 0.5)
    s = st.stats()
    kt = st.kernel_times()
    a, b = ws.E(lambda α: α, st), ws.E(lambda β: β, st)
    emit(config=f"C3 examples/linear_regression.jl N={n}, {npts} points, autoRW moves after each resample", seconds=dt,
         particle_updates_per_sec=n * npts / dt, resample_events=s["resamples_done"], moves=s["moves_run"],
         alpha=a, beta=b, truth=[1.0, -0.5], log_evidence=le, launches=s["kernel_launches"])


def c4():
    n = 1_000_000 if QUICK else 10_000_000
    timed_run((8, models.SCHOOLS_Y, models.SCHOOLS_SIGMA), models.SCHOOLS, 10_000, 0.5)
    st, dt, le = timed_run((8, models.SCHOOLS_Y, models.SCHOOLS_SIGMA), models.SCHOOLS, n, 0.5)
    s = st.stats()
    mu = ws.E(lambda μ: μ, st)
    tau = ws.E(lambda τ: τ, st)
    emit(config=f"C4 examples/eight_schools.jl N={n} (one GPU), J=8, diversity-gated autoRW moves", seconds=dt,
         particle_updates_per_sec=n * 8 / dt, mu=mu, tau=tau, log_evidence=le, resamples=s["resamples_done"], moves=s["moves_run"])
    # synthetic J = 512 variant for stable timings
    J = 64 if QUICK else 512
    rng = np.random.default_rng(1)
    sig = rng.uniform(9, 18, J)
    th = 4.0 + 3.0 * rng.standard_normal(J)
    y = th + sig * rng.standard_normal(J)
    n2 = n // 10
    st, dt, le = timed_run((J, list(y), list(sig)), models.SCHOOLS, n2, 0.5)
    s = st.stats()
    emit(config=f"C4 synthetic J={J} N={n2}", seconds=dt, particle_updates_per_sec=n2 * J / dt, mu=ws.E(lambda μ: μ, st),
         log_evidence=le, resamples=s["resamples_done"], moves=s["moves_run"])


def c5():
    """resampling microbenchmark: logw = s*z, payload of P planes, one Resample (reduce + CDF + search + gather)."""
    import torch
    sizes = (1_000_000, 10_000_000) if QUICK else (1_000_000, 10_000_000, 100_000_000, 1_000_000_000)
    for n in sizes:
        for P in (1, 6, 16):
            if n * P * 16 > 120e9:
                continue
            for label, s in (("s=0.5", 0.5), ("s=2", 2.0), ("s=4", 4.0), ("one-hot", None)):
                for scheme in ("stratified", "systematic", "multinomial"):
                    if scheme != "stratified" and (label not in ("s=2", "one-hot") or P != 6):
                        continue
                    st = ws.SMCState(n, ess_perc_min=float("inf"), seed=0x5EED, resampler=scheme)
                    store = st.store
                    store._call("ws_set_lazy_gather", 0)
                    # payload planes and log-weights are generated ON the device by the library itself
                    for p in range(P):
                        ws.Sample(f"p{p}", "Normal", (0.0, 1.0)).apply(st)
                    ws.Sample("z", "Normal", (0.0, 1.0)).apply(st)
                    reps = 5
                    store._call("ws_set_timing", 1)
                    times = []
                    for rep in range(reps + 1):
                        if s is None:
                            lw = ws.col("z") * 0.0
                            ws.Weight(None, (lw,)).apply(st)
                            # one particle carries 99 % of the mass: add log(0.99 N / 0.01) to particle chosen by z max is
                            # awkward on device; use a sharp quadratic well instead: exp(-(z-5)^2 * 1e6)
                            ws.Weight(None, (-(ws.col("z") - 5.0) * (ws.col("z") - 5.0) * 1e3,)).apply(st)
                        else:
                            ws.Weight(None, (ws.col("z") * s,)).apply(st)
                        st.sync()
                        store._call("ws_reset_kernel_times")
                        r = ws.Resample()
                        r.apply(st)
                        st.sync()
                        kt = st.kernel_times()
                        ms = kt["finalize"]["ms"] + kt["scan_search"]["ms"] + kt["gather"]["ms"] + kt["reduce"]["ms"]
                        if rep > 0:
                            times.append(ms)
                        ess = r.last.ess_perc
                        # refresh z so that the next round is weighted afresh (z was resampled)
                        ws.Sample("z", "Normal", (0.0, 1.0)).apply(st)
                    ms = float(np.median(times))
                    planes = P + 1
                    alg = 32 + 16 * planes
                    emit(config=f"C5 resample N={n} payload={P}+1 planes {label} {scheme}", ms=ms, ess_perc=ess,
                         particles_per_sec=n / (ms * 1e-3), alg_bytes_per_particle=alg,
                         achieved_gbs=alg * n / (ms * 1e-3) / 1e9, hbm_frac=alg * n / (ms * 1e-3) / 1e9 / HBM)
                    r = None            # (the transformer keeps its store alive)
                    st.store.close()    # free the device memory before the next state is created (N = 1e9 fills the GPU)
                    del st


if __name__ == "__main__":
    which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["c1", "lgssm", "c2hist", "c3", "c4", "hier", "c5"]
    for w in which:
        {"c1": c1, "lgssm": lgssm, "c2hist": c2hist, "hier": hier, "c3": c3, "c4": c4, "c5": c5}[w]()
