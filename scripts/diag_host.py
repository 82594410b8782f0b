"""Host overhead per statement: C3 (10k observes, each followed by a Resample) at small N under cProfile."""
import sys, time, cProfile, pstats
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, wsb200 as ws, models
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
npts = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
rng = np.random.default_rng(42)
xs = rng.uniform(0, 10, npts); ys = 1.0 - 0.5 * xs + 0.5 * rng.standard_normal(npts)
for rep in range(2):
    st = ws.SMCState(n, ess_perc_min=0.5, seed=1)
    root = ws.model(models.LINREG)(xs, ys)
    st.sync(); t0 = time.perf_counter()
    if rep == 1:
        pr = cProfile.Profile(); pr.enable()
    ws.run(root, st)
    st.sync()
    if rep == 1:
        pr.disable()
    dt = time.perf_counter() - t0
    print("N", n, "steps", npts, "seconds", round(dt, 4), "us/step", round(1e6 * dt / npts, 1), st.stats()["resamples_done"], flush=True)
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
