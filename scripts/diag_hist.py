"""Per-step time of the history-keeping 2D SSM as T grows (genealogy)."""
import sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, wsb200 as ws, models
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
rng = np.random.default_rng(42)
for T in (20, 80, 20, 80, 160, 320):
    obs = [rng.standard_normal(2) + np.array([t, 0.0]) for t in range(T)]
    st = ws.SMCState(n, ess_perc_min=1.0, seed=1, device=0)
    root = ws.model(models.SSM2D)(obs)
    st.store._call("ws_set_timing", 1)
    st.sync()
    t0 = time.perf_counter()
    ws.run(root, st)
    st.sync()
    dt = time.perf_counter() - t0
    kt = st.kernel_times()
    print(T, "ms/step", round(1e3 * dt / T, 4), {a: (round(b["ms"], 2), b["launches"]) for a, b in kt.items() if b["launches"]}, st.genealogy(), flush=True)
    del st
