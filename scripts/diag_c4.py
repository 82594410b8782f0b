"""C4 synthetic (J schools, wide score tape): time per config; env WSB200_SEG_REGS tunes the segment size."""
import sys, time, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, wsb200 as ws, models
J = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
rng = np.random.default_rng(1)
sig = rng.uniform(9, 18, J); th = 4.0 + 3.0 * rng.standard_normal(J); y = th + sig * rng.standard_normal(J)
st = ws.SMCState(n, ess_perc_min=0.5, seed=3)
st.store._call("ws_set_timing", 1)
st.sync(); t0 = time.perf_counter()
ws.run(ws.model(models.SCHOOLS)(J, list(y), list(sig)), st)
mu = ws.E(lambda μ: μ, st); st.sync()
dt = time.perf_counter() - t0
kt = st.kernel_times(); s = st.stats()
print("SEG_REGS", os.environ.get("WSB200_SEG_REGS"), "J", J, "N", n, "seconds", round(dt, 3), "mu", round(mu, 4), "logZ", round(ws.log_evidence(st), 4),
      "moves", s["moves_run"], {a: (round(b["ms"], 1), b["launches"]) for a, b in kt.items() if b["launches"]}, flush=True)
