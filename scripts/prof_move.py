"""ncu target: one autoRW move over a score tape of k Normal terms (C3's likelihood) at N particles."""
import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, wsb200 as ws, models
k, n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000, int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
rng = np.random.default_rng(42)
xs = rng.uniform(0, 10, k); ys = 1.0 - 0.5 * xs + 0.5 * rng.standard_normal(k)
st = ws.SMCState(n, ess_perc_min=0.0, seed=1)            # never resample: the moves below are what is measured
steps = [ws.Sample("α", "Normal", (0.0, 10.0)), ws.Sample("β", "Normal", (0.0, 10.0))]
steps += [ws.Observe(float(y), "Normal", (ws.col("α") + ws.col("β") * float(x), 1.0)) for x, y in zip(xs, ys)]
steps += [ws.Move(["α"], "autoRW"), ws.Move(["β"], "autoRW"), ws.Move(["α"], "autoRW")]
st.store._call("ws_set_timing", 1)
ws.run(ws.Sequence(*steps), st)
kt = st.kernel_times()
print({a: (round(b["ms"], 3), b["launches"]) for a, b in kt.items() if b["launches"]})
mv = kt["move"]
# per move: 2 moment passes + 1 move kernel; the move kernel folds k terms twice for n particles
print("terms/s ~", 3 * 2.0 * k * n / (mv["ms"] * 1e-3))
