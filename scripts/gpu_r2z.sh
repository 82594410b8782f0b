#!/bin/bash
# round 2, GPU call 26 (1 GPU): loop elements batched behind the boundary (ws_exec_n) — tests, small-N configs
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_r2z.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_r2z.log
tail -4 $OUT/pytest_r2z.log
timeout 600 python benchmarks/run_configs.py c1 lgssm --quick > $OUT/configs_r2z.jsonl 2> $OUT/configs_r2z.err; cut -c1-200 $OUT/configs_r2z.jsonl
timeout 600 python - <<'PY'
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, wsb200 as ws, models
rng = np.random.default_rng(0)
ys = list(rng.standard_normal(1000))
for n in (1000, 10_000, 100_000, 1_000_000):
    best = 1e9
    for rep in range(4):
        st = ws.SMCState(n, ess_perc_min=1.0, seed=rep + 1)
        root = ws.model(models.LGSSM1D)(ys, 0.9, 1.0, 0.5, 1.0)
        st.sync(); t0 = time.perf_counter(); ws.run(root, st); le = ws.log_evidence(st); st.sync()
        best = min(best, time.perf_counter() - t0)
    print(f"LGSSM-1D T=1000 N={n}: {best*1e3:.2f} ms total, {best*1e3:.3f} us per step", flush=True)
PY
