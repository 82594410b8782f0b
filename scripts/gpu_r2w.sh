#!/bin/bash
# round 2, GPU call 23 (1 GPU): compute-sanitizer (memcheck, racecheck) over the kernels written this round — chain form,
# prefetching search, checkpointed windows — at small sizes; multinomial microbenchmark after the window fix
OUT=gpurun_out; mkdir -p $OUT
cat > /tmp/san_case.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import wsb200 as ws
from models import LINREG, SSM2D_FILTER
form = sys.argv[1]
os.environ["WSB200_SCAN"] = form
os.environ["WSB200_SMALL_RESAMPLE"] = "0"
rng = np.random.default_rng(1)
# filter (mode 0 scan forms, gated async resample, deferred gather)
st = ws.SMCState(70_001, ess_perc_min=0.7, seed=3, device=0)
ws.run(ws.model(SSM2D_FILTER)([rng.standard_normal(2) + np.array([t, 0.0]) for t in range(6)]), st)
print(form, "filter", ws.log_evidence(st), st.stats()["resamples_done"])
# resample_indices: Philox grid, replayed grid, multinomial, one-hot
w = np.exp(2.0 * rng.standard_normal(50_003)); w /= w.sum()
for scheme in ("stratified", "systematic", "multinomial"):
    a = ws.resample_indices(w, st, scheme); print(form, scheme, int(a.sum() % 1000003))
a = ws.resample_indices(w, st, "stratified", uniforms=rng.random(50_003)); print(form, "replayed", int(a.sum() % 1000003))
w1 = np.zeros(50_003); w1[777] = 1.0
a = ws.resample_indices(w1, st, "stratified"); assert (a == 777).all()
# speculative blocks (both checkpoint kernels)
xs = rng.uniform(0, 10, 60); ys = 1 - 0.5 * xs + rng.standard_normal(60)
for vm in ("", "interp"):
    if vm: os.environ["WSB200_VM"] = vm
    s2 = ws.SMCState(30_011, ess_perc_min=0.5, seed=5, device=0)
    ws.run(ws.model(LINREG)(list(xs), list(ys)), s2)
    print(form, "linreg", vm or "sl", ws.log_evidence(s2), s2.stats()["resamples_done"], s2.stats()["moves_run"])
os.environ.pop("WSB200_VM", None)
PY
for tool in memcheck racecheck; do
  for form in 3pass chain; do
    timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python /tmp/san_case.py $form > $OUT/san_${tool}_$form.log 2>&1; echo "$tool $form rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" $OUT/san_${tool}_$form.log | tail -3
  done
done
timeout 600 python -m pytest tests/test_gpu_kernel_forms.py -m gpu -x -q -k speculative > $OUT/pytest_r2w.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_r2w.log
timeout 900 python - <<'PY' 2>&1 | tail -12
import sys, json, subprocess
out = subprocess.run([sys.executable, "benchmarks/run_configs.py", "c5", "--quick"], capture_output=True, text=True).stdout
for l in out.splitlines():
    if l.startswith("{") and ("multinomial" in l or "6+1 planes s=2 stratified" in l):
        d = json.loads(l); print(d["config"][12:], "ms", round(d["ms"], 3), "hbm", round(d["hbm_frac"], 3))
PY
