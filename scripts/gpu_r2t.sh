#!/bin/bash
# round 2, GPU call 20 (1 GPU): full suite, C5 resampling microbenchmark grid (stratified / systematic / multinomial)
OUT=gpurun_out; mkdir -p $OUT; rm -f $OUT/parity_attribution.jsonl
timeout 2400 python -m pytest tests -m gpu -x -q > $OUT/pytest_r2t.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_r2t.log
tail -12 $OUT/pytest_r2t.log | cut -c1-200
timeout 1500 python benchmarks/run_configs.py c5 > $OUT/configs_r2t_c5.jsonl 2> $OUT/configs_r2t_c5.err; python - <<'PY'
import json
for l in open("gpurun_out/configs_r2t_c5.jsonl"):
    d = json.loads(l); print(d["config"][12:], "ms", round(d["ms"], 3), "hbm", round(d["hbm_frac"], 3), "ess", f'{d["ess_perc"]:.2e}')
PY
tail -3 $OUT/configs_r2t_c5.err
timeout 600 python benchmarks/run_configs.py c4 > $OUT/configs_r2t_c4.jsonl 2> $OUT/configs_r2t_c4.err; cut -c1-300 $OUT/configs_r2t_c4.jsonl
