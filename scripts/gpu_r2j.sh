#!/bin/bash
# round 2, tenth GPU call (4 GPUs): the sharded parity tests on the final exchange code (2- and 4-rank cases), and the
# 2- and 4-GPU bench lines with the sharded == single-GPU check and the migration-heavy step
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus_r2j.txt
timeout 1500 python -m pytest tests/test_gpu_sharded.py -m gpu -v -s > $OUT/pytest_sharded_r2j.log 2>&1; echo "sharded rc=$?" | tee -a $OUT/pytest_sharded_r2j.log
grep -c PASSED $OUT/pytest_sharded_r2j.log; grep "particles differ\|passed\|failed\|skipped" $OUT/pytest_sharded_r2j.log | tail -6
for n in 2 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 30 > $OUT/bench_${n}gpu_r2j.log 2>&1
  python - $n <<'PY'
import json, sys
n = sys.argv[1]
for l in open(f"gpurun_out/bench_{n}gpu_r2j.log"):
    if l.startswith("{"):
        d = json.loads(l); print(n, "gpu ms/step", round(d["ms_per_step"], 3), "value", d["value"], "parity", json.dumps(d["sharded_parity"])[:500], "migration", json.dumps(d.get("migration_step"))[:400])
        break
else:
    print(n, "NO RESULT"); print(open(f"gpurun_out/bench_{n}gpu_r2j.log").read()[-1500:])
PY
done
