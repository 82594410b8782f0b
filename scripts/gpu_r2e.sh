#!/bin/bash
# round 2, fifth GPU call (2 GPUs): full GPU suite incl. the sharded tests, 2-GPU bench with the sharded parity check
OUT=gpurun_out; mkdir -p $OUT; rm -f $OUT/parity_attribution.jsonl
timeout 2400 python -m pytest tests -m gpu -x -q > $OUT/pytest_r2e.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_r2e.log
tail -8 $OUT/pytest_r2e.log
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q -s > $OUT/pytest_sharded_r2e.log 2>&1; echo "sharded rc=$?" | tee -a $OUT/pytest_sharded_r2e.log
grep "particles differ\|passed\|failed" $OUT/pytest_sharded_r2e.log | tail -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 30 > $OUT/bench_2gpu_r2e.log 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/bench_2gpu_r2e.log"):
    if l.startswith("{"):
        d = json.loads(l); print("2gpu ms/step", round(d["ms_per_step"], 3), "parity", json.dumps(d["sharded_parity"])[:600])
PY
tail -3 $OUT/bench_2gpu_r2e.log | cut -c1-300
