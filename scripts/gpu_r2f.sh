#!/bin/bash
# round 2, sixth GPU call (2 GPUs): GPU suite, sharded tests incl. multinomial, 1-GPU A/B of constant pinning, 2-GPU bench
OUT=gpurun_out; mkdir -p $OUT; rm -f $OUT/parity_attribution.jsonl
timeout 2400 python -m pytest tests -m gpu -x -q > $OUT/pytest_r2f.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_r2f.log
tail -6 $OUT/pytest_r2f.log
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q -s > $OUT/pytest_sharded_r2f.log 2>&1; echo "sharded rc=$?" | tee -a $OUT/pytest_sharded_r2f.log
grep "particles differ\|passed\|failed" $OUT/pytest_sharded_r2f.log | tail -12
timeout 600 python bench.py --steps 30 --no-cpu-baseline > $OUT/bench_r2f.log 2>&1; python scripts/brief.py $OUT/bench_r2f.log pinned
for v in nopin pinp3; do
  WSB200_LIB=$PWD/variants/$v.so timeout 600 python bench.py --steps 30 --no-cpu-baseline > $OUT/bench_r2f_$v.log 2>&1; python scripts/brief.py $OUT/bench_r2f_$v.log $v
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 30 > $OUT/bench_2gpu_r2f.log 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/bench_2gpu_r2f.log"):
    if l.startswith("{"):
        d = json.loads(l); print("2gpu ms/step", round(d["ms_per_step"], 3), "parity mismatches", d["sharded_parity"]["mismatches"], {k: round(v["avg_ms"], 3) for k, v in d["roofline"]["per_kernel"].items()})
PY
