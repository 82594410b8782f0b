#!/bin/bash
# round 2, GPU call 33 (4 GPUs): mailbox exchanges and the sharded genealogy on 4 ranks (the 4-rank cases of the sharded suite),
# 4-GPU bench, examples/2D_ssm.jl verbatim as one sharded filter with and without the genealogy
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -v -s -k "4-" > $OUT/pytest_sharded_r2ag.log 2>&1; echo "sharded rc=$?" | tee -a $OUT/pytest_sharded_r2ag.log
grep "particles differ\|particles differing\|passed\|failed\|skipped\|Error\|error\|assert" $OUT/pytest_sharded_r2ag.log | tail -12
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29671 bench.py --gpus 4 --steps 40 > $OUT/bench_r2ag_4gpu.log 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/bench_r2ag_4gpu.log"):
    if l.startswith("{"):
        d = json.loads(l); print("4gpu ms/step", round(d["ms_per_step"], 4), d.get("ms_per_step_chunks"), "value", d["value"], "parity mismatches", d["sharded_parity"]["mismatches"])
        break
else:
    print("NO RESULT"); print(open("gpurun_out/bench_r2ag_4gpu.log").read()[-1500:])
PY
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29672 benchmarks/run_sharded.py c2hist > $OUT/sharded_r2ag_c2hist_4gpu.jsonl 2> $OUT/sharded_r2ag_c2hist_4gpu.err
cut -c1-330 $OUT/sharded_r2ag_c2hist_4gpu.jsonl; tail -3 $OUT/sharded_r2ag_c2hist_4gpu.err
