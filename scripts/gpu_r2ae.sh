#!/bin/bash
# round 2, GPU call 31 (2 GPUs): mailbox exchanges (the step's small collectives done by the kernels over peer memory) —
# sharded tests (mailbox on and off), A/B of the 2-GPU bench
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -v -s -x > $OUT/pytest_sharded_r2ae.log 2>&1; echo "sharded rc=$?" | tee -a $OUT/pytest_sharded_r2ae.log
grep "particles differ\|passed\|failed\|skipped\|Error\|error" $OUT/pytest_sharded_r2ae.log | tail -16
for mb in 1 0; do
  WSB200_MAILBOX=$mb timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2965$mb bench.py --gpus 2 --steps 40 --skew 0 > $OUT/bench_r2ae_2gpu_mb$mb.log 2>&1
  python - $mb <<'PY'
import json, sys
mb = sys.argv[1]
for l in open(f"gpurun_out/bench_r2ae_2gpu_mb{mb}.log"):
    if l.startswith("{"):
        d = json.loads(l); print("mailbox", mb, "2gpu ms/step", round(d["ms_per_step"], 4), d.get("ms_per_step_chunks"), "parity", d["sharded_parity"])
        break
else:
    print("NO RESULT", mb); print(open(f"gpurun_out/bench_r2ae_2gpu_mb{mb}.log").read()[-1500:])
PY
done
