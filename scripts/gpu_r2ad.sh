#!/bin/bash
# round 2, GPU call 30 (2 GPUs): sharded steps with one host wait (decision read with the bounds) — sharded tests, A/B of the 2-GPU bench
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests/test_gpu_sharded.py -m gpu -v -s > $OUT/pytest_sharded_r2ad.log 2>&1; echo "sharded rc=$?" | tee -a $OUT/pytest_sharded_r2ad.log
grep "particles differ\|passed\|failed\|skipped\|Error" $OUT/pytest_sharded_r2ad.log | tail -12
for mw in 1 0; do
  WSB200_MERGED_WAIT=$mw timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2964$mw bench.py --gpus 2 --steps 40 --skew 0 > $OUT/bench_r2ad_2gpu_mw$mw.log 2>&1
  python - $mw <<'PY'
import json, sys
mw = sys.argv[1]
for l in open(f"gpurun_out/bench_r2ad_2gpu_mw{mw}.log"):
    if l.startswith("{"):
        d = json.loads(l); print("merged_wait", mw, "2gpu ms/step", round(d["ms_per_step"], 4), d["ms_per_step_chunks"], "parity mismatches", d["sharded_parity"]["mismatches"])
        break
else:
    print("NO RESULT", mw); print(open(f"gpurun_out/bench_r2ad_2gpu_mw{mw}.log").read()[-1200:])
PY
done
