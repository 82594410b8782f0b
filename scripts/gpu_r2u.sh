#!/bin/bash
# round 2, GPU call 21 (2 GPUs): full suite incl. the sharded tests, 1- and 2-GPU bench, sharded configs (C4, C5 incl.
# systematic / multinomial, rank-skewed migration), multinomial microbenchmark on one GPU
OUT=gpurun_out; mkdir -p $OUT; rm -f $OUT/parity_attribution.jsonl
timeout 2400 python -m pytest tests -m gpu -x -q > $OUT/pytest_r2u.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_r2u.log
tail -8 $OUT/pytest_r2u.log | cut -c1-200
timeout 600 python bench.py --steps 30 --no-cpu-baseline > $OUT/bench_r2u_1gpu.log 2>&1; python scripts/brief.py $OUT/bench_r2u_1gpu.log 1gpu
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 30 > $OUT/bench_r2u_2gpu.log 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/bench_r2u_2gpu.log"):
    if l.startswith("{"):
        d = json.loads(l); print("2gpu ms/step", round(d["ms_per_step"], 3), "value", f'{d["value"]:.4g}', "parity mismatches", d["sharded_parity"]["mismatches"], {k: round(v["avg_ms"], 3) for k, v in d["roofline"]["per_kernel"].items()}, "migration", json.dumps(d.get("migration"))[:300])
        break
else:
    print("2gpu NO RESULT"); print(open("gpurun_out/bench_r2u_2gpu.log").read()[-1500:])
PY
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622 benchmarks/run_sharded.py c4 c5 c5skew > $OUT/sharded_r2u_2gpu.jsonl 2> $OUT/sharded_r2u_2gpu.err
python - <<'PY'
import json
for l in open("gpurun_out/sharded_r2u_2gpu.jsonl"):
    if l.startswith("{"):
        d = json.loads(l); print(d["config"][:110], "|", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d.items() if k in ("ms", "seconds", "hbm_frac_per_gpu", "nvlink_frac_of_770", "nvlink_egress_gbs_max_rank")})
PY
tail -3 $OUT/sharded_r2u_2gpu.err
