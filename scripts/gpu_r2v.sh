#!/bin/bash
# round 2, GPU call 22 (8 GPUs): 8-GPU bench line (sharded parity + migration-heavy step), sharded configs over 8 GPUs
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 8 --steps 30 > $OUT/bench_r2v_8gpu.log 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/bench_r2v_8gpu.log"):
    if l.startswith("{"):
        d = json.loads(l); print("8gpu ms/step", round(d["ms_per_step"], 3), "value", f'{d["value"]:.4g}', "parity mismatches", d["sharded_parity"]["mismatches"], {k: round(v["avg_ms"], 3) for k, v in d["roofline"]["per_kernel"].items()}, "migration", json.dumps(d.get("migration"))[:400])
        break
else:
    print("8gpu NO RESULT"); print(open("gpurun_out/bench_r2v_8gpu.log").read()[-1500:])
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus 4 --steps 30 --skew 0 > $OUT/bench_r2v_4gpu.log 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/bench_r2v_4gpu.log"):
    if l.startswith("{"):
        d = json.loads(l); print("4gpu ms/step", round(d["ms_per_step"], 3), "value", f'{d["value"]:.4g}', "parity mismatches", d["sharded_parity"]["mismatches"])
        break
PY
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29633 benchmarks/run_sharded.py c4 c5 c5skew > $OUT/sharded_r2v_8gpu.jsonl 2> $OUT/sharded_r2v_8gpu.err
python - <<'PY'
import json
for l in open("gpurun_out/sharded_r2v_8gpu.jsonl"):
    if l.startswith("{"):
        d = json.loads(l); print(d["config"][:110], "|", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d.items() if k in ("ms", "seconds", "hbm_frac_per_gpu", "nvlink_frac_of_770", "nvlink_egress_gbs_max_rank")})
PY
tail -3 $OUT/sharded_r2v_8gpu.err | cut -c1-300
