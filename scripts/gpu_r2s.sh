#!/bin/bash
# round 2, GPU call 19 (1 GPU): speculative blocks, full suite, C3 + launch list of its kernels
OUT=gpurun_out; mkdir -p $OUT; rm -f $OUT/parity_attribution.jsonl
timeout 2400 python -m pytest tests -m gpu -x -q > $OUT/pytest_r2s.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_r2s.log
tail -25 $OUT/pytest_r2s.log | cut -c1-200
timeout 900 python benchmarks/run_configs.py c3 > $OUT/configs_r2s.jsonl 2> $OUT/configs_r2s.err; cut -c1-330 $OUT/configs_r2s.jsonl; tail -3 $OUT/configs_r2s.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_r2s_c3.csv \
    python benchmarks/run_configs.py c3 --quick > $OUT/ncu_launches_r2s.log 2>&1
python - $OUT/launches_r2s_c3.csv <<'PY'
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; kn = hdr.index("Kernel Name"); mv = hdr.index("Metric Value")
d = collections.defaultdict(list)
for r in rows[1:]:
    d[r[kn].split("(")[0][:60]].append(float(r[mv].replace(",", "")) / 1e3)
for k, v in d.items(): print(f"{k:62s} n={len(v):4d} median {sorted(v)[len(v)//2]:8.1f} us  total {sum(v)/1e3:8.2f} ms")
PY
