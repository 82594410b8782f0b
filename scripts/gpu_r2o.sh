#!/bin/bash
# round 2, GPU call 15 (1 GPU): per-kernel A/B of the prefetch variants (ncu launch durations), chain tile sizes
# (256 / 64 / 32 threads per CTA), fatter straight-line shapes
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_kernel_forms.py -m gpu -x -q > $OUT/pytest_r2o.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_r2o.log
tail -3 $OUT/pytest_r2o.log
WSB200_SCAN=chain WSB200_CHAIN_BLOCK=32 timeout 600 python -m pytest tests/test_gpu_kernel_forms.py -m gpu -x -q -k "scan or single_pass or slot_grid" > $OUT/pytest_r2o_b32.log 2>&1; echo "pytest b32 rc=$?"; tail -2 $OUT/pytest_r2o_b32.log
WSB200_SCAN=chain WSB200_CHAIN_BLOCK=64 timeout 600 python -m pytest tests/test_gpu_kernel_forms.py -m gpu -x -q -k "scan or single_pass or slot_grid" > $OUT/pytest_r2o_b64.log 2>&1; echo "pytest b64 rc=$?"; tail -2 $OUT/pytest_r2o_b64.log
timeout 600 python bench.py --steps 30 --no-cpu-baseline > $OUT/bench_r2o_base.log 2>&1; python scripts/brief.py $OUT/bench_r2o_base.log base_3pass
for b in 256 64 32; do
  WSB200_SCAN=chain WSB200_CHAIN_BLOCK=$b timeout 600 python bench.py --steps 30 --no-cpu-baseline > $OUT/bench_r2o_chain$b.log 2>&1; python scripts/brief.py $OUT/bench_r2o_chain$b.log chain$b
done
for v in cdfreg srchold bothold p5b3 p6b2 p5b2 p6b3 p8b2; do
  WSB200_LIB=$PWD/variants/$v.so timeout 600 python bench.py --steps 30 --no-cpu-baseline > $OUT/bench_r2o_$v.log 2>&1; python scripts/brief.py $OUT/bench_r2o_$v.log $v
done
for v in base bothold; do
  L=""; [ $v != base ] && L=$PWD/variants/$v.so
  WSB200_LIB=$L timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_r2o_$v.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline --profile-steps 2 > $OUT/ncu_launches_r2o_$v.log 2>&1
  python - $OUT/launches_r2o_$v.csv $v <<'PY'
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; kn = hdr.index("Kernel Name"); mv = hdr.index("Metric Value")
d = collections.defaultdict(list)
for r in rows[1:]:
    d[r[kn].split("(")[0][:40]].append(float(r[mv].replace(",", "")) / 1e3)
print(sys.argv[2], {k: round(sorted(v)[len(v) // 2], 1) for k, v in d.items() if sorted(v)[len(v) // 2] > 3})
PY
done
