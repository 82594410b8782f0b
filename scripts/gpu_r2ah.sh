#!/bin/bash
# round 2, GPU call 34 (1 GPU): consolidated run on the final code — full GPU suite, smoke, bench + reference arm, ncu launch
# list of the bench command
OUT=gpurun_out; mkdir -p $OUT; rm -f $OUT/parity_attribution.jsonl
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_r2ah.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_r2ah.log
tail -4 $OUT/pytest_r2ah.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > $OUT/bench_r2ah.log 2>&1; python scripts/brief.py $OUT/bench_r2ah.log default
timeout 600 python bench.py --impl reference --steps 3 > $OUT/bench_ref_r2ah.log 2>&1; tail -c 300 $OUT/bench_ref_r2ah.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_r2ah.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --profile-steps 2 > $OUT/ncu_launches_r2ah.log 2>&1
grep -c ws_ $OUT/launches_r2ah.csv
