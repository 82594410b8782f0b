#!/bin/bash
# round 2, fourth GPU call (1 GPU): full GPU suite, default bench + reference arm, configs, launch list, ncu of the hot kernels
OUT=gpurun_out; mkdir -p $OUT; rm -f $OUT/parity_attribution.jsonl
timeout 1800 python -m pytest tests -m gpu -x -q > $OUT/pytest_r2d.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_r2d.log
tail -6 $OUT/pytest_r2d.log
timeout 900 python bench.py > $OUT/bench_r2d.log 2>&1; python scripts/brief.py $OUT/bench_r2d.log default; tail -c 600 $OUT/bench_r2d.log
timeout 600 python bench.py --impl reference --steps 3 > $OUT/bench_ref_r2d.log 2>&1; tail -c 300 $OUT/bench_ref_r2d.log
timeout 900 python benchmarks/run_configs.py c1 lgssm c3 c4 > $OUT/configs_r2d.jsonl 2> $OUT/configs_r2d.err; cut -c1-250 $OUT/configs_r2d.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_r2d.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --profile-steps 3 > $OUT/ncu_launches_r2d.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ws_vm_sl_kernel|ws_search_kernel|ws_cdf_tiles_kernel|ws_cdf_offsets_kernel' \
    --launch-skip 8 --launch-count 4 -o $OUT/prof_r2d -f \
    python bench.py --particles 20000000 --steps 3 --warmup 3 --no-cpu-baseline --profile-steps 3 > $OUT/ncu_full_r2d.log 2>&1
ls -la $OUT/prof_r2d.ncu-rep
