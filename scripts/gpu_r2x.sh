#!/bin/bash
# round 2, GPU call 24 (1 GPU): final consolidated run — full suite, bench + reference arm, ncu launch list of the bench
# command, ncu --set full of every hot kernel at 2e7, all single-GPU configs
OUT=gpurun_out; mkdir -p $OUT; rm -f $OUT/parity_attribution.jsonl
timeout 2400 python -m pytest tests -m gpu -x -q > $OUT/pytest_r2x.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_r2x.log
tail -4 $OUT/pytest_r2x.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > $OUT/bench_r2x.log 2>&1; python scripts/brief.py $OUT/bench_r2x.log default
timeout 600 python bench.py --impl reference --steps 3 > $OUT/bench_ref_r2x.log 2>&1; tail -c 300 $OUT/bench_ref_r2x.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_r2x.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --profile-steps 2 > $OUT/ncu_launches_r2x.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ws_vm_sl_kernel|ws_cdf_tiles_kernel|ws_search_kernel|ws_cdf_group_offsets_kernel' \
    --launch-skip 12 --launch-count 4 -o $OUT/prof_r2x -f \
    python bench.py --particles 20000000 --steps 3 --warmup 3 --no-cpu-baseline --profile-steps 3 > $OUT/ncu_full_r2x.log 2>&1
ls -la $OUT/prof_r2x.ncu-rep
timeout 1500 python benchmarks/run_configs.py c1 lgssm c3 c2hist hier c4 > $OUT/configs_r2x.jsonl 2> $OUT/configs_r2x.err; cut -c1-200 $OUT/configs_r2x.jsonl
timeout 1500 python benchmarks/run_configs.py c5 > $OUT/configs_r2x_c5.jsonl 2> $OUT/configs_r2x_c5.err; grep -c config $OUT/configs_r2x_c5.jsonl; tail -2 $OUT/configs_r2x_c5.err
