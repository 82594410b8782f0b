#!/bin/bash
# round 2, seventh GPU call (1 GPU): sweep of the straight-line kernel's shape, ported reference tests, ncu of the best shape
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_reference_ports.py tests/test_gpu_kernel_forms.py -m gpu -x -q > $OUT/pytest_r2g.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_r2g.log
tail -4 $OUT/pytest_r2g.log
timeout 600 python bench.py --steps 30 --no-cpu-baseline > $OUT/bench_r2g.log 2>&1; python scripts/brief.py $OUT/bench_r2g.log default_p2b6
for v in p3b4 p3b5 p4b4 p4b3 p2b7 p3b4imm; do
  WSB200_LIB=$PWD/variants/$v.so timeout 600 python bench.py --steps 30 --no-cpu-baseline > $OUT/bench_r2g_$v.log 2>&1; python scripts/brief.py $OUT/bench_r2g_$v.log $v
done
WSB200_LIB=$PWD/variants/p3b4.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ws_vm_sl_kernel' \
    --launch-skip 3 --launch-count 1 -o $OUT/prof_r2g_p3b4 -f \
    python bench.py --particles 20000000 --steps 3 --warmup 3 --no-cpu-baseline --profile-steps 3 > $OUT/ncu_full_r2g.log 2>&1
ls -la $OUT/prof_r2g_p3b4.ncu-rep
