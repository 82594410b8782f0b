#!/bin/bash
# round 2, second GPU call: parity tests of the new kernel forms + oracle-tree replay tests, A/B benches, ncu of the two hot kernels
OUT=gpurun_out; mkdir -p $OUT; rm -f $OUT/parity_attribution.jsonl
timeout 1200 python -m pytest tests/test_gpu_kernel_forms.py tests/test_gpu_parity.py tests/test_gpu_moves.py tests/test_gpu_genealogy.py tests/test_gpu_expr_kernels.py -m gpu -x -q > $OUT/pytest_r2b.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_r2b.log
tail -8 $OUT/pytest_r2b.log
timeout 600 python bench.py --steps 30 --no-cpu-baseline > $OUT/bench_r2b.log 2>&1; python scripts/brief.py $OUT/bench_r2b.log default
WSB200_SCAN=3pass timeout 600 python bench.py --steps 30 --no-cpu-baseline > $OUT/bench_r2b_3pass.log 2>&1; python scripts/brief.py $OUT/bench_r2b_3pass.log 3pass
for v in slp4 fmb4 smb4; do
  WSB200_LIB=$PWD/variants/$v.so timeout 600 python bench.py --steps 30 --no-cpu-baseline > $OUT/bench_r2b_$v.log 2>&1; python scripts/brief.py $OUT/bench_r2b_$v.log $v
done
WSB200_LIB=$PWD/variants/smb4.so WSB200_SCAN=3pass timeout 600 python bench.py --steps 30 --no-cpu-baseline > $OUT/bench_r2b_smb4_3pass.log 2>&1; python scripts/brief.py $OUT/bench_r2b_smb4_3pass.log smb4_3pass
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ws_vm_sl_kernel|ws_scan_search_kernel' \
    --launch-skip 6 --launch-count 2 -o $OUT/prof_r2b -f \
    python bench.py --particles 20000000 --steps 3 --warmup 3 --no-cpu-baseline --profile-steps 3 > $OUT/ncu_full_r2b.log 2>&1
ls -la $OUT/prof_r2b.ncu-rep
