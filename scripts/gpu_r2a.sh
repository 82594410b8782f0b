#!/bin/bash
# round 2, first GPU call: new-kernel parity tests, bench of the default build and of the A/B variants
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_kernel_forms.py tests/test_gpu_parity.py -m gpu -x -q > $OUT/pytest_r2a.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_r2a.log
tail -8 $OUT/pytest_r2a.log
timeout 600 python bench.py --steps 30 --no-cpu-baseline > $OUT/bench_r2a.log 2>&1; python scripts/brief.py $OUT/bench_r2a.log default
WSB200_VM=interp timeout 600 python bench.py --steps 30 --no-cpu-baseline > $OUT/bench_r2a_interp.log 2>&1; python scripts/brief.py $OUT/bench_r2a_interp.log interp
WSB200_SCAN=3pass timeout 600 python bench.py --steps 30 --no-cpu-baseline > $OUT/bench_r2a_3pass.log 2>&1; python scripts/brief.py $OUT/bench_r2a_3pass.log 3pass
for v in slp1 slp3 slp4 fmb2; do
  WSB200_LIB=$PWD/variants/$v.so timeout 600 python bench.py --steps 30 --no-cpu-baseline > $OUT/bench_r2a_$v.log 2>&1; python scripts/brief.py $OUT/bench_r2a_$v.log $v
done
