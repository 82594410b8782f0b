#!/bin/bash
# round 2, GPU call 14 (1 GPU): three-pass form with prefetching search / tile kernels; sweep of the straight-line
# fused pass's shape (particles per thread x resident CTAs), CDF occupancy, search grid
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_kernel_forms.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q > $OUT/pytest_r2n.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_r2n.log
tail -3 $OUT/pytest_r2n.log
for f in 3pass chain; do
  WSB200_SCAN=$f timeout 600 python bench.py --steps 30 --no-cpu-baseline > $OUT/bench_r2n_$f.log 2>&1; python scripts/brief.py $OUT/bench_r2n_$f.log $f
done
for v in p3b4 p4b3 p3b5 p4b4 cdf6 srch6; do
  WSB200_LIB=$PWD/variants/$v.so timeout 600 python bench.py --steps 30 --no-cpu-baseline > $OUT/bench_r2n_$v.log 2>&1; python scripts/brief.py $OUT/bench_r2n_$v.log $v
done
