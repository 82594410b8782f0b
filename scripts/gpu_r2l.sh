#!/bin/bash
# round 2, GPU call 12 (1 GPU): chain form with the next tile's log-weights prefetched (cp.async) and two barriers per tile,
# interior fast path of the search (A/B), ncu of the chain kernel
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_kernel_forms.py tests/test_gpu_parity.py -m gpu -x -q > $OUT/pytest_r2l.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_r2l.log
tail -4 $OUT/pytest_r2l.log
for f in 3pass chain; do
  WSB200_SCAN=$f timeout 600 python bench.py --steps 30 --no-cpu-baseline > $OUT/bench_r2l_$f.log 2>&1; python scripts/brief.py $OUT/bench_r2l_$f.log $f
done
for f in 3pass chain; do
  WSB200_SCAN=$f WSB200_LIB=$PWD/variants/nointerior.so timeout 600 python bench.py --steps 30 --no-cpu-baseline > $OUT/bench_r2l_${f}_nointerior.log 2>&1; python scripts/brief.py $OUT/bench_r2l_${f}_nointerior.log ${f}_nointerior
done
WSB200_SCAN=chain timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ws_chain_kernel' \
    --launch-skip 3 --launch-count 1 -o $OUT/prof_r2l_chain -f \
    python bench.py --particles 20000000 --steps 3 --warmup 3 --no-cpu-baseline --profile-steps 3 > $OUT/ncu_full_r2l.log 2>&1
ls -la $OUT/prof_r2l_chain.ncu-rep
