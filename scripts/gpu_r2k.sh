#!/bin/bash
# round 2, GPU call 11 (1 GPU): chain form of CDF + search (look-back deferred by one tile): kernel-form tests, A/B against 3pass / 1pass
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_kernel_forms.py -m gpu -x -q > $OUT/pytest_r2k.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_r2k.log
tail -6 $OUT/pytest_r2k.log
for f in 3pass chain 1pass; do
  WSB200_SCAN=$f timeout 600 python bench.py --steps 30 --no-cpu-baseline > $OUT/bench_r2k_$f.log 2>&1; python scripts/brief.py $OUT/bench_r2k_$f.log $f
done
WSB200_SCAN=chain WSB200_LIB=$PWD/variants/chain4.so timeout 600 python bench.py --steps 30 --no-cpu-baseline > $OUT/bench_r2k_chain4.log 2>&1; python scripts/brief.py $OUT/bench_r2k_chain4.log chain_minb4
WSB200_SCAN=chain timeout 600 python bench.py --particles 20000000 --steps 30 --no-cpu-baseline > $OUT/bench_r2k_chain_2e7.log 2>&1; python scripts/brief.py $OUT/bench_r2k_chain_2e7.log chain_2e7
WSB200_SCAN=chain timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ws_chain_kernel' \
    --launch-skip 3 --launch-count 1 -o $OUT/prof_r2k_chain -f \
    python bench.py --particles 20000000 --steps 3 --warmup 3 --no-cpu-baseline --profile-steps 3 > $OUT/ncu_full_r2k.log 2>&1
ls -la $OUT/prof_r2k_chain.ncu-rep
