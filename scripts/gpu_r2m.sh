#!/bin/bash
# round 2, GPU call 13 (1 GPU): chain with preloaded look-back words; three-pass form with two-level tile offsets (no
# single-CTA scan of all tile words), trimmed fixed-point conversion; full suite
OUT=gpurun_out; mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -x -q > $OUT/pytest_r2m.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_r2m.log
tail -4 $OUT/pytest_r2m.log
for f in 3pass chain; do
  WSB200_SCAN=$f timeout 600 python bench.py --steps 30 --no-cpu-baseline > $OUT/bench_r2m_$f.log 2>&1; python scripts/brief.py $OUT/bench_r2m_$f.log $f
done
WSB200_SCAN=chain timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ws_chain_kernel' \
    --launch-skip 3 --launch-count 1 -o $OUT/prof_r2m_chain -f \
    python bench.py --particles 20000000 --steps 3 --warmup 3 --no-cpu-baseline --profile-steps 3 > $OUT/ncu_full_r2m.log 2>&1
WSB200_SCAN=3pass timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ws_cdf_tiles_kernel|ws_search_kernel|ws_cdf_group_offsets_kernel' \
    --launch-skip 9 --launch-count 3 -o $OUT/prof_r2m_3pass -f \
    python bench.py --particles 20000000 --steps 3 --warmup 3 --no-cpu-baseline --profile-steps 3 > $OUT/ncu_full_r2m_3pass.log 2>&1
ls -la $OUT/*.ncu-rep
