#!/bin/bash
# One gpurun call: GPU tests, default bench, ncu launch list, ncu --set full of the hot kernels.
# usage: scripts/gpu_profile.sh <tag> [skip-tests]
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p $OUT
if [ -z "$2" ]; then
  python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_$TAG.log
  tail -5 $OUT/pytest_$TAG.log
fi
python bench.py > $OUT/bench_$TAG.log 2>&1 || { echo BENCH FAILED; tail -20 $OUT/bench_$TAG.log; exit 1; }
tail -c 2500 $OUT/bench_$TAG.log
python bench.py --impl reference --steps 3 > $OUT/bench_ref_$TAG.log 2>&1; tail -c 600 $OUT/bench_ref_$TAG.log
# launch list of the same command (short), then full capture of one launch of each hot kernel at 2e7
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'ws_vm_kernel|ws_search_kernel|ws_cdf_tiles_kernel|ws_cdf_offsets_kernel' \
    --launch-skip 8 --launch-count 4 -o $OUT/prof_$TAG -f \
    python bench.py --particles 20000000 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_$TAG.log 2>&1
ls -la $OUT/prof_$TAG.ncu-rep
