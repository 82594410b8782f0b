#!/bin/bash
# A/B bench of variant builds without the test-suite, then ncu of one variant: scripts/gpu_variants2.sh tag prof_variant name1 name2 ...
TAG=$1; PV=$2; shift; shift
for v in "$@"; do
  WSB200_LIB=$PWD/variants/$v.so python bench.py --steps 30 --no-cpu-baseline > gpurun_out/bench_${TAG}_$v.log 2>&1; python scripts/brief.py gpurun_out/bench_${TAG}_$v.log $v
done
WSB200_LIB=$PWD/variants/$PV.so ncu --set full --clock-control none --import-source on -k regex:'ws_vm_kernel' \
    --launch-skip 4 --launch-count 1 -o gpurun_out/prof_${TAG}_$PV -f \
    python bench.py --particles 20000000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out/prof_${TAG}_$PV.ncu-rep
