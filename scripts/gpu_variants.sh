#!/bin/bash
# A/B bench of variant builds: scripts/gpu_variants.sh tag name1 name2 ...
TAG=$1; shift
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
python bench.py --steps 30 --no-cpu-baseline > gpurun_out/bench_$TAG.log 2>&1; python scripts/brief.py gpurun_out/bench_$TAG.log main
for v in "$@"; do
  WSB200_LIB=$PWD/variants/$v.so python bench.py --steps 30 --no-cpu-baseline > gpurun_out/bench_${TAG}_$v.log 2>&1; python scripts/brief.py gpurun_out/bench_${TAG}_$v.log $v
done
