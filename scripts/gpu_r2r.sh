#!/bin/bash
# round 2, GPU call 18 (1 GPU): speculative blocks with the register-resident block kernel (C3), search trim
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_kernel_forms.py -m gpu -x -q > $OUT/pytest_r2r.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_r2r.log
tail -30 $OUT/pytest_r2r.log | cut -c1-200
timeout 900 python benchmarks/run_configs.py c3 > $OUT/configs_r2r.jsonl 2> $OUT/configs_r2r.err; cut -c1-400 $OUT/configs_r2r.jsonl; tail -3 $OUT/configs_r2r.err
WSB200_VM=interp timeout 900 python benchmarks/run_configs.py c3 2>/dev/null | cut -c1-200
timeout 600 python bench.py --steps 30 --no-cpu-baseline > $OUT/bench_r2r.log 2>&1; python scripts/brief.py $OUT/bench_r2r.log default
