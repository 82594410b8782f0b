#!/bin/bash
# round 2, GPU call 17 (1 GPU): speculative blocks (C3)
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_kernel_forms.py -m gpu -x -q -k "speculative" > $OUT/pytest_r2q.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_r2q.log
tail -30 $OUT/pytest_r2q.log | cut -c1-200
timeout 900 python benchmarks/run_configs.py c3 > $OUT/configs_r2q.jsonl 2> $OUT/configs_r2q.err; cut -c1-400 $OUT/configs_r2q.jsonl; tail -3 $OUT/configs_r2q.err
WSB200_SPEC_BLOCKS=0 timeout 900 python benchmarks/run_configs.py c3 > $OUT/configs_r2q_nospec.jsonl 2> $OUT/configs_r2q_nospec.err; cut -c1-400 $OUT/configs_r2q_nospec.jsonl
for k in 8 32; do WSB200_SPEC_BLOCK_STEPS=$k timeout 900 python benchmarks/run_configs.py c3 2>/dev/null | cut -c1-200; done
