#!/bin/bash
# round 2, GPU call 27 (1 GPU): adaptive speculative blocks — tests (incl. every step firing), C3, hierarchical model
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests/test_gpu_kernel_forms.py tests/test_gpu_moves.py -m gpu -x -q > $OUT/pytest_r2aa.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_r2aa.log
tail -6 $OUT/pytest_r2aa.log | cut -c1-200
timeout 900 python benchmarks/run_configs.py c3 hier > $OUT/configs_r2aa.jsonl 2> $OUT/configs_r2aa.err; cut -c1-200 $OUT/configs_r2aa.jsonl; tail -2 $OUT/configs_r2aa.err
