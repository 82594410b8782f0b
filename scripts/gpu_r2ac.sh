#!/bin/bash
# round 2, GPU call 29 (2 GPUs): sharded linear regression test + the rest of the sharded tests on the final code
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests/test_gpu_sharded.py -m gpu -v -s > $OUT/pytest_sharded_r2ac.log 2>&1; echo "sharded rc=$?" | tee -a $OUT/pytest_sharded_r2ac.log
grep "particles differ\|passed\|failed\|skipped\|Error" $OUT/pytest_sharded_r2ac.log | tail -12
