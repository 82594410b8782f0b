#!/bin/bash
# round 2, GPU call 25 (4 GPUs): the sharded parity tests (2- and 4-rank cases) on the final code
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests/test_gpu_sharded.py -m gpu -v -s > $OUT/pytest_sharded_r2y.log 2>&1; echo "sharded rc=$?" | tee -a $OUT/pytest_sharded_r2y.log
grep -c PASSED $OUT/pytest_sharded_r2y.log; grep "particles differ\|passed\|failed\|skipped" $OUT/pytest_sharded_r2y.log | tail -8
