#!/bin/bash
# round 2, GPU call 32 (2 GPUs): sharded genealogy (history columns stay in place over the events of a sharded run) + the
# whole sharded suite on the new event bookkeeping
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_sharded.py -m gpu -v -s > $OUT/pytest_sharded_r2af.log 2>&1; echo "sharded rc=$?" | tee -a $OUT/pytest_sharded_r2af.log
grep "particles differ\|particles differing\|passed\|failed\|skipped\|Error\|error\|assert" $OUT/pytest_sharded_r2af.log | tail -24
timeout 300 python -m pytest tests/test_gpu_genealogy.py -m gpu -q 2>&1 | tail -3
