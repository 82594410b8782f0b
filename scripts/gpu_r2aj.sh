#!/bin/bash
# round 2, GPU call 37 (2 GPUs): sharded states grow their slabs geometrically (fewer IPC mapping rounds) — the genealogy tests and
# one parity case again, and the many-plane model's host phases
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -v -s -k "genealogy or 2-0.5-stratified-1 or direct_exchange" > $OUT/pytest_sharded_r2aj.log 2>&1; echo "sharded rc=$?" | tee -a $OUT/pytest_sharded_r2aj.log
grep "particles differ\|particles differing\|migrated\|passed\|failed\|skipped\|Error\|error\|assert" $OUT/pytest_sharded_r2aj.log | tail -12 | cut -c1-300
NGPU=2 bash scripts/gpu_r2ai.sh 2>&1 | cut -c1-420
