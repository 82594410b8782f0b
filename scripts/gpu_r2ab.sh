#!/bin/bash
# round 2, GPU call 28 (1 GPU): wide score tapes (C4 with J = 512) after the host-side fix
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests/test_gpu_moves.py tests/test_gpu_parity.py -m gpu -x -q > $OUT/pytest_r2ab.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_r2ab.log
tail -4 $OUT/pytest_r2ab.log | cut -c1-200
timeout 600 python benchmarks/run_configs.py c4 > $OUT/configs_r2ab.jsonl 2> $OUT/configs_r2ab.err; cut -c1-260 $OUT/configs_r2ab.jsonl; tail -2 $OUT/configs_r2ab.err
timeout 600 python scripts/diag_c4.py 512 1000000 2>&1 | tail -2 | cut -c1-400
