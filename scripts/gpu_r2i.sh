#!/bin/bash
# round 2, ninth GPU call (1 GPU): state check after the container was re-created — full suite, bench, reference arm, configs
OUT=gpurun_out; mkdir -p $OUT; rm -f $OUT/parity_attribution.jsonl
timeout 2400 python -m pytest tests -m gpu -x -q > $OUT/pytest_r2i.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_r2i.log
tail -5 $OUT/pytest_r2i.log
timeout 900 python bench.py > $OUT/bench_r2i.log 2>&1; python scripts/brief.py $OUT/bench_r2i.log default
timeout 600 python bench.py --impl reference --steps 3 > $OUT/bench_ref_r2i.log 2>&1; tail -c 400 $OUT/bench_ref_r2i.log
timeout 900 python benchmarks/run_configs.py c1 lgssm c3 > $OUT/configs_r2i.jsonl 2> $OUT/configs_r2i.err; cut -c1-260 $OUT/configs_r2i.jsonl
