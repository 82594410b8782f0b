#!/bin/bash
# round 2, third GPU call: async Resample + kernel-form tests, full GPU suite, A/B benches, small-N / C3 configs
OUT=gpurun_out; mkdir -p $OUT; rm -f $OUT/parity_attribution.jsonl
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_r2c.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_r2c.log
tail -8 $OUT/pytest_r2c.log
timeout 600 python bench.py --steps 30 --no-cpu-baseline > $OUT/bench_r2c.log 2>&1; python scripts/brief.py $OUT/bench_r2c.log default
WSB200_ASYNC_RESAMPLE=0 timeout 600 python bench.py --steps 30 --no-cpu-baseline > $OUT/bench_r2c_sync.log 2>&1; python scripts/brief.py $OUT/bench_r2c_sync.log sync_resample
for v in imm slp3 smb4; do
  WSB200_LIB=$PWD/variants/$v.so timeout 600 python bench.py --steps 30 --no-cpu-baseline > $OUT/bench_r2c_$v.log 2>&1; python scripts/brief.py $OUT/bench_r2c_$v.log $v
done
timeout 900 python benchmarks/run_configs.py c1 lgssm c3 > $OUT/configs_r2c.jsonl 2> $OUT/configs_r2c.err; cut -c1-330 $OUT/configs_r2c.jsonl
WSB200_ASYNC_RESAMPLE=0 timeout 600 python benchmarks/run_configs.py c1 lgssm --quick > $OUT/configs_r2c_sync.jsonl 2>&1; cut -c1-200 $OUT/configs_r2c_sync.jsonl
