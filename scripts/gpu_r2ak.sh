#!/bin/bash
# round 2, GPU call 38 (2 GPUs): the 2-GPU bench line on the final code, driver flags
OUT=gpurun_out; mkdir -p $OUT
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29681 bench.py --gpus 2 --steps 20 --warmup 3 > $OUT/bench_r2ak_2gpu.log 2>&1
grep '^{' $OUT/bench_r2ak_2gpu.log | cut -c1-400 || tail -20 $OUT/bench_r2ak_2gpu.log
