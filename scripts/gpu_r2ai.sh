#!/bin/bash
# round 2, GPU calls 35, 36 (2 and 4 GPUs): where a sharded step of a many-plane model spends its host time (examples/2D_ssm.jl verbatim, 806 planes)
OUT=gpurun_out; mkdir -p $OUT
WSB200_TRACE=1 WSB200_C2HIST_QUICK=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU:-2} --master-addr 127.0.0.1 --master-port 29673 benchmarks/run_sharded.py c2hist > $OUT/sharded_r2aj_c2hist_${NGPU:-2}gpu.jsonl 2> $OUT/sharded_r2aj_c2hist_${NGPU:-2}gpu.err
cut -c1-300 $OUT/sharded_r2aj_c2hist_${NGPU:-2}gpu.jsonl; grep wsb200 $OUT/sharded_r2aj_c2hist_${NGPU:-2}gpu.err | tail -8
