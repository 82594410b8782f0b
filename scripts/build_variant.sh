#!/bin/bash
# builds a variant of libwsb200.so with extra -D flags into gpurun_out/variants/<name>.so (scratch, for A/B runs)
# usage: scripts/build_variant.sh <name> "-DWS_VM_P=4 -DWS_VM_MINB=5"
set -e
NAME=$1; DEFS=$2
D=${WS_SRC:-/root/repo/weightedsampling.jl_b200/csrc}
O=/tmp/wsb200_variant_$NAME
mkdir -p $O /root/repo/variants
for f in ws_runtime ws_kernels ws_kernels_move ws_kernels_stats; do
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC $DEFS -Xptxas -v -c $D/$f.cu -o $O/$f.o 2> $O/$f.log &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o /root/repo/variants/$NAME.so $O/ws_runtime.o $O/ws_kernels.o $O/ws_kernels_move.o $O/ws_kernels_stats.o -lcudart -ldl
grep -A2 "ws_vm_kernelILb1" $O/ws_kernels.log | grep "Used\|spill"
