#!/bin/bash
# round 2, eighth GPU call (1 GPU): fatter-thread sweep of the straight-line kernel, single-kernel small-N resampler, full suite
OUT=gpurun_out; mkdir -p $OUT; rm -f $OUT/parity_attribution.jsonl
make -s -j8 -C weightedsampling.jl_b200/csrc > $OUT/build_r2h.log 2>&1 || { tail -20 $OUT/build_r2h.log; exit 1; }
for v in "p4b3:-DWS_SL_P=4 -DWS_SL_MINB=3" "p5b3:-DWS_SL_P=5 -DWS_SL_MINB=3" "p6b2:-DWS_SL_P=6 -DWS_SL_MINB=2" "p6b3:-DWS_SL_P=6 -DWS_SL_MINB=3" "p8b2:-DWS_SL_P=8 -DWS_SL_MINB=2" "p4b2:-DWS_SL_P=4 -DWS_SL_MINB=2"; do
  WS_SRC=$PWD/weightedsampling.jl_b200/csrc scripts/build_variant.sh ${v%%:*} "${v#*:}" > $OUT/build_r2h_${v%%:*}.log 2>&1 &
done
wait
for v in p4b3 p5b3 p6b2 p6b3 p8b2 p4b2; do echo -n "$v: "; grep -A2 "ws_vm_sl_kernelI10WsSigSsm2dLi" /tmp/wsb200_variant_$v/ws_kernels.log | grep -o "Used [0-9]* registers\|[0-9]* bytes spill stores" | tr '\n' ' '; echo; done
timeout 2400 python -m pytest tests -m gpu -x -q > $OUT/pytest_r2h.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_r2h.log
tail -5 $OUT/pytest_r2h.log
for v in p4b3 p5b3 p6b2 p6b3 p8b2 p4b2; do
  WSB200_LIB=$PWD/variants/$v.so timeout 600 python bench.py --steps 30 --no-cpu-baseline > $OUT/bench_r2h_$v.log 2>&1; python scripts/brief.py $OUT/bench_r2h_$v.log $v
done
timeout 600 python benchmarks/run_configs.py c1 lgssm --quick > $OUT/configs_r2h.jsonl 2> $OUT/configs_r2h.err; cut -c1-200 $OUT/configs_r2h.jsonl | head -3
