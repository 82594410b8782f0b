#!/bin/bash
# round 2, seventh GPU call (1 GPU): sweep of the straight-line kernel's shape (variants built on the box from this
# snapshot), ported reference tests, loop templates, small-N configs, ncu of the best shape
OUT=gpurun_out; mkdir -p $OUT
make -s -j8 -C weightedsampling.jl_b200/csrc > $OUT/build_r2g.log 2>&1 || { tail -20 $OUT/build_r2g.log; exit 1; }
for v in "p3b4:-DWS_SL_P=3 -DWS_SL_MINB=4" "p3b5:-DWS_SL_P=3 -DWS_SL_MINB=5" "p4b4:-DWS_SL_P=4 -DWS_SL_MINB=4" "p4b3:-DWS_SL_P=4 -DWS_SL_MINB=3" "p2b7:-DWS_SL_P=2 -DWS_SL_MINB=7"; do
  WS_SRC=$PWD/weightedsampling.jl_b200/csrc scripts/build_variant.sh ${v%%:*} "${v#*:}" > /dev/null 2>&1 &
done
wait
ls variants/
timeout 1200 python -m pytest tests/test_reference_ports.py tests/test_gpu_kernel_forms.py tests/test_abi.py -m gpu -x -q > $OUT/pytest_r2g.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_r2g.log
tail -4 $OUT/pytest_r2g.log
timeout 600 python bench.py --steps 30 --no-cpu-baseline > $OUT/bench_r2g.log 2>&1; python scripts/brief.py $OUT/bench_r2g.log default_p2b6
for v in p3b4 p3b5 p4b4 p4b3 p2b7; do
  WSB200_LIB=$PWD/variants/$v.so timeout 600 python bench.py --steps 30 --no-cpu-baseline > $OUT/bench_r2g_$v.log 2>&1; python scripts/brief.py $OUT/bench_r2g_$v.log $v
done
timeout 600 python benchmarks/run_configs.py c1 lgssm --quick > $OUT/configs_r2g.jsonl 2> $OUT/configs_r2g.err; cut -c1-220 $OUT/configs_r2g.jsonl
WSB200_LOOP_TEMPLATE=0 timeout 600 python benchmarks/run_configs.py lgssm --quick > $OUT/configs_r2g_notmpl.jsonl 2>&1; cut -c1-220 $OUT/configs_r2g_notmpl.jsonl | head -3
WSB200_LIB=$PWD/variants/p3b4.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ws_vm_sl_kernel' \
    --launch-skip 3 --launch-count 1 -o $OUT/prof_r2g_p3b4 -f \
    python bench.py --particles 20000000 --steps 3 --warmup 3 --no-cpu-baseline --profile-steps 3 > $OUT/ncu_full_r2g.log 2>&1
ls -la $OUT/prof_r2g_p3b4.ncu-rep
