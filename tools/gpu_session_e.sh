#!/bin/bash
# GPU session: the two recalibrated tests, then this round's ncu evidence (launch list of the bench step, full-set capture of
# the expert GEMM launches and of the 64K scan).
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -k "c4_7b_dims_bf16 or fp16_autocast" > $O/two_tests.log 2>&1; echo "rc=$?" >> $O/two_tests.log
grep -E "^E  |passed|failed|^FAILED|rc=" $O/two_tests.log | cut -c1-1500 | head -20
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r2a_launches.csv python bench.py --steps 2 --warmup 3 --graph off --no-c4 --no-gpu-reference --no-cpu-baseline > $O/r2a_bench_under_ncu.log 2>&1; tail -2 $O/r2a_bench_under_ncu.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:grouped_gemm_kernel -s 27 -c 9 -o $O/r2a_gemm python tools/gemm_bench.py --iters 1 > $O/r2a_gemm_ncu.log 2>&1; tail -2 $O/r2a_gemm_ncu.log
ls -la $O/*.ncu-rep
