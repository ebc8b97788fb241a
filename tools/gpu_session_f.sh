#!/bin/bash
# GPU session: GEMM kernel tests first (fail fast), then the whole suite, per-shape GEMM timing, bench, step profile.
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "gemm or dense or linear" > $O/gemm_tests.log 2>&1; rc=$?; echo "rc=$rc" >> $O/gemm_tests.log
grep -E "^E  |passed|failed|^FAILED|rc=" $O/gemm_tests.log | cut -c1-300 | head -20
if [ $rc -ne 0 ]; then tail -40 $O/gemm_tests.log | cut -c1-300; exit 0; fi
timeout 200 python tools/gemm_bench.py > $O/gemm_per_shape.txt 2>&1; cat $O/gemm_per_shape.txt
timeout 1500 python -m pytest tests -q -m gpu > $O/gpu_tests.log 2>&1; echo "tests rc=$?" >> $O/gpu_tests.log
grep -E "^E  |passed|failed|^FAILED|rc=" $O/gpu_tests.log | cut -c1-400 | head -30
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_f.json 2> $O/bench_f.err; python tools/bench_brief.py $O/bench_f.json; grep -v Warning $O/bench_f.err | tail -5
timeout 300 python tools/step_kernels.py 8 40 > $O/step_kernels.txt 2>&1; head -45 $O/step_kernels.txt | cut -c1-160
