#!/bin/bash
# GPU session: GPU test-suite, GEMM per-shape bench, block bench (N=1), per-kernel device times of a step.
set -u
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu > $O/gpu_tests.log 2>&1; echo "tests rc=$?" >> $O/gpu_tests.log
grep -E "^E  |passed|failed|^FAILED" $O/gpu_tests.log | cut -c1-250 | head -30
timeout 200 python tools/gemm_bench.py > $O/gemm_per_shape.txt 2>&1; cat $O/gemm_per_shape.txt
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_d.json 2> $O/bench_d.err; tail -c 6000 $O/bench_d.json; grep -v Warning $O/bench_d.err | tail -5
timeout 300 python tools/step_kernels.py 8 70 > $O/step_kernels.txt 2>&1; head -60 $O/step_kernels.txt
