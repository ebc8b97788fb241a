#!/bin/bash
# GPU session: kernel + module tests, bench (N=1), per-kernel device times of a step.
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_modules.py -q > $O/k_tests.log 2>&1; echo "rc=$?" >> $O/k_tests.log
grep -E "^E  |passed|failed|^FAILED|rc=" $O/k_tests.log | cut -c1-300 | head -12
timeout 900 python bench.py --steps 10 --warmup 3 --no-gpu-reference --no-cpu-baseline > $O/bench_k.json 2> $O/bench_k.err; python tools/bench_brief.py $O/bench_k.json | head -12
timeout 300 python tools/step_kernels.py 8 24 > $O/step_kernels.txt 2>&1; grep -v Warn $O/step_kernels.txt | head -27 | cut -c1-150
