#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_reference_model.py -q -k "gemm or dense or linear or patched" > $O/h_tests.log 2>&1; echo "rc=$?" >> $O/h_tests.log
grep -E "^E  |passed|failed|^FAILED|rc=" $O/h_tests.log | cut -c1-600 | head -30
timeout 200 python tools/gemm_bench.py > $O/gemm_per_shape.txt 2>&1; cat $O/gemm_per_shape.txt
