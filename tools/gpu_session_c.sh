#!/bin/bash
# GPU session: model-level tests, per-kernel device time of a step (torch profiler), ncu launch list of a step.
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_reference_model.py -q > $O/ref_model_tests.log 2>&1; echo "tests rc=$?" >> $O/ref_model_tests.log
tail -15 $O/ref_model_tests.log
timeout 300 python tools/step_kernels.py 8 60 > $O/step_kernels.txt 2>&1
cat $O/step_kernels.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_c2_b8.csv python bench.py --steps 2 --warmup 1 --graph off --no-cpu-baseline > $O/bench_ncu.log 2>&1
tail -2 $O/bench_ncu.log | cut -c1-300
