#!/bin/bash
# Final single-GPU session: smoke(), the GPU test-suite, the bench line (default flags), the reference arm.
set -u
O=gpurun_out
mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -q -m gpu > $O/gpu_tests.log 2>&1; echo "tests rc=$?" >> $O/gpu_tests.log
grep -E "^E  |passed|failed|^FAILED|rc=" $O/gpu_tests.log | cut -c1-400 | head -20
timeout 900 python bench.py > $O/bench_final.json 2> $O/bench_final.err; python tools/bench_brief.py $O/bench_final.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; tail -c 600 $O/bench_ref.json
timeout 300 python tools/step_kernels.py 8 60 > $O/step_kernels.txt 2>&1
