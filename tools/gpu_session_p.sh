#!/bin/bash
# GPU session: this round's evidence - full-set ncu captures of the expert GEMM launches and of the 64K scan, the launch
# list of an eager bench step, the GPU test-suite and the bench line.
set -u
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu > $O/gpu_tests.log 2>&1; echo "tests rc=$?" >> $O/gpu_tests.log
grep -E "^E  |passed|failed|^FAILED|rc=" $O/gpu_tests.log | cut -c1-400 | head -20
timeout 600 ncu --set full --clock-control none --import-source on -k regex:grouped_gemm_kernel -s 27 -c 9 -o $O/r2e_gemm python tools/gemm_bench.py --iters 1 > $O/r2e_gemm_ncu.log 2>&1; tail -1 $O/r2e_gemm_ncu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_rounds_ -s 6 -c 3 -o $O/r2e_scan python tools/scan_bench.py --mode rounds --seqs 65536 --iters 1 > $O/r2e_scan_ncu.log 2>&1; tail -1 $O/r2e_scan_ncu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/r2e_launches.csv python bench.py --steps 2 --warmup 3 --graph off --no-c4 --no-gpu-reference --no-cpu-baseline > $O/r2e_bench_under_ncu.log 2>&1; tail -c 300 $O/r2e_bench_under_ncu.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_p.json 2> $O/bench_p.err; python tools/bench_brief.py $O/bench_p.json
timeout 200 python tools/scan_bench.py --mode rounds --graph > $O/r2e_scan_sweep.txt 2>&1; cat $O/r2e_scan_sweep.txt
