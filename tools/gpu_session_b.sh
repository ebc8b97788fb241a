#!/bin/bash
# GPU session: full GPU test-suite, scan sweep (graph-timed), a short bench.
set -u
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/gpu_tests.log 2>&1; echo "tests rc=$?" >> $O/gpu_tests.log
tail -25 $O/gpu_tests.log
timeout 300 python tools/scan_bench.py --mode rounds --graph > $O/scan_rounds_sweep_graph.txt 2>&1
timeout 100 python tools/scan_bench.py --mode rounds --seqs 65536 --iters 10 >> $O/scan_rounds_sweep_graph.txt 2>&1
timeout 100 python tools/scan_bench.py --mode rounds --graph --heads 11 --batch 8 --seqs 4096 --iters 10 >> $O/scan_rounds_sweep_graph.txt 2>&1
cat $O/scan_rounds_sweep_graph.txt
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_b.json 2> $O/bench_b.err; tail -c 3000 $O/bench_b.json; tail -5 $O/bench_b.err
