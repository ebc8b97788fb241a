#!/bin/bash
# One gpurun session for the rounds scan: parity tests, sweep, tuning grid, ncu capture.  Output under gpurun_out/.
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_scan_rounds.py -x -q > $O/scan_tests.log 2>&1; echo "tests rc=$?" >> $O/scan_tests.log
tail -5 $O/scan_tests.log
timeout 300 python tools/scan_bench.py --mode rounds --json $O/scan_rounds_bf16.json > $O/scan_rounds_sweep.txt 2>&1
timeout 120 python tools/scan_bench.py --mode pipe --seqs 65536 --iters 10 >> $O/scan_rounds_sweep.txt 2>&1
for nst in 2 3; do for tc in 32 64 128 256; do
  echo "# tc=$tc nst=$nst" >> $O/scan_rounds_tune.txt
  timeout 120 python tools/scan_bench.py --mode rounds --seqs 16384,65536 --iters 10 --tc-fwd $tc --tc-bwd $tc --nst-fwd $nst --nst-bwd $nst >> $O/scan_rounds_tune.txt 2>&1
done; done
echo "# f32" >> $O/scan_rounds_sweep.txt
timeout 200 python tools/scan_bench.py --mode rounds --dtype f32 --seqs 16384,65536 --iters 10 >> $O/scan_rounds_sweep.txt 2>&1
echo "# block shape B8 H11" >> $O/scan_rounds_sweep.txt
timeout 200 python tools/scan_bench.py --mode rounds --heads 11 --batch 8 --seqs 4096 --iters 10 >> $O/scan_rounds_sweep.txt 2>&1
timeout 200 python tools/scan_bench.py --mode rounds --heads 25 --batch 8 --seqs 4096 --iters 10 >> $O/scan_rounds_sweep.txt 2>&1
cat $O/scan_rounds_sweep.txt $O/scan_rounds_tune.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_rounds_ -s 6 -c 4 -o $O/scan_rounds_64k python tools/scan_bench.py --mode rounds --seqs 65536 --iters 1 > $O/ncu.log 2>&1
tail -3 $O/ncu.log
