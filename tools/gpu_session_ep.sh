#!/bin/bash
# Multi-GPU session (N = $1): NCCL + peer-memory parity tests of the expert-parallel layer, then the bench at N GPUs with
# each transport.
set -u
N=${1:-2}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_ep.py -q -m gpu > $O/ep${N}_test.log 2>&1; echo "rc=$?" >> $O/ep${N}_test.log; grep -E "^E  |passed|failed|rc=|gpu ep ok" $O/ep${N}_test.log | cut -c1-400 | tail -12
for T in ${2:-peer nccl}; do
  APERTIS_B200_EP=$T timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_ep${N}_$T.json 2> $O/bench_ep${N}_$T.err
  echo "== transport $T"; python tools/bench_brief.py $O/bench_ep${N}_$T.json; grep -v -i "warn\|OMP_NUM\|^\*\*\*\|run_backward" $O/bench_ep${N}_$T.err | tail -6 | cut -c1-400
done
