#!/bin/bash
# same-box A/B on the GPU box: unit tests of the changed kernels, then the C3 step with a switch on / off, alternating.
#   usage: tools/ab_round.sh VAR "v1 v2 ..." [rounds]     e.g. tools/ab_round.sh SLAM_PAIR512 "long 0 all"
cd "$(dirname "$0")/.."
VAR=${1:-SLAM_PAIR512}; VALS=${2:-"long 0 all"}; ROUNDS=${3:-2}
B="--steps 20 --warmup 5 --skip-cpu --skip-eager --skip-recipe --skip-traffic"
timeout 900 python -m pytest tests/test_ops_gpu.py -q -x -k "gemm" 2>&1 | tail -4 | cut -c1-300
timeout 900 python -m pytest tests/test_step_parity_gpu.py tests/test_ref_parity_gpu.py -q -x 2>&1 | tail -3 | cut -c1-300
for r in $(seq $ROUNDS); do
  for v in $VALS; do
    env $VAR=$v timeout 300 python bench.py $B 2>/dev/null | python tools/bench_line.py "$VAR=$v"
  done
done
