#!/bin/bash
# same-box A/B on the GPU box: unit tests, then the C3 step of the previous commit (_ab_old worktree, built beforehand) against the working tree,
# alternating; then the in-kernel GEMM trace of the working tree's debug build (if present).
cd "$(dirname "$0")/.."
ROUNDS=${1:-2}
B="--steps 20 --warmup 5 --skip-cpu --skip-eager --skip-recipe --skip-traffic"
timeout 900 python -m pytest tests/test_ops_gpu.py -q -x -k "gemm" 2>&1 | tail -4 | cut -c1-300
timeout 900 python -m pytest tests/test_step_parity_gpu.py tests/test_ref_parity_gpu.py -q -x 2>&1 | tail -3 | cut -c1-300
for r in $(seq $ROUNDS); do
  (cd _ab_old && timeout 300 python bench.py $B 2>/dev/null | python ../tools/bench_line.py old)
  timeout 300 python bench.py $B 2>/dev/null | python tools/bench_line.py new
done
if [ -f slam_llm_b200/libslam_b200_trace.so ]; then
  SLAM_B200_LIB=slam_llm_b200/libslam_b200_trace.so timeout 300 python tools/gemm_trace.py > gpurun_out/gemm_trace3.log 2> gpurun_out/gemm_trace3.err
  tail -1 gpurun_out/gemm_trace3.log | cut -c1-300
fi
