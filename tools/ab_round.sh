#!/bin/bash
# same-box A/B on the GPU box: unit tests of the changed kernels, then the C3 step of the previous commit (_ab_old worktree) against the
# working tree with the new switches on / off, alternating; then the in-kernel GEMM trace of the new build.
cd "$(dirname "$0")/.."
B="--steps 20 --warmup 5 --skip-cpu --skip-eager --skip-recipe"
timeout 900 python -m pytest tests/test_ops_gpu.py -q -x -k "gemm" 2>&1 | tail -4 | cut -c1-300
timeout 900 python -m pytest tests/test_step_parity_gpu.py tests/test_ref_parity_gpu.py -q -x 2>&1 | tail -3 | cut -c1-300
for r in 1 2; do
  (cd _ab_old && timeout 300 python bench.py $B 2>/dev/null | python ../tools/bench_line.py old)
  timeout 300 python bench.py $B 2>/dev/null | python tools/bench_line.py new
  SLAM_STATIC_W=0 timeout 300 python bench.py $B 2>/dev/null | python tools/bench_line.py new_static0
  SLAM_THIN_CLUSTER=0 timeout 300 python bench.py $B 2>/dev/null | python tools/bench_line.py new_thin0
done
SLAM_B200_LIB=slam_llm_b200/libslam_b200_trace.so timeout 300 python tools/gemm_trace.py > gpurun_out/gemm_trace2.log 2> gpurun_out/gemm_trace2.err
tail -1 gpurun_out/gemm_trace2.log | cut -c1-300
