#!/bin/bash
# Round-2 profiling recipe (run under gpurun, ONE GPU).  Outputs in gpurun_out/ (summaries are copied to profiles/ by hand).
#  1. every launch of ONE eager C3 step with time + DRAM bytes + tensor-pipe %   -> ${TAG}_step_metrics.csv  (tools/step_kernel_table.py)
#  2. --set full captures (source-level, for reading here) of the HBM-bound kernels the north star names and of the attention kernels
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --ncu-step --graph 0 --skip-cpu --skip-eager --skip-recipe --skip-traffic"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
timeout 900 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file $OUT/${TAG}_step_metrics.csv $B > $OUT/${TAG}_ncu_step.log 2>&1
python tools/step_kernel_table.py $OUT/${TAG}_step_metrics.csv $OUT/${TAG}_step_kernels > /dev/null 2>&1
for K in logmel_kernel layernorm_reg rmsnorm_fwd_reg rmsnorm_bwd_reg rope_kernel cross_entropy_kernel adamw_kernel embed_merge_kernel im2col_kernel; do
  timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$K -c 2 -o $OUT/${TAG}_prof_$K $B > /dev/null 2>&1
  ncu -i $OUT/${TAG}_prof_$K.ncu-rep --page details --csv 2>/dev/null | grep -E "DRAM Throughput|Memory Throughput|Duration|Achieved Occupancy|Registers Per|L2 Cache Throughput|Compute \(SM\) Throughput" > $OUT/${TAG}_prof_$K.txt
done
ls -la $OUT | tail -30
