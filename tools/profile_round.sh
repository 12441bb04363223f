#!/bin/bash
# Round profiling recipe (run under gpurun, ONE GPU): launch list of a full C3 step, dram traffic of every GEMM launch,
# and full-set captures of the GEMM and attention kernels.  Outputs land in gpurun_out/ (copy summaries to profiles/).
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --steps 1 --warmup 3 --skip-cpu"
# launches per step ~1300: skip 3 warm-up steps (the >=3 forced warm-ups) then capture one timed step
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 3900 -c 1400 --csv --log-file $OUT/${TAG}_launches.csv $B > /dev/null 2>&1
python tools/ncu_summary.py $OUT/${TAG}_launches.csv > $OUT/${TAG}_launches_summary.txt 2>/dev/null
timeout 500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:gemm_tcgen05 -s 1400 -c 470 --csv --log-file $OUT/${TAG}_gemm_traffic.csv $B > /dev/null 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 1500 -c 3 -o $OUT/${TAG}_prof_gemm $B > /dev/null 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 40 -c 4 -o $OUT/${TAG}_prof_attn $B > /dev/null 2>&1
ls -la $OUT
