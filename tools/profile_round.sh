#!/bin/bash
# Round profiling recipe (run under gpurun, ONE GPU): launch list of a full C3 step, dram traffic of every GEMM launch,
# and full-set captures of the GEMM (single-CTA and CTA-pair), FMHA and attention-backward kernels.
# Outputs land in gpurun_out/ (copy the summaries to profiles/).
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --steps 1 --warmup 3 --skip-cpu"
LPS=${LAUNCHES_PER_STEP:-1325}
# skip the 3 forced warm-up steps, then capture one timed step
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s $((3 * LPS)) -c $((LPS + 5)) --csv --log-file $OUT/${TAG}_launches.csv $B > /dev/null 2>&1
python tools/ncu_summary.py $OUT/${TAG}_launches.csv > $OUT/${TAG}_launches_summary.txt 2>/dev/null
timeout 500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:gemm_tcgen05 -s 1400 -c 470 --csv --log-file $OUT/${TAG}_gemm_traffic.csv $B > /dev/null 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_kernel -s 1500 -c 3 -o $OUT/${TAG}_prof_gemm $B > /dev/null 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_pair -s 300 -c 3 -o $OUT/${TAG}_prof_gemm_pair $B > /dev/null 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:fmha_fwd_tc -s 30 -c 4 -o $OUT/${TAG}_prof_fmha $B > /dev/null 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_kernel -s 5 -c 1 -o $OUT/${TAG}_prof_attn_bwd $B > /dev/null 2>&1
ls -la $OUT | tail -12
