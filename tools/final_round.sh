#!/bin/bash
# Round-end check on the GPU box: the whole GPU suite, the per-kernel ncu table of one eager C3 step, the default bench line.
cd "$(dirname "$0")/.."
TAG=${1:-r02c}
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4 | cut -c1-300
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
timeout 900 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/${TAG}_step_metrics.csv \
  python bench.py --ncu-step --graph 0 --skip-cpu --skip-eager --skip-recipe --skip-traffic > gpurun_out/${TAG}_ncu_step.log 2>&1
python tools/step_kernel_table.py gpurun_out/${TAG}_step_metrics.csv gpurun_out/${TAG}_step_kernels > /dev/null 2>&1
head -12 gpurun_out/${TAG}_step_kernels.txt | cut -c1-200
timeout 900 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
tail -1 gpurun_out/${TAG}_bench_n1.json | python tools/bench_line.py final
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
