#!/bin/bash
# Source-level ncu captures of the SMALL kernels of a config-2 update (everything but the rollout), which are latency
# chains rather than throughput kernels:  gpurun --timeout 900 -- bash tools/profile_small.sh r2
# Text exports land in gpurun_out/<round>_small_<kernel>_{summary,lines}.txt; the .ncu-rep files stay on the box.
R=${1:-r2}
mkdir -p gpurun_out
python tools/prof_target.py cfg2 6 > /dev/null 2>&1 || exit 1
for k in k_finish k_weights k_gradient_reduce k_gradient k_sample_columns; do
  ncu --set full --clock-control none --import-source on --warp-sampling-interval 0 -k regex:"^${k}(<|\$)" --launch-skip 4 --launch-count 1 -f -o /tmp/${R}_small_$k python tools/prof_target.py cfg2 6 > gpurun_out/${R}_small_$k.log 2>&1
  python tools/ncu_summary.py /tmp/${R}_small_$k.ncu-rep > gpurun_out/${R}_small_${k}_summary.txt 2>&1
  python tools/ncu_lines.py /tmp/${R}_small_$k.ncu-rep 60 > gpurun_out/${R}_small_${k}_lines.txt 2>&1
done
ls -la gpurun_out | head -40
