#!/bin/bash
# Round 2 GPU call: A/B of the switches left unmeasured at the end of round 1, then ncu captures of the kernels as shipped.
# The .ncu-rep files stay on the box (3 x 25 MB exceeds the 64 MiB gpurun_out limit): only their text exports travel.
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
nvidia-smi -q -d PERFORMANCE > gpurun_out/r2_nvsmi_perf.txt 2>&1
bash tools/ab_next.sh > gpurun_out/ab_summary.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
for w in cfg2 cfg3 big; do
  python tools/prof_target.py $w 6 > gpurun_out/r2_prof_$w.log 2>&1 || continue
  ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2_launches_$w.csv python tools/prof_target.py $w 6 > /dev/null 2>&1
  $NCU --launch-skip 32 --launch-count 10 -f -o /tmp/r2_$w python tools/prof_target.py $w 6 > gpurun_out/r2_ncu_$w.log 2>&1
  python tools/ncu_summary.py /tmp/r2_$w.ncu-rep > gpurun_out/r2_${w}_summary.txt 2>&1
  python tools/ncu_lines.py /tmp/r2_$w.ncu-rep 60 > gpurun_out/r2_${w}_lines.txt 2>&1
  ncu -i /tmp/r2_$w.ncu-rep --page raw --csv > gpurun_out/r2_${w}_raw.csv 2>/dev/null
done
ls -la gpurun_out; du -sh gpurun_out
