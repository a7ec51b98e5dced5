#!/bin/bash
# The ncu evidence of a round, one GPU call:  gpurun --timeout 1500 -- bash tools/profile.sh r2
# For every workload of tools/prof_target.py: the launch list (gpu__time_duration), one `--set full` capture of a steady-state
# update exported to text (the .ncu-rep files are 25 MB each and stay on the box), and the metric pass that
# tools/ncu_counters.py turns into profiles/<round>_counters.json. Everything lands in gpurun_out/; copy what is to be
# judged into profiles/.
R=${1:-r2}
mkdir -p gpurun_out
nvidia-smi -q -d PERFORMANCE > gpurun_out/${R}_nvsmi_perf.txt 2>&1
M=$(python -c "import sys; sys.path.insert(0,'tools'); import ncu_counters; print(ncu_counters.METRICS)")
for w in cfg2 cfg2_f32 cfg3 cfg4_f64 cfg4_f32; do
  python tools/prof_target.py $w 6 > gpurun_out/${R}_prof_$w.log 2>&1 || continue
  ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${R}_launches_$w.csv python tools/prof_target.py $w 6 > /dev/null 2>&1
  ncu --csv --metrics $M --clock-control none --launch-skip 18 --launch-count 8 python tools/prof_target.py $w 6 > gpurun_out/${R}_counters_$w.csv 2> gpurun_out/${R}_counters_$w.err
done
for w in cfg2 cfg3 cfg4_f64 cfg4_f32; do
  ncu --set full --clock-control none --import-source on --launch-skip 18 --launch-count 6 -f -o /tmp/${R}_$w python tools/prof_target.py $w 6 > gpurun_out/${R}_ncu_$w.log 2>&1
  python tools/ncu_summary.py /tmp/${R}_$w.ncu-rep > gpurun_out/${R}_${w}_summary.txt 2>&1
  python tools/ncu_lines.py /tmp/${R}_$w.ncu-rep 40 > gpurun_out/${R}_${w}_lines.txt 2>&1
done
ls -la gpurun_out | head -60
