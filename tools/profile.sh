#!/bin/bash
# The ncu evidence of a round, one GPU call:  gpurun --timeout 1500 -- bash tools/profile.sh r2
# For every workload of tools/prof_target.py: the launch list (gpu__time_duration), one `--set full` capture of a steady-state
# update exported to text (the .ncu-rep files are 25 MB each and stay on the box), and the metric pass that
# tools/ncu_counters.py turns into profiles/<round>_counters.json. Everything lands in gpurun_out/; copy what is to be
# judged into profiles/. LISTS="cfg2 cfg2_f32" FULL="cfg2" restrict the workloads of the two loops.
R=${1:-r2}
mkdir -p gpurun_out
nvidia-smi -q -d PERFORMANCE > gpurun_out/${R}_nvsmi_perf.txt 2>&1
M=$(python -c "import sys; sys.path.insert(0,'tools'); import ncu_counters; print(ncu_counters.METRICS)")
for w in ${LISTS:-cfg2 cfg2_f32 cfg3 cfg4_f64 cfg4_f32}; do
  python tools/prof_target.py $w 6 > gpurun_out/${R}_prof_$w.log 2>&1 || continue
  ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${R}_launches_$w.csv python tools/prof_target.py $w 6 > /dev/null 2>&1
  # (the small unsharded configurations run three kernels per update — rollout with its own sampling, weighted sum with the weights,
  # finish with the second stage — the others six; the window starts at the fourth update either way)
  case $w in cfg2*) SKIP=9; COUNT=4;; *) SKIP=18; COUNT=8;; esac
  ncu --csv --metrics $M --clock-control none --launch-skip $SKIP --launch-count $COUNT python tools/prof_target.py $w 6 > gpurun_out/${R}_counters_$w.csv 2> gpurun_out/${R}_counters_$w.err
done
for w in ${FULL:-cfg2 cfg3 cfg4_f64 cfg4_f32}; do
  case $w in cfg2*) SKIP=9; COUNT=3;; *) SKIP=18; COUNT=6;; esac
  ncu --set full --clock-control none --import-source on --launch-skip $SKIP --launch-count $COUNT -f -o /tmp/${R}_$w python tools/prof_target.py $w 6 > gpurun_out/${R}_ncu_$w.log 2>&1
  python tools/ncu_summary.py /tmp/${R}_$w.ncu-rep > gpurun_out/${R}_${w}_summary.txt 2>&1
  python tools/ncu_lines.py /tmp/${R}_$w.ncu-rep 40 > gpurun_out/${R}_${w}_lines.txt 2>&1
done
ls -la gpurun_out | head -60
