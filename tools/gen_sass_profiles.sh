#!/bin/bash
# Regenerates profiles/${R}_sass_model.txt and profiles/${R}_sass_phases.txt from the objects build() leaves in csrc/build
# (static: no GPU). Run from the repo root after `make -C assistedmanipulation_b200/csrc`.
set -e
R=${1:-r2}
T=tools/sass_cycles.py; O=assistedmanipulation_b200/csrc/build
# the range-reduction slow path of the FP64 sines / cosines (behind a forward branch taken only for |angle| >= 2^31)
FOLD=$(grep -n "if (!(fabs(a\[i\]) < 2147483648.0))" assistedmanipulation_b200/csrc/robot_fast.cuh | cut -d: -f1)
SK="--skip robot_fast.cuh:$FOLD"
{
echo "# Static issue-cycle model (tools/sass_cycles.py) of the rollout kernels as committed (round $R); one warp per SM sub-partition."
echo "# Calibration on B200: cfg2 kernel of r1_cfg2_final.ncu-rep models 7597 cycles/step, measures 9015; the kernel of r1_quick_check_after_model_work.log models 5372, measures 6509."
echo "# FP64 kernels: the range-reduction slow path of the shared sines / cosines (robot_fast.cuh:$FOLD, behind a forward branch taken only for |angle| >= 2^31) is left out (--skip)."
echo
echo "## lean reach-to-pose kernel, FP64, unrolled build — configs 2 and 4 (default for every rollout count)   [start: 3120 instructions, 2515 FP64; 2445 / 1956 / 5083 cycles before the joint placements' structural zeros]"
python $T $O/k_rollout_f64.o 'Li1ELb0ENS_11TrackPointPIdEELb1ELb0EEE' --lines 12 $SK
echo
echo "## the same kernel with the sampling warps of the noise chase (config 2 and every other small reach-to-pose rollout set; the 5-instruction inner loop is the chunk wait at the latch, modelled with 7 trips where it normally takes one)"
python $T $O/k_rollout_f64.o 'Li1ELb0ENS_11TrackPointPIdEELb1ELb1EEE' --lines 6 $SK
echo
echo "## same, loop-body build (MPPI_B200_BIG_FROM; config 2 ran it when the round's GPU numbers were taken)   [start: 3668 instructions, 2585 FP64, 7597 cycles; 2918 / 2173 / 5140 before the structural zeros]"
python $T $O/k_rollout_f64.o 'Li1ELb0ENS_11TrackPointPIdEELb0ELb0EEE' --lines 12 $SK
echo
echo "## config 3 / 5 kernel (FP32 assisted manipulation + energy tank)   [start: 8832 instructions, 18154 cycles; 5895 / 7629 before the self-collision pairs became one basic block; 5646 / 6773 with the solver's arm joints as a loop]"
python $T $O/k_rollout_f32.o 'IfLi4ELb0ENS_9AssistedPIfEELb0ELb0EEE' --lines 12
echo
echo "## lean reach-to-pose kernel in FP32, unrolled build (default)"
python $T $O/k_rollout_f32.o 'IfLi1ELb0ENS_11TrackPointPIfEELb1ELb0EEE' --fp64-issue 1
echo
echo "## same, loop-body build"
python $T $O/k_rollout_f32.o 'IfLi1ELb0ENS_11TrackPointPIfEELb0ELb0EEE' --fp64-issue 1
} > profiles/${R}_sass_model.txt
{
echo "# Call-site attribution of one rollout step (tools/sass_cycles.py --phases rollout_core.cuh): instructions and stall cycles per call made by"
echo "# rollout_franka, through the inline chains of the SASS line table (nvdisasm -gi). Static, no GPU; kernels as committed."
echo
echo "## lean reach-to-pose kernel, FP64, unrolled build (configs 2 and 4)"
python $T $O/k_rollout_f64.o 'k_rolloutIdLi1ELb0ENS_11TrackPointPIdEELb1ELb0' --phases rollout_core.cuh --phase-depth 1 $SK
echo
echo "## same, loop-body build"
python $T $O/k_rollout_f64.o 'k_rolloutIdLi1ELb0ENS_11TrackPointPIdEELb0ELb0' --phases rollout_core.cuh --phase-depth 1 $SK
echo
echo "## config 3 / 5 kernel (FP32 assisted manipulation + energy tank)"
python $T $O/k_rollout_f32.o 'k_rolloutIfLi4ELb0ENS_9AssistedPIfEELb0ELb0' --phases rollout_core.cuh --phase-depth 2 --fp64-issue 1 | grep -v "robot_fast.cuh"
} > profiles/${R}_sass_phases.txt
grep -H "per step" profiles/${R}_sass_model.txt
