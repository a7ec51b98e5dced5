#!/bin/bash
# Round 2 GPU call 3: (1) re-rollout on the same kernel build as the grid it overlaps, (2) FP64 state path of the FP32 fast mode:
# timing and flip counts of the builds (nomix / state only / state path with the unrolled or the loop-body solver).
mkdir -p gpurun_out
Q="python tools/quick_check.py --no-smoke"
CS=assistedmanipulation_b200/csrc
{
echo "== default (level 2, unrolled solver, same-build re-rollout, reserve 200)"; $Q --only cfg2,cfg2f32,cfg3,cfg5,bigf32 --flips
echo "== other-build re-rollout"; MPPI_B200_OPTIMAL_OTHER_BUILD=1 $Q --only cfg2,cfg2f32
echo "== same build, reserve 0"; MPPI_B200_OPTIMAL_RESERVE_KB=0 $Q --only cfg2,cfg2f32
echo "== no optimal"; MPPI_B200_NO_OPTIMAL=1 $Q --only cfg2,cfg2f32,cfg3
echo "== loop-body FP64 solver"; MPPI_B200_LIB=$CS/libmppi_b200_vsu1.so $Q --only cfg3,cfg5 --flips
echo "== level 1"; MPPI_B200_LIB=$CS/libmppi_b200_vmix1.so $Q --only cfg3,cfg5
echo "== nomix"; MPPI_B200_LIB=$CS/libmppi_b200_vnomix.so $Q --only cfg3,cfg5
echo "== default block 128"; MPPI_B200_AM_BLOCK=128 $Q --only cfg3,cfg5
} > gpurun_out/c3_ab.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c3_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c3_pytest.log
tail -5 gpurun_out/c3_pytest.log
