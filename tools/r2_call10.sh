#!/bin/bash
# Round 2 GPU call 10: rolled kinematics / RNEA / objective loops (instruction-fetch-bound kernels): parity + A/B timing
mkdir -p gpurun_out
CS=assistedmanipulation_b200/csrc
Q="python tools/quick_check.py --no-smoke"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c10_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c10_pytest.log
{
echo "== rolled (default)"; $Q --only cfg3,cfg5 --flips
echo "== rolled, one warp"; MPPI_B200_SPLIT=0 $Q --only cfg3,cfg5
echo "== rolled, loop-body FP64 solver"; MPPI_B200_LIB=$CS/libmppi_b200_vsu1.so $Q --only cfg3,cfg5
echo "== rolled, loop-body FP64 solver, one warp"; MPPI_B200_SPLIT=0 MPPI_B200_LIB=$CS/libmppi_b200_vsu1.so $Q --only cfg3,cfg5
echo "== unrolled"; MPPI_B200_LIB=$CS/libmppi_b200_vunrolled.so $Q --only cfg3,cfg5
echo "== block 128 one warp rolled"; MPPI_B200_SPLIT=0 MPPI_B200_AM_BLOCK=128 $Q --only cfg3,cfg5
echo "== block 32 one warp rolled"; MPPI_B200_SPLIT=0 MPPI_B200_AM_BLOCK=32 $Q --only cfg3,cfg5
} > gpurun_out/c10_ab.log 2>&1
tail -4 gpurun_out/c10_pytest.log; cat gpurun_out/c10_ab.log | cut -c1-200
