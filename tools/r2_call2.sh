#!/bin/bash
# Round 2 GPU call 2: parity suite, then A/B timing of (1) the SM reservation of the side-stream re-rollout, (2) lockstep
# barriers / block size of the assisted-manipulation kernel, (3) the FP64-carried state of the FP32 fast mode.
mkdir -p gpurun_out
Q="python tools/quick_check.py --no-smoke"
CS=assistedmanipulation_b200/csrc
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c2_pytest.log
{
echo "== default"; $Q --only cfg2,cfg2f32
echo "== reserve 0"; MPPI_B200_OPTIMAL_RESERVE_KB=0 $Q --only cfg2,cfg2f32
echo "== no optimal"; MPPI_B200_NO_OPTIMAL=1 $Q --only cfg2,cfg2f32
echo "== cfg3 default"; $Q --only cfg3,cfg5
for blk in 64 128; do for ls in 0 1 2 3; do
echo "== cfg3 block $blk lockstep $ls"; MPPI_B200_AM_BLOCK=$blk MPPI_B200_LOCKSTEP=$ls $Q --only cfg3,cfg5
done; done
echo "== cfg3 block 32 lockstep 0"; MPPI_B200_AM_BLOCK=32 $Q --only cfg3
echo "== cfg3 block 256 lockstep 3"; MPPI_B200_AM_BLOCK=256 MPPI_B200_LOCKSTEP=3 $Q --only cfg3
echo "== nomix"; MPPI_B200_LIB=$CS/libmppi_b200_vnomix.so $Q --only cfg3,cfg2f32,bigf32
echo "== mixed"; $Q --only bigf32
echo "== flips mixed"; $Q --only none --flips
echo "== flips nomix"; MPPI_B200_LIB=$CS/libmppi_b200_vnomix.so $Q --only none --flips
} > gpurun_out/c2_ab.log 2>&1
tail -5 gpurun_out/c2_pytest.log
