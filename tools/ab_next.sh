#!/bin/bash
# A/B of the changes made without GPU time at the end of round 1. One GPU call:
#   gpurun --timeout 900 -- bash tools/ab_next.sh
# Logs land in gpurun_out/ab_*.log; the library reads the switches once per process, hence one process per variant.
mkdir -p gpurun_out
python tools/quick_check.py                                  > gpurun_out/ab_default.log 2>&1
MPPI_B200_BIG_FROM=12288 python tools/quick_check.py         > gpurun_out/ab_loop_body_below_12k.log 2>&1
MPPI_B200_BIG_FROM=1000000000 python tools/quick_check.py    > gpurun_out/ab_loop_body_always.log 2>&1
MPPI_B200_SAMPLE_TILE=1 python tools/quick_check.py          > gpurun_out/ab_sample_tile.log 2>&1
grep -H "device us\|replay" gpurun_out/ab_*.log
