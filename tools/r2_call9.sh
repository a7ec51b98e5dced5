#!/bin/bash
# Round 2 GPU call 9 (2 GPUs): whole GPU suite on the flag-in-data exchange + a timeout check
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c9_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c9_pytest.log
tail -5 gpurun_out/c9_pytest.log
