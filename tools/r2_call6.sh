#!/bin/bash
# Round 2 GPU call 6 (2 GPUs): two-warp rollout kernel (split) against the one-warp path, 256-bit stores of the sampling kernel,
# single-warp exchange push
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c6_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c6_pytest.log
{
echo "== default"; python tools/quick_check.py --no-smoke --only cfg3,cfg5,big,bigf32
echo "== one-warp AM"; MPPI_B200_SPLIT=0 python tools/quick_check.py --no-smoke --only cfg3,cfg5
} > gpurun_out/c6_quick.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 20 --no-secondary > gpurun_out/c6_bench2.json 2> gpurun_out/c6_bench2.err; echo "bench2 rc $?" >> gpurun_out/c6_bench2.err
timeout 600 python bench.py --steps 200 --warmup 20 --no-secondary --no-cpu-baseline > gpurun_out/c6_bench1.json 2> gpurun_out/c6_bench1.err
tail -4 gpurun_out/c6_pytest.log; tail -3 gpurun_out/c6_bench2.err; cat gpurun_out/c6_quick.log
