#!/bin/bash
# Round 2 GPU call 5 (2 GPUs): parity suite incl. the 2-GPU exchange tests on the folded exchange, timings of the new
# sampling / weighted-sum kernels, the 2-GPU bench line with its secondary block (sharded parity, cfg4 strong, cfg5).
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c5_gpus.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c5_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c5_pytest.log
{
echo "== default"; python tools/quick_check.py --no-smoke --only cfg2,cfg2f32,cfg3,big,bigf32
echo "== quads kernel"; MPPI_B200_SAMPLE_TILE=2 python tools/quick_check.py --no-smoke --only cfg2,big,bigf32
} > gpurun_out/c5_quick.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 20 > gpurun_out/c5_bench2.json 2> gpurun_out/c5_bench2.err; echo "bench2 rc $?" >> gpurun_out/c5_bench2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 20 --exchange nccl --no-secondary > gpurun_out/c5_bench2_nccl.json 2> gpurun_out/c5_bench2_nccl.err; echo "bench2 nccl rc $?" >> gpurun_out/c5_bench2_nccl.err
tail -4 gpurun_out/c5_pytest.log; tail -3 gpurun_out/c5_bench2.err; tail -3 gpurun_out/c5_bench2_nccl.err; cat gpurun_out/c5_quick.log
