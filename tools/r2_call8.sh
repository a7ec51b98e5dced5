#!/bin/bash
# Round 2 GPU call 8 (2 GPUs): exchange description in the parameter bank, device-side error flag; bench with / without the per-step host barrier
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/c8_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c8_pytest.log
for v in free barrier; do
FLAG=""; [ $v = barrier ] && FLAG="--step-barrier"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 300 --warmup 20 --no-secondary $FLAG > gpurun_out/c8_bench2_$v.json 2> gpurun_out/c8_bench2_$v.err
done
timeout 600 python bench.py --steps 300 --warmup 20 --no-secondary --no-cpu-baseline > gpurun_out/c8_bench1.json 2> gpurun_out/c8_bench1.err
tail -3 gpurun_out/c8_pytest.log
