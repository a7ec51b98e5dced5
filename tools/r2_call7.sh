#!/bin/bash
# Round 2 GPU call 7 (2 GPUs): exchange push with 256 workers; weighted-sum kernel by group count; split heuristic
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_variants.py tests/test_gpu_batch.py -m gpu -x -q > gpurun_out/c7_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c7_pytest.log
python tools/quick_check.py --no-smoke --only cfg2,cfg2f32,cfg5,bigf32 > gpurun_out/c7_quick.log 2>&1
for i in 1 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$i bench.py --gpus 2 --steps 300 --warmup 20 --no-secondary > gpurun_out/c7_bench2_$i.json 2> gpurun_out/c7_bench2_$i.err
timeout 600 python bench.py --steps 300 --warmup 20 --no-secondary --no-cpu-baseline > gpurun_out/c7_bench1_$i.json 2> gpurun_out/c7_bench1_$i.err
done
tail -3 gpurun_out/c7_pytest.log; cat gpurun_out/c7_quick.log
