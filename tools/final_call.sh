#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/f_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; echo "smoke rc $?" >> gpurun_out/f_smoke.log
timeout 900 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc $?" >> gpurun_out/f_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err
bash tools/profile.sh r2 > gpurun_out/f_profile.log 2>&1
tail -3 gpurun_out/f_pytest.log; cat gpurun_out/f_smoke.log | tail -2; tail -2 gpurun_out/f_bench.err
