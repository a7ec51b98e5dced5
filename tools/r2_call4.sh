#!/bin/bash
# Round 2 GPU call 4: GPU parity suite on the new build, the full bench line, ncu metric passes for profiles/r2_counters.json
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c4_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c4_pytest.log
python tools/quick_check.py --no-smoke --only cfg2,cfg2f32,cfg3,cfg5,big,bigf32 > gpurun_out/c4_quick.log 2>&1
timeout 900 python bench.py > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err; echo "bench rc $?" >> gpurun_out/c4_bench.err
M=$(python -c "import sys; sys.path.insert(0,'tools'); import ncu_counters; print(ncu_counters.METRICS)")
for w in cfg2 cfg2_f32 cfg3 cfg4_f64 cfg4_f32; do
  ncu --csv --metrics $M --clock-control none --launch-skip 18 --launch-count 8 python tools/prof_target.py $w 6 > gpurun_out/r2_counters_$w.csv 2> gpurun_out/r2_counters_$w.err
done
tail -3 gpurun_out/c4_pytest.log; tail -2 gpurun_out/c4_bench.err; cat gpurun_out/c4_quick.log
