#!/bin/bash
# GPU check of the diagnostic fused into the push: its tests, the per-step cost on the staged shapes, then configs[2]
# per GPU with and without it.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_fused_moments.py tests/test_gpu_boris.py -x -q -m gpu > gpurun_out/fused_tests.log 2>&1; echo "pytest rc $?"; tail -5 gpurun_out/fused_tests.log
python tools/run_fused_ke.py > gpurun_out/fused_ke_staged.jsonl 2> gpurun_out/fused_ke_staged.err; echo "run_fused_ke rc $?"; cat gpurun_out/fused_ke_staged.jsonl | cut -c1-330
OUT=gpurun_out/fused_ke_ab.txt; : > $OUT
for F in 1 0 1 0; do
  TURBOPY_B200_FUSED_MOMENTS=$F python bench.py --workload config3 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-extras 2> gpurun_out/fused_err_$F.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fused=$F', round(d['value']/1e9,2), 'G pushes/s', round(d['roofline']['frac'],4), round(d['ms_per_step'],4), 'ms/step launches', d['gpu_launches'], d['parity'], d['clocks']['reasons'])" >> $OUT 2>&1
done
cat $OUT
python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-extras 2>/dev/null | tail -1 | cut -c1-400
