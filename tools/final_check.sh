#!/bin/bash
# Round-end rehearsal on one B200: bench A-B at the driver's settings (idle gap in front of the timed region, eager vs
# CUDA graph), the GPU test suite, smoke(), both bench arms.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
OUT=gpurun_out/final_ab.txt; : > $OUT
line() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e9,2), 'G pushes/s', round(d['roofline']['frac'],4), 'launches', d['gpu_launches'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"; }
for R in ${AB_RUNS:-1 2 3}; do
  TURBOPY_B200_BENCH_GAP_MS=50 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-extras --no-graph 2>/dev/null | tail -1 | line "gap50 eager" >> $OUT
  TURBOPY_B200_BENCH_GAP_MS=0  python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-extras --no-graph 2>/dev/null | tail -1 | line "gap0  eager" >> $OUT
  TURBOPY_B200_BENCH_GAP_MS=0  python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-extras --graph 2>/dev/null | tail -1 | line "gap0  graph" >> $OUT
done
cat $OUT
python -m pytest tests -m gpu -x -q -rs > gpurun_out/final_gpu_tests.log 2>&1; echo "pytest rc $?" | tee -a $OUT; tail -3 gpurun_out/final_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc $?" | tee -a $OUT; tail -3 gpurun_out/final_smoke.log
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_bench_ref_n1.json 2> gpurun_out/final_bench_ref_n1.err; echo "ref rc $?" | tee -a $OUT
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench rc $?" | tee -a $OUT
tail -c 600 gpurun_out/final_bench_n1.json
