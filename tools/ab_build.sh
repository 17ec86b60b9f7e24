#!/bin/bash
# usage: tools/ab_build.sh "<nvcc flags A>" "<nvcc flags B>" ...  -- rebuilds the library with each flag set (on the GPU
# box) and prints the headline and the six-component figures (A-B runs of compile-time switches).
for FLAGS in "$@"; do
  TURBOPY_B200_NVCC_FLAGS="$FLAGS" python -m turbopy_b200.build --force > /dev/null 2>&1
  echo "=== FLAGS: '$FLAGS'"
  TURBOPY_B200_NVCC_FLAGS="$FLAGS" python bench.py --no-e2e --no-cpu-baseline --no-extras --steps 200 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('headline', round(d['value']/1e9,2), round(d['roofline']['frac'],4))"
  for f in interp interp1d_nd; do
    TURBOPY_B200_NVCC_FLAGS="$FLAGS" python tools/run_gather6.py --ng 1025 --order random --formula $f 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ng1025 random', d['formula'], round(d['frac_of_hbm_peak'],4))"
  done
  TURBOPY_B200_NVCC_FLAGS="$FLAGS" python tools/run_gather6.py --ng 1048577 --n 125000000 --order sorted --steps 100 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('config3 100 steps', round(d['frac_of_hbm_peak'],4))"
done
