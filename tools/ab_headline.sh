#!/bin/bash
# usage: tools/ab_headline.sh "<nvcc flags A>" "<nvcc flags B>" ...: rebuild with each flag set, headline bench at the driver's step count
for FLAGS in "$@"; do
  TURBOPY_B200_NVCC_FLAGS="$FLAGS" python -m turbopy_b200.build --force > /dev/null 2>&1
  for M in 1 0; do for R in 1 2 3; do
    TURBOPY_B200_NVCC_FLAGS="$FLAGS" TURBOPY_B200_MASK_KERNELS=$M python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-extras 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('flags [$FLAGS] mask $M K=20', round(d['value']/1e9,2), round(d['roofline']['frac'],4))"
  done; done
done
