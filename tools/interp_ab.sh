for U in 1 2 3; do
  TURBOPY_B200_NVCC_FLAGS=-DTP_INTERP_UNROLL=$U python -m turbopy_b200.build --force > /dev/null 2>&1
  echo "UNROLL $U"; TURBOPY_B200_NVCC_FLAGS=-DTP_INTERP_UNROLL=$U python tools/bench_kernels.py --quick 2>&1 | grep -i "interpolate1D"
done
