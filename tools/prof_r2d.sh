set -x
cd $GRAFT_REPO_ROOT
python tools/run_interp.py --n 16777216 --ng 1025 --iters 4 > /dev/null 2>&1 || exit 1
python tools/run_gather6.py --ng 1025 --n 16777216 --order random --formula create_interpolator --steps 4 > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:interp_cells_ring -s 3 -c 1 -f -o /tmp/p_interp python tools/run_interp.py --n 16777216 --ng 1025 --iters 4 > gpurun_out/r2d_ncu_interp.log 2>&1
ncu -i /tmp/p_interp.ncu-rep --page raw --csv > gpurun_out/r2d_interp_ring_raw.csv 2>/dev/null
ncu -i /tmp/p_interp.ncu-rep --page source --csv > gpurun_out/r2d_interp_ring_src.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:gather_push_tma -s 3 -c 1 -f -o /tmp/p_fa python tools/run_gather6.py --ng 1025 --n 16777216 --order random --formula create_interpolator --steps 4 > gpurun_out/r2d_ncu_fa.log 2>&1
ncu -i /tmp/p_fa.ncu-rep --page raw --csv > gpurun_out/r2d_formula_a_raw.csv 2>/dev/null
ncu -i /tmp/p_fa.ncu-rep --page source --csv > gpurun_out/r2d_formula_a_src.csv 2>/dev/null
ls -la gpurun_out/r2d_*
