#!/bin/bash
# ncu of the fused-diagnostic instance of the hot kernel (pif shape): launch list of [diagnose(); gather_push()] steps
# and one --set full capture (FUSED_KE_ONLY restricts the run to the fused pif variant, so the name filter only sees the <MOM> instance).  Run only after tools/run_fused_ke.py exited 0 without ncu.
mkdir -p gpurun_out
export FUSED_KE_ONLY=pif:fused FUSED_KE_STEPS=10
python tools/run_fused_ke.py > gpurun_out/prof_fused_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof_fused_plain.log; exit 1; }
[ -n "$SKIP_LIST" ] || ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_fused_ke_launches.csv python tools/run_fused_ke.py > gpurun_out/prof_fused_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gather_push_tma_kernel -s 6 -c 1 -f -o gpurun_out/r2_fused_ke_mom python tools/run_fused_ke.py > gpurun_out/prof_fused_full.log 2>&1
ls -la gpurun_out/r2_fused_ke_mom.ncu-rep; tail -3 gpurun_out/prof_fused_full.log; head -5 gpurun_out/r2_fused_ke_launches.csv | cut -c1-300
