python -m pytest tests/test_gpu_fused_moments.py tests/test_gpu_boris.py -x -q -m gpu > gpurun_out/fused_tests2.log 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/fused_tests2.log
python tools/run_fused_ke.py > gpurun_out/fused_ke_staged.jsonl 2> gpurun_out/fused_ke_staged.err; echo rc $?; cut -c1-300 gpurun_out/fused_ke_staged.jsonl
