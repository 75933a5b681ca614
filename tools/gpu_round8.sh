mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tracking_gpu.py -x -q > gpurun_out/pytest_gpu8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu8.log
timeout 300 python tools/debug_force_E.py > gpurun_out/debug_force_E.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:b2_step_kernel -s 245 -c 1 -f -o gpurun_out/r02_final_step_kernel python tools/profile_step.py 4096 ncu > gpurun_out/ncu_full8.log 2>&1
tail -3 gpurun_out/pytest_gpu8.log; cat gpurun_out/debug_force_E.log | cut -c1-600
