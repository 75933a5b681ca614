mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu3.log
timeout 300 python tools/time_variants.py > gpurun_out/variants3.log 2>&1
timeout 300 python tools/time_options.py > gpurun_out/options3.log 2>&1
tail -5 gpurun_out/pytest_gpu3.log; cat gpurun_out/variants3.log gpurun_out/options3.log
