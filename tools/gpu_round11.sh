mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu11.log
timeout 300 python tools/time_options.py > gpurun_out/options11.log 2>&1
timeout 300 python bench.py > gpurun_out/bench_B11.json 2> gpurun_out/bench_B11.err
tail -3 gpurun_out/pytest_gpu11.log; head -3 gpurun_out/options11.log; tail -3 gpurun_out/options11.log; head -c 300 gpurun_out/bench_B11.json
