mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu5.log
timeout 300 python tools/time_options.py > gpurun_out/options5.log 2>&1
timeout 300 python bench.py > gpurun_out/bench_B5.json 2> gpurun_out/bench_B5.err
timeout 900 python tools/parity_table.py 1024 > gpurun_out/parity_table.md 2> gpurun_out/parity_table.err
tail -5 gpurun_out/pytest_gpu5.log; head -3 gpurun_out/options5.log; tail -2 gpurun_out/options5.log; cat gpurun_out/bench_B5.json | head -c 400
