mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu6.log
timeout 300 python tools/time_options.py > gpurun_out/options6.log 2>&1
timeout 300 python bench.py > gpurun_out/bench_B6.json 2> gpurun_out/bench_B6.err
timeout 900 python tools/parity_table.py 1024 > gpurun_out/parity_table.md 2> gpurun_out/parity_table.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches6.csv python bench.py --steps 2 --warmup 3 > gpurun_out/b_ncu6.log 2>&1
tail -3 gpurun_out/pytest_gpu6.log; head -3 gpurun_out/options6.log; tail -1 gpurun_out/options6.log; cat gpurun_out/bench_B6.json | head -c 300
