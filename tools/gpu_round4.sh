mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu4.log
timeout 400 python tools/time_variants.py > gpurun_out/variants4.log 2>&1
timeout 300 python tools/time_options.py > gpurun_out/options4.log 2>&1
timeout 300 python bench.py --workload F > gpurun_out/bench_F.json 2> gpurun_out/bench_F.err
tail -5 gpurun_out/pytest_gpu4.log; cat gpurun_out/variants4.log gpurun_out/options4.log; head -c 300 gpurun_out/bench_F.json
