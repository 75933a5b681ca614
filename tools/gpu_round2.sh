mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu2.log
timeout 300 python tools/time_variants.py > gpurun_out/variants.log 2>&1
timeout 300 python tools/time_options.py > gpurun_out/options.log 2>&1
timeout 300 python bench.py --workload F > gpurun_out/bench_F.json 2> gpurun_out/bench_F.err
timeout 300 python bench.py > gpurun_out/bench_B2.json 2> gpurun_out/bench_B2.err
tail -5 gpurun_out/pytest_gpu2.log; cat gpurun_out/variants.log gpurun_out/options.log; head -c 600 gpurun_out/bench_F.json; tail -3 gpurun_out/bench_F.err
