mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu7.log
timeout 300 python bench.py --workload C > gpurun_out/bench_C7.json 2> gpurun_out/bench_C7.err
timeout 300 python bench.py --workload E > gpurun_out/bench_E7.json 2> gpurun_out/bench_E7.err
timeout 300 python bench.py --workload F > gpurun_out/bench_F7.json 2> gpurun_out/bench_F7.err
tail -4 gpurun_out/pytest_gpu7.log; head -c 250 gpurun_out/bench_C7.json; echo; head -c 250 gpurun_out/bench_E7.json; echo; head -c 250 gpurun_out/bench_F7.json; tail -3 gpurun_out/bench_C7.err
