mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_fin.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_fin.log
for w in B C E F; do timeout 300 python bench.py --workload $w > gpurun_out/bench_${w}_fin.json 2> gpurun_out/bench_${w}_fin.err; done
timeout 300 python tools/time_options.py > gpurun_out/options_fin.log 2>&1
tail -4 gpurun_out/pytest_gpu_fin.log; for w in B C E F; do head -c 200 gpurun_out/bench_${w}_fin.json; echo; done; head -3 gpurun_out/options_fin.log
