mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_ls.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_ls.log
timeout 300 python bench.py > gpurun_out/bench_B_ls.json 2> gpurun_out/bench_B_ls.err
timeout 300 python bench.py --workload C > gpurun_out/bench_C_ls.json 2> gpurun_out/bench_C_ls.err
timeout 300 python bench.py --workload E > gpurun_out/bench_E_ls.json 2> gpurun_out/bench_E_ls.err
tail -3 gpurun_out/pytest_gpu_ls.log; for w in B C E; do head -c 230 gpurun_out/bench_${w}_ls.json; echo; done
