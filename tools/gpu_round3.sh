mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu3.log
timeout 300 python tools/time_options.py > gpurun_out/options3.log 2>&1
B2SIM_LIB=mjlab_b200/csrc/variants/libb2sim_timing.so timeout 300 python tools/phase_breakdown.py > gpurun_out/phases3.log 2>&1
timeout 300 python tools/workload_stats.py > gpurun_out/workload3.log 2>&1
timeout 900 python tools/parity_table.py 1024 > gpurun_out/parity_table.md 2> gpurun_out/parity_table.err
timeout 300 python bench.py --workload F > gpurun_out/bench_F.json 2> gpurun_out/bench_F.err
tail -5 gpurun_out/pytest_gpu3.log; cat gpurun_out/options3.log; tail -40 gpurun_out/phases3.log; tail -30 gpurun_out/workload3.log
