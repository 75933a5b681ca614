mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py > gpurun_out/bench_B.json 2> gpurun_out/bench_B.err
timeout 300 python bench.py --workload C > gpurun_out/bench_C.json 2> gpurun_out/bench_C.err
timeout 300 python bench.py --workload E > gpurun_out/bench_E.json 2> gpurun_out/bench_E.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/b_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:b2_step_kernel -s 245 -c 1 -f -o gpurun_out/r02_step_kernel python tools/profile_step.py 4096 ncu > gpurun_out/ncu_full.log 2>&1
timeout 200 python tools/profile_step.py 4096 time > gpurun_out/profile_time.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_B.json; tail -12 gpurun_out/profile_time.log
