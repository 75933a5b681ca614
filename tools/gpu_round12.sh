mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_hfield_terrain.py tests/test_convex_scene.py -m gpu -x -q > gpurun_out/pytest_gpu12.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu12.log
timeout 300 python bench.py --workload F > gpurun_out/bench_F12.json 2> gpurun_out/bench_F12.err
tail -3 gpurun_out/pytest_gpu12.log; head -c 300 gpurun_out/bench_F12.json
