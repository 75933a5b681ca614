mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu9.log
timeout 900 python tools/parity_table.py 1024 > gpurun_out/parity_table.md 2> gpurun_out/parity_table.err
timeout 300 python bench.py > gpurun_out/bench_B9.json 2> gpurun_out/bench_B9.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref9.json 2> gpurun_out/bench_ref9.err
tail -3 gpurun_out/pytest_gpu9.log; grep "| default | E | forward | contact_force" gpurun_out/parity_table.md; head -c 400 gpurun_out/bench_B9.json; echo; head -c 600 gpurun_out/bench_ref9.json
