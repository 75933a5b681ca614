mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_ls2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_ls2.log
timeout 900 python tools/parity_table.py 1024 > gpurun_out/parity_table.md 2> gpurun_out/parity_table.err
tail -6 gpurun_out/pytest_gpu_ls2.log; grep "| default |" gpurun_out/parity_table.md | grep "qacc \|qvel"
