mkdir -p gpurun_out
timeout 300 python tools/sort_oracle.py > gpurun_out/sort_oracle.log 2>&1; cat gpurun_out/sort_oracle.log | tail -3
