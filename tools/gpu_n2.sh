mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 60 --warmup 10 > gpurun_out/bench_B_n2.json 2> gpurun_out/bench_B_n2.err
timeout 300 python bench.py --steps 60 --warmup 10 > gpurun_out/bench_B_n1_same_box.json 2> gpurun_out/bench_B_n1_same_box.err
timeout 300 python tools/sort_oracle.py > gpurun_out/sort_oracle.log 2>&1
grep '^{' gpurun_out/bench_B_n2.json | head -c 300; echo; head -c 300 gpurun_out/bench_B_n1_same_box.json; echo; tail -2 gpurun_out/sort_oracle.log
