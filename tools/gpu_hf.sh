#!/bin/bash
# height-field path: GPU test suite, workload F / B bench lines, phase breakdown of F (timing variant)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --workload F --steps 50 --warmup 10 > gpurun_out/hf_bench_F.json 2> gpurun_out/hf_bench_F.err
python -c "import json;d=json.load(open('gpurun_out/hf_bench_F.json'));print('F', d['value'], d['e2e']['value'], d['roofline']['kernel_ms'])"
timeout 600 python bench.py --steps 50 --warmup 10 > gpurun_out/hf_bench_B.json 2> gpurun_out/hf_bench_B.err
python -c "import json;d=json.load(open('gpurun_out/hf_bench_B.json'));print('B', d['value'], d['e2e']['value'], d['roofline']['kernel_ms'])"
B2_WORKLOAD=F B2SIM_LIB=mjlab_b200/csrc/variants/libb2sim_timing.so timeout 600 python tools/phase_breakdown.py 4096 100 > gpurun_out/hf_phase_F.txt 2>&1; tail -26 gpurun_out/hf_phase_F.txt
