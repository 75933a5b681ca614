#!/bin/bash
# lane-distributed height-field narrowphase: parity tests, workload F bench line, phase breakdown
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hfield_terrain.py tests/test_convex_scene.py tests/test_terrain_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --workload F --steps 50 --warmup 10 > gpurun_out/hf_bench_F.json 2> gpurun_out/hf_bench_F.err
python -c "import json;d=json.load(open('gpurun_out/hf_bench_F.json'));print(d['value'], d['e2e']['value'], d.get('kernel_ms'), d['roofline'])"
B2_WORKLOAD=F B2SIM_LIB=mjlab_b200/csrc/variants/libb2sim_timing.so timeout 600 python tools/phase_breakdown.py 4096 100 > gpurun_out/hf_phase_F.txt 2>&1; tail -20 gpurun_out/hf_phase_F.txt
