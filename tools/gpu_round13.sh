mkdir -p gpurun_out
for w in E F; do B2_WORKLOAD=$w B2SIM_LIB=mjlab_b200/csrc/variants/libb2sim_timing.so timeout 300 python tools/phase_breakdown.py > gpurun_out/phases_$w.log 2>&1; done
cat gpurun_out/phases_E.log gpurun_out/phases_F.log
