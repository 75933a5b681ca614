#!/bin/bash
# one ncu --set full capture of the step kernel (MODE 3: step + convex routines) in workload F
mkdir -p gpurun_out
B2_WORKLOAD=F timeout 900 ncu --set full --clock-control none --import-source on -k regex:b2_step_kernel -s 1500 -c 1 -f -o gpurun_out/step_kernel_F python tools/profile_step.py 4096 ncu 200 > gpurun_out/ncu_F.log 2>&1
tail -3 gpurun_out/ncu_F.log; ls -la gpurun_out/
