#!/bin/bash
# Round 2, GPU call 3B (1 GPU): ncu of the DEM contact kernels in the settled bed.
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:pb_k_dem_force -s 3000 -c 1 -o gpurun_out/r3b_ncu_dem_force python tools/profile_dem.py 3010 > gpurun_out/r3b_ncu_1.log 2>&1
timeout 600 $NCU -k regex:pb_k_dem_detect -s 3000 -c 1 -o gpurun_out/r3b_ncu_dem_detect python tools/profile_dem.py 3010 > gpurun_out/r3b_ncu_2.log 2>&1
tail -3 gpurun_out/r3b_ncu_1.log; ls -la gpurun_out/r3b_*
