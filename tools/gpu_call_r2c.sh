#!/bin/bash
# Round 2, GPU call C: tile prototype v2 (super-column tiles, cp.async staging, list prefetch): timings + ncu.
mkdir -p gpurun_out
B=tools/micro/_bin/tile_force
timeout 200 $B 40 > gpurun_out/r2c_tile_40.jsonl 2>&1
timeout 400 $B 100 > gpurun_out/r2c_tile_100.jsonl 2>&1
cat gpurun_out/r2c_tile_100.jsonl
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 200 $NCU -k regex:k_force_tile -s 14 -c 1 -o gpurun_out/r2c_ncu_force_tile $B 63 0.12 0 > gpurun_out/r2c_ncu_1.log 2>&1
timeout 200 $NCU -k regex:k_force_tile -s 26 -c 1 -o gpurun_out/r2c_ncu_force_tile_fma $B 63 0.12 0 > gpurun_out/r2c_ncu_2.log 2>&1
timeout 200 $NCU -k regex:k_build_tile -s 2 -c 1 -o gpurun_out/r2c_ncu_build_tile $B 63 0.12 0 > gpurun_out/r2c_ncu_3.log 2>&1
