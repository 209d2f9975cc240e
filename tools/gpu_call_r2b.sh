#!/bin/bash
# Round 2, GPU call B: tile prototype with the segment fix (timings + ncu), and the existing union pair lists on 4 M atoms.
mkdir -p gpurun_out
B=tools/micro/_bin/tile_force
timeout 200 $B 40 > gpurun_out/r2b_tile_40.jsonl 2>&1
timeout 400 $B 100 > gpurun_out/r2b_tile_100.jsonl 2>&1
tail -30 gpurun_out/r2b_tile_100.jsonl
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 200 $NCU -k regex:k_force_tile -s 2 -c 1 -o gpurun_out/r2b_ncu_force_tile $B 63 0.12 0 > gpurun_out/r2b_ncu_3.log 2>&1
timeout 200 $NCU -k regex:k_force_tile -s 14 -c 1 -o gpurun_out/r2b_ncu_force_tile_fma $B 63 0.12 0 > gpurun_out/r2b_ncu_4.log 2>&1
timeout 200 $NCU -k regex:k_build_tile -s 2 -c 1 -o gpurun_out/r2b_ncu_build_tile $B 63 0.12 0 > gpurun_out/r2b_ncu_5.log 2>&1
timeout 300 python tools/bench_pair_lists.py 100 > gpurun_out/r2b_bench_pair_lists.json 2> gpurun_out/r2b_bench_pair_lists.err
cat gpurun_out/r2b_bench_pair_lists.json
