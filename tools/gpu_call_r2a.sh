#!/bin/bash
# Round 2, GPU call A: (1) tile prototype timings + ncu, (2) the whole GPU suite with the non-strict xfails removed.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2a_gpu.txt 2>&1
B=tools/micro/_bin/tile_force
timeout 200 $B 40 > gpurun_out/r2a_tile_40.jsonl 2>&1
timeout 400 $B 100 > gpurun_out/r2a_tile_100.jsonl 2>&1
tail -40 gpurun_out/r2a_tile_100.jsonl
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 200 $NCU -k regex:k_force_base -s 2 -c 1 -o gpurun_out/r2a_ncu_force_base $B 63 0.12 2 > gpurun_out/r2a_ncu_1.log 2>&1
timeout 200 $NCU -k regex:k_force_base -s 14 -c 1 -o gpurun_out/r2a_ncu_force_base_fma $B 63 0.12 2 > gpurun_out/r2a_ncu_2.log 2>&1
timeout 200 $NCU -k regex:k_force_tile -s 2 -c 1 -o gpurun_out/r2a_ncu_force_tile $B 63 0.12 2 > gpurun_out/r2a_ncu_3.log 2>&1
timeout 200 $NCU -k regex:k_force_tile -s 14 -c 1 -o gpurun_out/r2a_ncu_force_tile_fma $B 63 0.12 2 > gpurun_out/r2a_ncu_4.log 2>&1
timeout 200 $NCU -k regex:k_build_tile -s 2 -c 1 -o gpurun_out/r2a_ncu_build_tile $B 63 0.12 2 > gpurun_out/r2a_ncu_5.log 2>&1
timeout 200 $NCU -k regex:k_build_base -s 2 -c 1 -o gpurun_out/r2a_ncu_build_base $B 63 0.12 2 > gpurun_out/r2a_ncu_6.log 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r2a_gpu_suite.log 2>&1
echo "suite exit $?" >> gpurun_out/r2a_gpu_suite.log
tail -60 gpurun_out/r2a_gpu_suite.log
