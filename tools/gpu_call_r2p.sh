#!/bin/bash
# Round 2, GPU call P (1 GPU): TMA staging in the prototype; where bench.py --workload dem dies.
mkdir -p gpurun_out
B=tools/micro/_bin/tile_force
timeout 300 $B 100 0.12 0 > gpurun_out/r2p_tile_100_m256.jsonl 2>&1
grep -h "fast_u8\|tma\|exact" gpurun_out/r2p_tile_100_m256.jsonl | cut -c1-200
timeout 600 python -X faulthandler bench.py --workload dem --steps 50 --warmup 5 --dem-settle 400 --no-cpu-baseline > gpurun_out/r2p_bench_dem.json 2> gpurun_out/r2p_bench_dem.err
echo "dem exit $?"; cut -c1-600 gpurun_out/r2p_bench_dem.json; tail -30 gpurun_out/r2p_bench_dem.err
