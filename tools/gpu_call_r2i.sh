#!/bin/bash
# Round 2, GPU call I (2 GPUs): branch-free pair arithmetic in the prototype (M = 256 and 384); N = 2 parity module.
mkdir -p gpurun_out
B=tools/micro/_bin/tile_force
timeout 300 $B 100 0.12 0 > gpurun_out/r2i_tile_100_m256.jsonl 2>&1
timeout 300 $B 100 0.12 4 > gpurun_out/r2i_tile_100_m384.jsonl 2>&1
grep -h -v build_base gpurun_out/r2i_tile_100_m256.jsonl gpurun_out/r2i_tile_100_m384.jsonl | cut -c1-200
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q --tb=short -p no:cacheprovider -k "md" > gpurun_out/r2i_multi.log 2>&1
tail -5 gpurun_out/r2i_multi.log | cut -c1-1500
