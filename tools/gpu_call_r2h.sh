#!/bin/bash
# Round 2, GPU call H (2 GPUs): prototype sweep (M / CAP, conflict-aware list order) + ncu; N = 2 parity (bench + pytest).
mkdir -p gpurun_out
B=tools/micro/_bin/tile_force
timeout 600 $B 100 > gpurun_out/r2h_tile_100.jsonl 2>&1
grep -v build_base gpurun_out/r2h_tile_100.jsonl | cut -c1-260
NCU="ncu --set full --clock-control none --import-source on -f"
# variant 0: build x5, force x12, prefetch x12, fma2 x12, reorder x5, reordered x12, fma2 reordered x12
timeout 200 $NCU -k regex:k_force_tile -s 50 -c 1 -o gpurun_out/r2h_ncu_force_tile_fma_reordered $B 63 0.12 0 > gpurun_out/r2h_ncu_1.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 60 --warmup 20 > gpurun_out/r2h_bench_n2.json 2> gpurun_out/r2h_bench_n2.err
echo "bench n2 exit $?"
tail -c 3500 gpurun_out/r2h_bench_n2.json; grep -n "Error" -B2 -A6 gpurun_out/r2h_bench_n2.err | head -40
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q --tb=short -p no:cacheprovider > gpurun_out/r2h_multi.log 2>&1
tail -5 gpurun_out/r2h_multi.log | cut -c1-800
