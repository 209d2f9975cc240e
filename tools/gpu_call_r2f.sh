#!/bin/bash
# Round 2, GPU call F (2 GPUs): tile build with 1-byte slot meta; tests; timings; N = 2 bench with the N-rank parity check.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tiles.py tests/test_gpu_md.py tests/test_gpu_edge_cases.py -x -q --tb=short -p no:cacheprovider > gpurun_out/r2f_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2f_tests.log
tail -15 gpurun_out/r2f_tests.log
timeout 300 python tools/bench_tiles.py 100 > gpurun_out/r2f_bench_tiles.json 2> gpurun_out/r2f_bench_tiles.err
cat gpurun_out/r2f_bench_tiles.json; tail -5 gpurun_out/r2f_bench_tiles.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 60 --warmup 20 > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err
echo "bench n2 exit $?"
tail -c 3000 gpurun_out/r2f_bench_n2.json; tail -30 gpurun_out/r2f_bench_n2.err
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q --tb=short -p no:cacheprovider > gpurun_out/r2f_multi.log 2>&1
tail -30 gpurun_out/r2f_multi.log
