#!/bin/bash
# Round 2, GPU call D: tile lists in the library: new parity tests, the whole GPU suite, tile on/off timings, the bench line.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tiles.py -x -q --tb=short -p no:cacheprovider > gpurun_out/r2d_tiles.log 2>&1
echo "tiles exit $?" >> gpurun_out/r2d_tiles.log
tail -40 gpurun_out/r2d_tiles.log
timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r2d_gpu_suite.log 2>&1
echo "suite exit $?" >> gpurun_out/r2d_gpu_suite.log
tail -40 gpurun_out/r2d_gpu_suite.log
timeout 300 python tools/bench_tiles.py 100 > gpurun_out/r2d_bench_tiles.json 2> gpurun_out/r2d_bench_tiles.err
cat gpurun_out/r2d_bench_tiles.json; tail -5 gpurun_out/r2d_bench_tiles.err
timeout 600 python bench.py --steps 100 --warmup 20 --no-dem > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
tail -c 2500 gpurun_out/r2d_bench.json; tail -5 gpurun_out/r2d_bench.err
