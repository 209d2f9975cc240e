#!/bin/bash
# Round 2, GPU call E: flattened tile build + prefetched epilogue: tests, timings, ncu of the two tile kernels in the library.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tiles.py tests/test_gpu_props.py tests/test_gpu_md.py -x -q --tb=short -p no:cacheprovider > gpurun_out/r2e_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2e_tests.log
tail -15 gpurun_out/r2e_tests.log
timeout 300 python tools/bench_tiles.py 100 > gpurun_out/r2e_bench_tiles.json 2> gpurun_out/r2e_bench_tiles.err
cat gpurun_out/r2e_bench_tiles.json; tail -5 gpurun_out/r2e_bench_tiles.err
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:pb_k_tile_lj -s 30 -c 1 -o gpurun_out/r2e_ncu_tile_lj python tools/prof_md.py 63 45 > gpurun_out/r2e_ncu_1.log 2>&1
timeout 300 $NCU -k regex:pb_k_tile_build -s 1 -c 1 -o gpurun_out/r2e_ncu_tile_build python tools/prof_md.py 63 25 > gpurun_out/r2e_ncu_2.log 2>&1
ls -la gpurun_out/r2e*.ncu-rep
