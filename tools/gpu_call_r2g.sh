#!/bin/bash
# Round 2, GPU call G (2 GPUs): runs-unchanged tests, C2 at size against the oracle, N = 2 bench with the N-rank parity check.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_examples.py "tests/test_gpu_full_size.py::test_config_c2_four_million_atoms_against_the_oracle" -q --tb=short -p no:cacheprovider > gpurun_out/r2g_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2g_tests.log
tail -40 gpurun_out/r2g_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 60 --warmup 20 > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err
echo "bench n2 exit $?"
tail -c 3500 gpurun_out/r2g_bench_n2.json; grep -n "Error" -B2 -A6 gpurun_out/r2g_bench_n2.err | head -40
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q --tb=short -p no:cacheprovider > gpurun_out/r2g_multi.log 2>&1
tail -30 gpurun_out/r2g_multi.log
