#!/bin/bash
# Round 2, GPU call O (2 GPUs): everything multi-GPU that has never run (N-rank props, DEM N = 2), DEM bench line at N = 1 and 2,
# LJ bench at N = 2 with the parity block.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_props.py -q --tb=short -p no:cacheprovider -k "multi or nranks or n_rank or world" > gpurun_out/r2o_multi.log 2>&1
echo "multi exit $?" >> gpurun_out/r2o_multi.log; tail -6 gpurun_out/r2o_multi.log | cut -c1-600
timeout 900 python bench.py --workload dem --steps 200 --warmup 20 > gpurun_out/r2o_bench_dem_n1.json 2> gpurun_out/r2o_bench_dem_n1.err
echo "dem n1 exit $?"; cut -c1-1800 gpurun_out/r2o_bench_dem_n1.json; tail -3 gpurun_out/r2o_bench_dem_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --workload dem --gpus 2 --steps 200 --warmup 20 > gpurun_out/r2o_bench_dem_n2.json 2> gpurun_out/r2o_bench_dem_n2.err
echo "dem n2 exit $?"; cut -c1-900 gpurun_out/r2o_bench_dem_n2.json; tail -3 gpurun_out/r2o_bench_dem_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 100 --warmup 20 > gpurun_out/r2o_bench_n2.json 2> gpurun_out/r2o_bench_n2.err
echo "lj n2 exit $?"; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2o_bench_n2.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "e2e")}, d.get("parity_nranks"), d["roofline"]["frac"], d.get("reneighbor"))
except Exception as e:
    print("no line", e)
PY
tail -3 gpurun_out/r2o_bench_n2.err
