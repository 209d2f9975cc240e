#!/bin/bash
# Round 2, GPU call Q (2 GPUs): TMA-staged tiles from the cell-ordered mirror in the library: MD suites, bench_tiles, N = 2 parity + bench, DEM bench.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_tiles.py tests/test_gpu_md.py tests/test_gpu_edge_cases.py tests/test_gpu_props.py tests/test_gpu_multi.py -q --tb=short -p no:cacheprovider -x > gpurun_out/r2q_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2q_tests.log
tail -15 gpurun_out/r2q_tests.log | cut -c1-400
timeout 600 python tools/bench_tiles.py 100 > gpurun_out/r2q_bench_tiles.json 2> gpurun_out/r2q_bench_tiles.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2q_bench_tiles.json"))
for k, v in d.items():
    if isinstance(v, dict):
        print(k, round(v["ms_per_step"], 4), {a: round(b["ms_per_call"], 4) for a, b in v["stages"].items()})
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 100 --warmup 20 > gpurun_out/r2q_bench_n2.json 2> gpurun_out/r2q_bench_n2.err
echo "lj n2 exit $?"; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2q_bench_n2.json"))
    print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], (d.get("parity_nranks") or {}).get("ok"), d["roofline"]["frac"], d.get("reneighbor"))
except Exception as e:
    print("no line", e)
PY
tail -3 gpurun_out/r2q_bench_n2.err
timeout 900 python -X faulthandler bench.py --workload dem --steps 200 --warmup 20 > gpurun_out/r2q_bench_dem_n1.json 2> gpurun_out/r2q_bench_dem_n1.err
echo "dem n1 exit $?"; cut -c1-2500 gpurun_out/r2q_bench_dem_n1.json; tail -5 gpurun_out/r2q_bench_dem_n1.err
