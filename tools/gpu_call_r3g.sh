#!/bin/bash
# Round 2, GPU call 3G (2 GPUs): DEM bench at N = 2 with the N-rank parity block; DEM multi test; clock sampler at 20 steps.
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --workload dem --gpus 2 --steps 200 --warmup 20 > gpurun_out/r3g_bench_dem_n2.json 2> gpurun_out/r3g_bench_dem_n2.err
echo "dem n2 exit $?"; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r3g_bench_dem_n2.json"))
    print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["roofline"]["frac"], d["falling_phase"], json.dumps(d.get("parity_nranks"))[:600])
except Exception as e:
    print("no line", e)
PY
tail -4 gpurun_out/r3g_bench_dem_n2.err | cut -c1-300
timeout 600 python -m pytest "tests/test_gpu_multi.py::test_multi_gpu_dem_matches_single_gpu" -q --tb=short -p no:cacheprovider > gpurun_out/r3g_multi.log 2>&1; tail -3 gpurun_out/r3g_multi.log | cut -c1-300
for k in 1 2; do timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-dem --no-cpu-baseline > gpurun_out/r3g_bench_k20_$k.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r3g_bench_k20_$k.json')); print(d['value'], d['e2e']['value'], d['clocks'])"; done
