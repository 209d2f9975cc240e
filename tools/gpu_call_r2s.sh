#!/bin/bash
# Round 2, GPU call S (1 GPU): list words through a cp.async ring: tests + bench; DEM contact kernel under register caps.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tiles.py tests/test_gpu_md.py tests/test_gpu_edge_cases.py -q --tb=short -p no:cacheprovider -x > gpurun_out/r2s_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2s_tests.log
tail -8 gpurun_out/r2s_tests.log | cut -c1-400
timeout 600 python tools/bench_tiles.py 100 > gpurun_out/r2s_bench_tiles.json 2> gpurun_out/r2s_bench_tiles.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2s_bench_tiles.json"))
for k, v in d.items():
    if isinstance(v, dict):
        print(k, round(v["ms_per_step"], 4), {a: round(b["ms_per_call"], 4) for a, b in v["stages"].items()})
PY
for R in 0 128 112 96 80; do
  PB_DEM_MAXREG=$R timeout 300 python tools/bench_dem.py 4000 > gpurun_out/r2s_dem_maxreg_$R.json 2> gpurun_out/r2s_dem_maxreg_$R.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2s_dem_maxreg_$R.json"))
    print("maxreg $R", "falling", round(d["falling"]["ms_per_step"], 4), "settled", round(d["settled"]["ms_per_step"], 4), "lsd", round(d["settled"]["stages_ms_per_step"]["linear_spring_dashpot"], 4), "contacts", round(d["settled"]["mean_contacts"], 2))
except Exception as e:
    print("maxreg $R failed", e)
PY
done
