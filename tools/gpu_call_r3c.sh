#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dem.py -q --tb=short -p no:cacheprovider -x > gpurun_out/r3c_tests.log 2>&1; tail -3 gpurun_out/r3c_tests.log
timeout 600 python tools/bench_dem.py 8000 > gpurun_out/r3c_bench_dem.json 2> gpurun_out/r3c_bench_dem.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r3c_bench_dem.json"))
for k in ("falling", "settled"):
    print(k, round(d[k]["ms_per_step"], 4), {a: round(b, 4) for a, b in d[k]["stages_ms_per_step"].items() if b > 0}, "contacts", round(d[k]["mean_contacts"], 2))
PY
