#!/bin/bash
# Round 2, GPU call 3H (2 GPUs): whole GPU suite after the contact-table generalisation (further contact properties in extra lanes);
# DEM bench at N = 1 to confirm the library's contact kernel did not move.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r3h_suite.log 2>&1; echo "suite exit $?"; tail -15 gpurun_out/r3h_suite.log | cut -c1-400
timeout 600 python bench.py --workload dem --gpus 1 --steps 200 --warmup 20 > gpurun_out/r3h_bench_dem.json 2> gpurun_out/r3h_bench_dem.err; echo "dem exit $?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r3h_bench_dem.json"))
    print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["roofline"], d.get("falling_phase"))
except Exception as e:
    print("no line", e)
PY
