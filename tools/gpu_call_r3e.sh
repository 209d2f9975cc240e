#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tiles.py tests/test_gpu_md.py tests/test_gpu_edge_cases.py -q --tb=short -p no:cacheprovider -x > gpurun_out/r3e_tests.log 2>&1; tail -3 gpurun_out/r3e_tests.log
timeout 600 python tools/bench_tiles.py 100 > gpurun_out/r3e_bench_tiles.json 2> gpurun_out/r3e_bench_tiles.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r3e_bench_tiles.json"))
for k, v in d.items():
    if isinstance(v, dict):
        print(k, round(v["ms_per_step"], 4), {a: round(b["ms_per_call"], 4) for a, b in v["stages"].items()})
PY
