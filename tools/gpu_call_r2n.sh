#!/bin/bash
# Round 2, GPU call N (1 GPU): L1 prefetch of list words; where the e2e leg spends its time.
mkdir -p gpurun_out
timeout 600 python tools/bench_tiles.py 100 > gpurun_out/r2n_bench_tiles.json 2> gpurun_out/r2n_bench_tiles.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2n_bench_tiles.json"))
for k, v in d.items():
    if isinstance(v, dict):
        print(k, round(v["ms_per_step"], 4), {a: round(b["ms_per_call"], 4) for a, b in v["stages"].items()})
PY
timeout 600 python tools/e2e_probe.py 100 300 > gpurun_out/r2n_e2e_probe.json 2>&1; cat gpurun_out/r2n_e2e_probe.json
timeout 600 python tools/e2e_probe.py 100 100 > gpurun_out/r2n_e2e_probe100.json 2>&1; cat gpurun_out/r2n_e2e_probe100.json
