#!/bin/bash
# Round 2, GPU call 3D (1 GPU): persistent transfer staging: e2e probe at the driver's K = 20, MD tests, bench at the driver's arguments.
mkdir -p gpurun_out
timeout 600 python tools/e2e_probe.py 100 20 > gpurun_out/r3d_e2e_probe20.json 2>&1; cat gpurun_out/r3d_e2e_probe20.json | cut -c1-700
timeout 600 python tools/e2e_probe.py 100 20 > gpurun_out/r3d_e2e_probe20b.json 2>&1; cat gpurun_out/r3d_e2e_probe20b.json | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_md.py tests/test_cabi_and_host.py -q --tb=short -p no:cacheprovider -x > gpurun_out/r3d_tests.log 2>&1; tail -3 gpurun_out/r3d_tests.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r3d_bench_k20.json 2> gpurun_out/r3d_bench_k20.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r3d_bench_k20.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"], d["roofline"]["frac"], d["clocks"], d["cpu_baseline"])
PY
