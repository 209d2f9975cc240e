#!/bin/bash
# Round 2, GPU call J (2 GPUs): branch-free force kernel in the library (tests + bench), N = 2 parity module.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tiles.py tests/test_gpu_md.py -x -q --tb=short -p no:cacheprovider > gpurun_out/r2j_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2j_tests.log
tail -15 gpurun_out/r2j_tests.log | cut -c1-300
timeout 600 python tools/bench_tiles.py 100 > gpurun_out/r2j_bench_tiles.json 2> gpurun_out/r2j_bench_tiles.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2j_bench_tiles.json"))
for k, v in d.items():
    if isinstance(v, dict):
        print(k, round(v["ms_per_step"], 4), {a: round(b["ms_per_call"], 4) for a, b in v["stages"].items()})
PY
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q --tb=short -p no:cacheprovider -k "md" > gpurun_out/r2j_multi.log 2>&1
tail -5 gpurun_out/r2j_multi.log | cut -c1-1500
grep -n "rank0\]:" gpurun_out/r2j_multi.log | head -12 | cut -c1-300
