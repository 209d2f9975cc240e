#!/bin/bash
# Round 2, GPU call L (1 GPU): restructured tile build + conflict-aware list order: tests, bench, ncu of build and force.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tiles.py tests/test_gpu_md.py tests/test_gpu_edge_cases.py -x -q --tb=short -p no:cacheprovider > gpurun_out/r2l_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2l_tests.log
tail -15 gpurun_out/r2l_tests.log | cut -c1-300
timeout 600 python tools/bench_tiles.py 100 > gpurun_out/r2l_bench_tiles.json 2> gpurun_out/r2l_bench_tiles.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2l_bench_tiles.json"))
for k, v in d.items():
    if isinstance(v, dict):
        print(k, round(v["ms_per_step"], 4), {a: round(b["ms_per_call"], 4) for a, b in v["stages"].items()})
PY
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:pb_k_tile_lj -s 30 -c 1 -o gpurun_out/r2l_ncu_tile_lj python tools/prof_md.py 100 45 1 1 > gpurun_out/r2l_ncu_1.log 2>&1
timeout 300 $NCU -k regex:pb_k_tile_build -s 1 -c 1 -o gpurun_out/r2l_ncu_tile_build python tools/prof_md.py 100 45 1 1 > gpurun_out/r2l_ncu_2.log 2>&1
ls -la gpurun_out/r2l_*
