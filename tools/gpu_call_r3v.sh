#!/bin/bash
# Round 2, GPU call 3V (1 GPU): list rows of a tile sorted by length (force-kernel warps of equal iteration count)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_tiles.py tests/test_gpu_md.py tests/test_gpu_edge_cases.py -q --tb=short -p no:cacheprovider > gpurun_out/r3v_suite.log 2>&1; echo "suite exit $?"; tail -12 gpurun_out/r3v_suite.log | cut -c1-300
for k in 1 2; do
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-dem --no-cpu-baseline > gpurun_out/r3v_bench_k20_$k.json 2> gpurun_out/r3v_bench_k20_$k.err; echo "bench exit $?"
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r3v_bench_k20_$k.json"))
    print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], d["reneighbor"]["ms_per_rebuild"], d["clocks"])
except Exception as e:
    print("no line", e)
PY
done
timeout 300 python tools/e2e_probe.py 100 20 2>&1 | tail -1 | cut -c1-900
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pb_k_tile_lj -s 3 -c 1 -o gpurun_out/r3v_ncu_tile_lj -f python bench.py --gpus 1 --steps 2 --warmup 1 --no-dem --no-cpu-baseline > gpurun_out/r3v_ncu.log 2>&1; echo "ncu exit $?"
