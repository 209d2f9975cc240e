#!/bin/bash
# Round 2, GPU call 3J (1 GPU): residue histogram moved out of the accept path of the tile build; pb_md_run_from_host
# (upload of velocities / masses overlapped with the first list build).  Tile + MD suites, then the driver's bench command.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tiles.py tests/test_gpu_md.py tests/test_gpu_edge_cases.py -q --tb=short -p no:cacheprovider > gpurun_out/r3j_suite.log 2>&1; echo "suite exit $?"; tail -12 gpurun_out/r3j_suite.log | cut -c1-300
for k in 1 2; do
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-dem --no-cpu-baseline > gpurun_out/r3j_bench_k20_$k.json 2> gpurun_out/r3j_bench_k20_$k.err; echo "bench exit $?"
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r3j_bench_k20_$k.json"))
    print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["seconds_per_repetition"], "frac", d["roofline"]["frac"], d["reneighbor"]["ms_per_rebuild"], d["clocks"])
except Exception as e:
    print("no line", e)
PY
done
timeout 300 python tools/e2e_probe.py 100 20 2>&1 | tail -1 | cut -c1-900
