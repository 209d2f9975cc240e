#!/bin/bash
# Round 2, GPU call 3W (2 GPUs): whole GPU suite on the final code; smoke(); the default bench command (N = 1, all legs)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r3w_suite.log 2>&1; echo "suite exit $?"; tail -6 gpurun_out/r3w_suite.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3w_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r3w_smoke.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/r3w_bench.json 2> gpurun_out/r3w_bench.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r3w_bench.json"))
    print(d["value"], d["ms_per_step"], d["steps"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], d["reneighbor"]["ms_per_rebuild"], d["clocks"], d["cpu_baseline"]["value"], d.get("gpu_launches"))
except Exception as e:
    print("no line", e)
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r3w_bench_n2.json 2> gpurun_out/r3w_bench_n2.err; echo "n2 exit $?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r3w_bench_n2.json"))
    print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], d["reneighbor"]["ms_per_rebuild"], d.get("parity_nranks", {}).get("ok"))
except Exception as e:
    print("no line", e)
PY
