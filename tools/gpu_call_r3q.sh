#!/bin/bash
# Round 2, GPU call 3Q (4 GPUs): weak scaling on the final code, N = 4 at the driver's 20 steps
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r3q_bench_n4_k20.json 2> gpurun_out/r3q_bench_n4_k20.err; echo "n4 exit $?"
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r3q_bench_n4_k20.json"))
    print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], d["reneighbor"]["ms_per_rebuild"], d.get("parity_nranks", {}).get("ok"))
except Exception as e:
    print("no line", e)
PY
