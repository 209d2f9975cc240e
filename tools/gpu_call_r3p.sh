#!/bin/bash
# Round 2, GPU call 3P (8 GPUs): weak scaling on the final code, N = 8 at the driver's 20 steps and at 100 steps (N-rank parity inside)
mkdir -p gpurun_out
for K in 20 100; do
W=5; [ $K = 100 ] && W=20
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2957$((K/20)) bench.py --gpus 8 --steps $K --warmup $W > gpurun_out/r3p_bench_n8_k$K.json 2> gpurun_out/r3p_bench_n8_k$K.err; echo "n8 k$K exit $?"
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r3p_bench_n8_k$K.json"))
    print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], d["reneighbor"], d.get("parity_nranks", {}).get("ok"), d["clocks"])
except Exception as e:
    print("no line", e)
PY
tail -3 gpurun_out/r3p_bench_n8_k$K.err | cut -c1-300
done
