#!/bin/bash
# Round 2, GPU call Y (8 GPUs): LJ bench at N = 8 and N = 4 with the N-rank parity block; multi-GPU tests at 4 and 8.
mkdir -p gpurun_out
for N in 8 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520 + N)) bench.py --gpus $N --steps 100 --warmup 20 > gpurun_out/r2y_bench_n$N.json 2> gpurun_out/r2y_bench_n$N.err
  echo "lj n$N exit $?"; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2y_bench_n$N.json"))
    print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], (d.get("parity_nranks") or {}).get("ok"), (d.get("parity_nranks") or {}).get("seconds"), d.get("reneighbor"))
    print(d["stages_ms"])
except Exception as e:
    print("no line", e)
PY
  tail -3 gpurun_out/r2y_bench_n$N.err
done
timeout 900 python -m pytest tests/test_gpu_multi.py -q --tb=short -p no:cacheprovider -k "4 or 8" > gpurun_out/r2y_multi.log 2>&1
tail -5 gpurun_out/r2y_multi.log | cut -c1-600
