#!/bin/bash
# Round 2, GPU call 3F (2 GPUs): the whole GPU suite, smoke(), the three bench arms at the driver's arguments, default bench line.
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r3f_suite.log 2>&1
echo "suite exit $?" >> gpurun_out/r3f_suite.log; tail -6 gpurun_out/r3f_suite.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3f_smoke.log 2>&1; tail -2 gpurun_out/r3f_smoke.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r3f_bench_reference_arm.json 2> gpurun_out/r3f_bench_reference_arm.err
echo "ref arm exit $?"; cut -c1-400 gpurun_out/r3f_bench_reference_arm.json
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r3f_bench_k20.json 2> gpurun_out/r3f_bench_k20.err
timeout 900 python bench.py > gpurun_out/r3f_bench.json 2> gpurun_out/r3f_bench.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r3f_bench_n2_k20.json 2> gpurun_out/r3f_bench_n2_k20.err
python - <<'PY'
import json
for f in ("r3f_bench_k20", "r3f_bench", "r3f_bench_n2_k20"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, {k: d[k] for k in ("value", "ms_per_step", "steps")}, "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "parity", (d.get("parity_nranks") or {}).get("ok"), d["clocks"]["samples"])
    except Exception as e:
        print(f, "no line", e)
PY
