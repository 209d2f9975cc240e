#!/bin/bash
# Round 2, GPU call R (1 GPU): ncu of the TMA-staged force / build kernels at 4 M atoms, launch list of a bench run, default bench line.
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:pb_k_tile_lj -s 30 -c 1 -o gpurun_out/r2r_ncu_tile_lj python tools/prof_md.py 100 45 1 1 > gpurun_out/r2r_ncu_1.log 2>&1
timeout 300 $NCU -k regex:pb_k_tile_build -s 1 -c 1 -o gpurun_out/r2r_ncu_tile_build python tools/prof_md.py 100 45 1 1 > gpurun_out/r2r_ncu_2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2r_launches.csv python bench.py --steps 40 --warmup 20 --no-cpu-baseline --no-dem > gpurun_out/r2r_bench_under_ncu.log 2>&1
timeout 900 python bench.py > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err
echo "bench exit $?"; python - <<'PY'
import json
d = json.load(open("gpurun_out/r2r_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"], d["roofline"]["frac"], d["roofline"]["ms_per_step_in_kernel"], d.get("reneighbor"), d["clocks"])
PY
ls -la gpurun_out/r2r_*
