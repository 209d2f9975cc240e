#!/bin/bash
# What has not run on a GPU yet, in one gpurun call (1 GPU, ~6 minutes):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/next_gpu_call.sh'
# 1. what was written after the GPU budget of round 1 was spent -- user-defined properties (csrc/props.cu), the rest of the generic
#    vocabulary, generated DEM contact models and per-particle kernels in DEM scripts: tests/test_gpu_props.py with the xfail
#    marker ignored
# 2. the whole GPU suite (the border / exchange kernels got a run-time record stride)
# 3. per-particle vs pair lists on the 4 M-atom workload (tools/bench_pair_lists.py): the experiment DESIGN.md 6b item 2 describes
# 3b. the generic path and the user-property rows next to the hand-written kernels (tools/bench_generic.py)
# 4. the headline bench line, to see that nothing moved
# Outputs land in gpurun_out/.  For 2 or 4 GPUs: gpurun --gpus 2 -- 'python -m pytest tests/test_gpu_props.py -q --runxfail -k between_ranks'
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_props.py -q --runxfail -x > gpurun_out/props_gpu.log 2>&1
echo "props exit $?" >> gpurun_out/props_gpu.log
tail -30 gpurun_out/props_gpu.log
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/gpu_suite.log 2>&1
echo "suite exit $?" >> gpurun_out/gpu_suite.log
tail -5 gpurun_out/gpu_suite.log
timeout 300 python tools/bench_pair_lists.py 100 > gpurun_out/bench_pair_lists.json 2> gpurun_out/bench_pair_lists.err
tail -c 1200 gpurun_out/bench_pair_lists.json
# one full ncu capture of the pair-list force kernel and of the per-particle one it competes with (1 M atoms keeps the replays short)
cat > /tmp/pl_prof.py <<'PY'
import sys
sys.path.insert(0, ".")
from pairs_b200 import backend
nx = 63
L = nx * pow(4.0 / 0.8442, 1.0 / 3.0)
ctx = backend.Context(0)
ctx.init_domain([0, L, 0, L, 0, L])
ctx.set_option("pair_lists", int(sys.argv[1]))
ctx.copper_fcc_lattice(nx, nx, nx, 0.8442, 4)
ctx.adjust_thermo(1.44)
ctx.set_lj_params(4, [1.0] * 16, [1.0] * 16)
ctx.md_run(0, 30, 0.005, 2.5, 2.8, 2.8, 20, 0)
PY
for on in 0 1; do
    timeout 400 ncu --set full --clock-control none --import-source on -k regex:'pb_k_lj_pairs|pb_k_lennard_jones' -s 22 -c 2 -f \
        -o gpurun_out/prof_force_pairlists_$on python /tmp/pl_prof.py $on > gpurun_out/ncu_force_pairlists_$on.log 2>&1
done
timeout 600 python tools/bench_generic.py 63 100 > gpurun_out/bench_generic.json 2> gpurun_out/bench_generic.err
tail -c 1500 gpurun_out/bench_generic.json
timeout 600 python bench.py --steps 100 --warmup 20 > gpurun_out/bench_next.json 2> gpurun_out/bench_next.err
tail -c 600 gpurun_out/bench_next.json
