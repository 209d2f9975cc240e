#!/bin/bash
# Round 2, GPU call 3A (2 GPUs): DEM contact kernel with CTA-local sort by partner count: DEM suites (incl. generated models, C3, 2-GPU), DEM bench.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_gpu_dem.py tests/test_gpu_props.py tests/test_gpu_checkpoint.py tests/test_gpu_examples.py "tests/test_gpu_full_size.py::test_config_c3_one_million_spheres_against_the_reference" "tests/test_gpu_full_size.py::test_config_c3_settled_bed_invariants" "tests/test_gpu_multi.py::test_multi_gpu_dem_matches_single_gpu" -q --tb=short -p no:cacheprovider > gpurun_out/r3a_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r3a_tests.log
tail -25 gpurun_out/r3a_tests.log | cut -c1-400
timeout 600 python tools/bench_dem.py 8000 > gpurun_out/r3a_bench_dem.json 2> gpurun_out/r3a_bench_dem.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r3a_bench_dem.json"))
for k in ("falling", "settled"):
    print(k, round(d[k]["ms_per_step"], 4), {a: round(b, 4) for a, b in d[k]["stages_ms_per_step"].items() if b > 0}, "contacts", round(d[k]["mean_contacts"], 2))
PY
