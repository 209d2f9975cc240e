#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/restart_probe.py > gpurun_out/r2u_restart_probe.log 2>&1; cat gpurun_out/r2u_restart_probe.log | tail -8
timeout 2400 python -m pytest "tests/test_gpu_full_size.py::test_config_c3_one_million_spheres_against_the_reference" "tests/test_gpu_full_size.py::test_config_c3_settled_bed_invariants" -q --tb=short -p no:cacheprovider -s > gpurun_out/r2u_tests.log 2>&1
tail -12 gpurun_out/r2u_tests.log | cut -c1-400
