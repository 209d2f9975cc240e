#!/bin/bash
# Round 2, GPU call T (1 GPU): checkpoint / restart, contact-capacity growth, C3 at size.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_gpu_checkpoint.py tests/test_gpu_dem.py "tests/test_gpu_full_size.py::test_config_c3_one_million_spheres_against_the_reference" "tests/test_gpu_full_size.py::test_config_c3_settled_bed_invariants" -q --tb=short -p no:cacheprovider > gpurun_out/r2t_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2t_tests.log
tail -40 gpurun_out/r2t_tests.log | cut -c1-500
