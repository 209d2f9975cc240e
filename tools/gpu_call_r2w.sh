#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_gpu_checkpoint.py "tests/test_gpu_full_size.py::test_config_c3_one_million_spheres_against_the_reference" -q --tb=short -p no:cacheprovider -s > gpurun_out/r2w_tests.log 2>&1
tail -12 gpurun_out/r2w_tests.log | cut -c1-400
