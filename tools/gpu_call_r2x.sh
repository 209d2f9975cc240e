#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_dem.py -q --tb=short -p no:cacheprovider -s -k "corrected or grows" > gpurun_out/r2x_tests.log 2>&1
grep -n "corrected oracle\|passed\|failed\|Error" gpurun_out/r2x_tests.log | cut -c1-300
