#!/bin/bash
# Round 2, GPU call M (1 GPU): whole GPU suite (no -x) after the list reorder; default bench line.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2m_suite.log 2>&1
echo "suite exit $?" >> gpurun_out/r2m_suite.log
tail -25 gpurun_out/r2m_suite.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
echo "bench exit $?"; cut -c1-1500 gpurun_out/r2m_bench.json; tail -3 gpurun_out/r2m_bench.err
