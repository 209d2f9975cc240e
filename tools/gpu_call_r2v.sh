#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/restart_probe.py > gpurun_out/r2v_restart_probe.log 2>&1; cat gpurun_out/r2v_restart_probe.log | tail -8
