#!/bin/bash
# Round 2, GPU call 3R (1 GPU): compute-sanitizer memcheck over the paths added late in the round (small cases)
mkdir -p gpurun_out
for w in md dem; do
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 5 python tools/sanitize_probe.py $w > gpurun_out/r3r_memcheck_$w.log 2>&1; echo "memcheck $w exit $?"
grep -E "ERROR SUMMARY|probe ok|Invalid|out of bounds" gpurun_out/r3r_memcheck_$w.log | head -6 | cut -c1-200
done
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 5 python tools/sanitize_probe.py md > gpurun_out/r3r_racecheck_md.log 2>&1; echo "racecheck md exit $?"
grep -E "RACECHECK SUMMARY|probe ok|hazard" gpurun_out/r3r_racecheck_md.log | head -6 | cut -c1-200
