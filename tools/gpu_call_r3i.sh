#!/bin/bash
# Round 2, GPU call 3I (2 GPUs): the multi-rank check of the further contact properties (checkpoint-based)
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29852 tests/scripts/mgpu_contact_props_check.py > gpurun_out/r3i_contact_props.log 2>&1; echo "exit $?"
grep -v "^DEM\|^Domain\|^Spacing\|^Diameter\|^Initial\|^Particle\|^Number" gpurun_out/r3i_contact_props.log | tail -25 | cut -c1-400
