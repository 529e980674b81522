#!/bin/bash
# Compare builds of the library (build/libs/*.so) on the same pre-rolled records:
# throughput on REPS concatenated copies and single-longest-search latency.
REC=${1:-build/road_triple_speed_r32.npz}
REPS=${2:-8}
mkdir -p gpurun_out
for lib in build/libs/*.so; do
  echo "### $lib"
  PDMPC_LIB=$lib python tools/profile_batch.py $REC 3 $REPS 1 | tail -2
done
