#!/bin/bash
# development helper: per-op timings (one B200)
out=gpurun_out/sweep_r2.log
: > $out
run() { echo "## $DIMC $*" >> $out; env "$@" timeout 300 python tools/microbench.py $DIMC 10 >> $out 2>&1; }
DIMC="3 16 6"
run HMG_GROUP_WIDTH=32
run HMG_GROUP_WIDTH=16
run HMG_GROUP_WIDTH=16 HMG_APPLY_RING_ROWS=1200
run HMG_GROUP_WIDTH=16 HMG_APPLY_RUN=2
DIMC="3 20 5"
run HMG_GROUP_WIDTH=16
