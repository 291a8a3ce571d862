#!/bin/bash
# development helper: launch-shape sweep of the 3D apply kernel (one B200)
out=gpurun_out/sweep_r2.log
: > $out
run() { echo "## $DIMC $*" >> $out; env "$@" timeout 300 python tools/microbench.py $DIMC 10 v >> $out 2>&1; }
DIMC="3 20 5"
run HMG_APPLY_RUN=2 HMG_APPLY_WARPS=12
run HMG_APPLY_RUN=2 HMG_APPLY_WARPS=13
run HMG_APPLY_RUN=2 HMG_APPLY_WARPS=14
run HMG_APPLY_RUN=1 HMG_APPLY_WARPS=12
DIMC="3 16 6"
run HMG_APPLY_RUN=2 HMG_APPLY_CHUNK_SHIFT=5
run HMG_APPLY_RUN=1 HMG_APPLY_CHUNK_SHIFT=6 HMG_APPLY_WARPS=12
run HMG_APPLY_RUN=1 HMG_APPLY_CHUNK_SHIFT=6 HMG_APPLY_WARPS=14
DIMC="2 96 8"
run HMG_APPLY_WARPS=12
run HMG_APPLY_SEG_SHIFT=6
