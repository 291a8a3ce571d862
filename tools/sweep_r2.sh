#!/bin/bash
# development helper: apply-kernel launch-shape sweep (one B200)
out=gpurun_out/sweep_r2.log
: > $out
run() { echo "## $DIMC $*" >> $out; env "$@" timeout 200 python tools/microbench.py $DIMC 10 >> $out 2>&1; }
DIMC="3 20 5"
run HMG_X=0
run HMG_APPLY_RUN=2
run HMG_APPLY_RUN=1
run HMG_APPLY_CHUNK_SHIFT=6
run HMG_APPLY_WARPS=12
DIMC="2 64 8"
run HMG_X=0
run HMG_APPLY_WARPS=8
run HMG_APPLY_SEG_SHIFT=6
run HMG_APPLY_SEG_SHIFT=4
DIMC="3 8 6"
run HMG_X=0
run HMG_APPLY_CHUNK_SHIFT=6
