#!/bin/bash
# development helper: per-op timings (one B200)
out=gpurun_out/sweep_r2.log
: > $out
run() { echo "## $DIMC $*" >> $out; env "$@" timeout 300 python tools/microbench.py $DIMC 10 v >> $out 2>&1; }
DIMC="3 16 6"
run HMG_X=0
run HMG_FUSE_P=0
run HMG_APPLY_CONVERTERS=1
DIMC="2 96 8"
run HMG_APPLY_SLOT_SHIFT=1
