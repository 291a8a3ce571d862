#!/bin/bash
out=gpurun_out/sweep_r2.log
: > $out
run() { echo "## $DIMC $*" >> $out; env "$@" timeout 300 python tools/microbench.py $DIMC 10 v >> $out 2>&1; }
DIMC="3 16 6"
run HMG_X=0
DIMC="3 20 5"
run HMG_X=0
