#!/bin/bash
# development helper: per-op timings for a few launch-shape knobs on one B200
#   tools/sweep.sh "3 20 5" HMG_APPLY_WARPS=12 HMG_APPLY_RUN=2 ...   (one run per knob setting)
# knobs (read at context creation): HMG_APPLY_WARPS, HMG_APPLY_RUN, HMG_APPLY_CHUNK_SHIFT, HMG_APPLY_SEG_SHIFT,
# HMG_APPLY_SEG3_SHIFT, HMG_APPLY_RING_ROWS, HMG_APPLY_OVERSUB, HMG_APPLY_CONVERTERS, HMG_APPLY_SLOT_SHIFT, HMG_FUSE_P,
# HMG_GRAPH, HMG_PEER,
# HMG_DEBUG_CFG=1 prints the launch shapes chosen per level
shape=$1; shift
out=gpurun_out/sweep.log
mkdir -p gpurun_out; : > $out
echo "## $shape (defaults)" >> $out; timeout 300 python tools/microbench.py $shape 10 v >> $out 2>&1
for kv in "$@"; do echo "## $shape $kv" >> $out; env $kv timeout 300 python tools/microbench.py $shape 10 v >> $out 2>&1; done
