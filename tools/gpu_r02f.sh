#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - t0 ))s] $*"; }
el start
timeout 200 python tools/microbench.py 3 16 6 2 > gpurun_out/r02f_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:"apply_kernel<.int.3, .int.32, .int.0, .bool.0, .bool.0, .bool.1>|interface_kernel<.int.3, .int.0, .bool.0>" -s 3 -c 6 \
    -o gpurun_out/prof_r02f python tools/microbench.py 3 16 6 2 > gpurun_out/r02f_ncu.log 2>&1
el "ncu rc=$?"; tail -5 gpurun_out/r02f_ncu.log
ls -la gpurun_out/prof_r02f* 2>/dev/null
el done
