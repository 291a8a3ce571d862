#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - t0 ))s] $*"; }
el start
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_pytest_gpu.log 2>&1; el "pytest gpu rc=$?"; tail -5 gpurun_out/r02c_pytest_gpu.log
timeout 300 python tools/microbench.py 3 32 6 10 v > gpurun_out/r02c_mb3d.log 2>&1; el "mb3d rc=$?"
HMG_CG_PAIRS=0 timeout 300 python tools/microbench.py 3 32 6 10 v > gpurun_out/r02c_mb3d_nopairs.log 2>&1; el "mb3d nopairs rc=$?"
timeout 300 python tools/microbench.py 2 192 8 10 v > gpurun_out/r02c_mb2d.log 2>&1; el "mb2d rc=$?"
HMG_CG_PAIRS=0 timeout 300 python tools/microbench.py 2 192 8 10 v > gpurun_out/r02c_mb2d_nopairs.log 2>&1; el "mb2d nopairs rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02c_mb*.log')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], {k:(d[k]['ms'] if isinstance(d[k],dict) else d[k]) for k in ('apply','interface','interface_pairs','interface_multi','global_product','cg_update','cg_update_pairs','x_update','vcycle') if k in d})
    except Exception as ex:
        print(f,'unreadable',ex)
PY
el done
