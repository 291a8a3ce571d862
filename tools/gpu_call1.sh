#!/bin/bash
# one-shot GPU check of the round's last additions (rows N2-N4, interface-kernel variants)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - t0 ))s] $*"; }
el start
timeout 110 python -m pytest tests/test_gpu_next_rows.py -x -q > gpurun_out/r01r_next_rows.log 2>&1; el "next_rows rc=$?"; tail -5 gpurun_out/r01r_next_rows.log
HMG_IFACE_VARIANT=1 timeout 70 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r01r_parity_v1.log 2>&1; el "parity v1 rc=$?"; tail -3 gpurun_out/r01r_parity_v1.log
for v in 0 1 2; do
  HMG_IFACE_VARIANT=$v timeout 45 python tools/microbench.py 3 16 6 20 v > gpurun_out/r01r_mb3d_v$v.log 2>&1; el "mb3d v$v rc=$?"
done
for v in 0 1; do
  HMG_IFACE_VARIANT=$v timeout 45 python tools/microbench.py 2 96 8 20 v > gpurun_out/r01r_mb2d_v$v.log 2>&1; el "mb2d v$v rc=$?"
done
HMG_IFACE_VARIANT=2 timeout 60 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r01r_parity_v2.log 2>&1; el "parity v2 rc=$?"; tail -3 gpurun_out/r01r_parity_v2.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r01r_mb*.log')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k:(d[k]['ms'] if isinstance(d[k],dict) else d[k]) for k in ('apply','interface','global_product','residual','vcycle') if k in d})
    except Exception as ex:
        print(f,'unreadable',ex)
PY
el done
