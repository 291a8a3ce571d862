#!/bin/bash
# GPU check used at the end of round 1: interface kernel split (pairs / multi-owner cells), the multi-owner variant,
# an L2-resident problem, and the full GPU suite.  Run through gpurun from the repository root.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - t0 ))s] $*"; }
el start
timeout 90 python tools/microbench.py 3 16 6 20 v > gpurun_out/r01s_mb3d_default.log 2>&1; el "mb3d default rc=$?"
HMG_IFACE_MULTI=1 timeout 45 python tools/microbench.py 3 16 6 20 v > gpurun_out/r01s_mb3d_multi1.log 2>&1; el "mb3d multi1 rc=$?"
HMG_IFACE_MULTI=1 timeout 60 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r01s_parity_multi1.log 2>&1; el "parity multi1 rc=$?"; tail -2 gpurun_out/r01s_parity_multi1.log
# face nodes enumerated along lattice lines (HMG_FACE_ORDER=1, untimed so far): parity, then the same microbench
HMG_FACE_ORDER=1 timeout 60 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_tests.py -x -q > gpurun_out/r01s_parity_faceorder1.log 2>&1; el "parity faceorder1 rc=$?"; tail -2 gpurun_out/r01s_parity_faceorder1.log
HMG_FACE_ORDER=1 timeout 60 python tools/microbench.py 3 16 6 20 v > gpurun_out/r01s_mb3d_faceorder1.log 2>&1; el "mb3d faceorder1 rc=$?"
# multi-owner cells as a second kernel on a second stream (HMG_IFACE_SPLIT=1, untimed so far)
HMG_IFACE_SPLIT=1 timeout 60 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r01s_parity_split1.log 2>&1; el "parity split1 rc=$?"; tail -2 gpurun_out/r01s_parity_split1.log
HMG_IFACE_SPLIT=1 timeout 60 python tools/microbench.py 3 16 6 20 v > gpurun_out/r01s_mb3d_split1.log 2>&1; el "mb3d split1 rc=$?"
timeout 45 python tools/microbench.py 3 6 6 20 > gpurun_out/r01s_mb3d_c6_L2resident.log 2>&1; el "mb3d c6 rc=$?"
timeout 100 python -m pytest tests -m gpu -x -q > gpurun_out/r01s_pytest_gpu.log 2>&1; el "pytest gpu rc=$?"; tail -3 gpurun_out/r01s_pytest_gpu.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r01s_mb*.log')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k:(d[k]['ms'] if isinstance(d[k],dict) else d[k]) for k in ('apply','interface','interface_pairs','interface_multi','global_product','vcycle') if k in d})
    except Exception as ex:
        print(f,'unreadable',ex)
PY
el done
