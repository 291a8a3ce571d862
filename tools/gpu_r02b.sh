#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - t0 ))s] $*"; }
el start
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest_gpu.log 2>&1; el "pytest gpu rc=$?"; tail -5 gpurun_out/r02b_pytest_gpu.log
timeout 300 python tools/microbench.py 3 32 6 10 v > gpurun_out/r02b_mb3d.log 2>&1; el "mb3d rc=$?"
timeout 300 python tools/microbench.py 2 192 8 10 v > gpurun_out/r02b_mb2d.log 2>&1; el "mb2d rc=$?"
# compute-sanitizer on the mbarrier ring / converter / interface kernels (small cases)
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "(tet-c2-L5 or tri-c3-L7) and (global_product or smoothing or interface or residual_history)" > gpurun_out/r02b_sanitizer_$tool.log 2>&1; el "sanitizer $tool rc=$?"; tail -4 gpurun_out/r02b_sanitizer_$tool.log
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02b_mb*.log')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], {k:(d[k]['ms'] if isinstance(d[k],dict) else d[k]) for k in ('apply','apply_dot','interface','interface_pairs','interface_multi','global_product','residual','fused_p_product','cg_update','restrict','interp','vcycle') if k in d})
    except Exception as ex:
        print(f,'unreadable',ex)
PY
el done
