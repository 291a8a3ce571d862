#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - t0 ))s] $*"; }
el start
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_knobs.py -x -q > gpurun_out/r02m_pytest.log 2>&1; el "pytest rc=$?"; tail -2 gpurun_out/r02m_pytest.log
mb() { local name=$1; shift; local shape=$1; shift
  env "$@" timeout 300 python tools/microbench.py $shape 10 v > gpurun_out/r02m_mb_$name.log 2> gpurun_out/r02m_mb_$name.err; el "mb $name rc=$?"; }
mb c4_td2 "3 32 6" HMG_NOP=1
mb c4_td3 "3 32 6" HMG_LIB=$PWD/variants/libhmg_td3.so
mb c4_td2_noseg "3 32 6" HMG_APPLY_SEG3_SHIFT=30
mb c2 "2 192 8" HMG_NOP=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02m_mb_*.log')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], {k:(d[k]['ms'] if isinstance(d[k],dict) else d[k]) for k in ('apply','apply_dot','residual','mul','fused_p_product','vcycle') if k in d})
    except Exception as ex:
        print(f,'unreadable',ex)
PY
el done
