#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - t0 ))s] $*"; }
el start
timeout 600 python -m pytest tests/test_gpu_knobs.py -x -q -k "graph_replay or coarse_inverse" > gpurun_out/r02t_pytest.log 2>&1; el "pytest rc=$?"; tail -2 gpurun_out/r02t_pytest.log
for v in four u4; do HMG_LIB=$PWD/variants/libhmg_$v.so timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "tet-c2-L6 or tri-c3-L7 or tet-c2-L5" > gpurun_out/r02t_parity_$v.log 2>&1; el "parity $v rc=$?"; tail -1 gpurun_out/r02t_parity_$v.log; done
mb() { local name=$1; shift; local shape=$1; shift
  env "$@" timeout 300 python tools/microbench.py $shape 10 v > gpurun_out/r02t_mb_$name.log 2> gpurun_out/r02t_mb_$name.err; el "mb $name rc=$?"; }
mb c4_base "3 32 6" HMG_NOP=1
mb c4_four "3 32 6" HMG_LIB=$PWD/variants/libhmg_four.so
mb c4_u4 "3 32 6" HMG_LIB=$PWD/variants/libhmg_u4.so
mb c4_base2 "3 32 6" HMG_NOP=1
mb c2_base "2 192 8" HMG_NOP=1
mb c2_four "2 192 8" HMG_LIB=$PWD/variants/libhmg_four.so
mb c2_u4 "2 192 8" HMG_LIB=$PWD/variants/libhmg_u4.so
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02t_mb_*.log')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], {k:(d[k]['ms'] if isinstance(d[k],dict) else d[k]) for k in ('apply','apply_dot','residual','mul','fused_p_product','vcycle') if k in d})
    except Exception as ex:
        print(f,'unreadable',ex)
PY
el done
