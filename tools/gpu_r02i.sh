#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - t0 ))s] $*"; }
el start
mb() { local name=$1; shift; local shape=$1; shift
  env "$@" timeout 300 python tools/microbench.py $shape 10 v > gpurun_out/r02i_mb_$name.log 2> gpurun_out/r02i_mb_$name.err; el "mb $name rc=$?"; }
S=$PWD/variants/libhmg_serial.so
mb c4_batched "3 32 6" HMG_NOP=1
mb c4_serial "3 32 6" HMG_LIB=$S
mb l5_batched "3 32 5" HMG_NOP=1
mb l5_serial "3 32 5" HMG_LIB=$S
mb c2_batched "2 192 8" HMG_NOP=1
mb c2_serial "2 192 8" HMG_LIB=$S
mb c4_batched2 "3 32 6" HMG_NOP=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02i_mb_*.log')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], {k:(d[k]['ms'] if isinstance(d[k],dict) else d[k]) for k in ('apply','apply_dot','residual','mul','fused_p_product','cg_update','vcycle') if k in d}, 'apply frac', d['apply']['hbm_frac'])
    except Exception as ex:
        print(f,'unreadable',ex)
PY
el done
