#!/bin/bash
# End-of-round evidence on one B200: the full GPU suite, bench.py as the driver runs it, the reference arm, the C5 sweep,
# and the ncu launch list of the bench command (library kernels only).  TAG names the files under gpurun_out/.
TAG=${1:-r02k}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - t0 ))s] $*"; }
el start
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; el "pytest gpu rc=$?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench_default.log 2> gpurun_out/${TAG}_bench_default.err; el "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_reference.log 2>&1; el "bench reference rc=$?"
timeout 900 python tools/sweep_c5.py > gpurun_out/${TAG}_c5_sweep.jsonl 2> gpurun_out/${TAG}_c5_sweep.err; el "c5 sweep rc=$?"
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-also"
timeout 600 $B > gpurun_out/${TAG}_plain.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:"hmg::" -c 8000 --csv \
    --log-file gpurun_out/${TAG}_launches_C4.csv $B > gpurun_out/${TAG}_ncu.log 2>&1
el "launch list rc=$?"
python - <<PY
import json
for f in ('gpurun_out/${TAG}_bench_default.log','gpurun_out/${TAG}_bench_reference.log'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], {k:d.get(k) for k in ('value','ms_per_step','gpu_launches')}, 'ax', d.get('ax'), 'roofline frac', (d.get('roofline') or {}).get('frac'), 'e2e', (d.get('e2e') or {}).get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
        if d.get('also'): print('  also', {k:(v['value'], v['ms_per_step'], v['ax']['value'], v['roofline']['frac']) for k,v in d['also'].items()})
    except Exception as ex:
        print(f,'unreadable',ex)
PY
el done
