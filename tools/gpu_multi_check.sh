#!/bin/bash
# Multi-GPU check (run through `gpurun --gpus N -- bash tools/gpu_multi_check.sh N TAG`): the partitioned-context tests
# against the oracle, then bench.py at N ranks -- peer memory (default), CUDA graphs on top, and the NCCL path.
N=${1:-2}; TAG=${2:-r02}; SEL=${3:-gpus}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - t0 ))s] $*"; }
run() { # name, env..., then bench args after --
  local name=$1; shift
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) \
      bench.py --gpus $N "$@" > gpurun_out/${TAG}_bench_n${N}_$name.log 2> gpurun_out/${TAG}_bench_n${N}_$name.err
  el "bench $name rc=$?"; tail -c 400 gpurun_out/${TAG}_bench_n${N}_$name.err | tail -3
}
el start
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 1500 python -m pytest tests/test_gpu_multi.py -q -x -k "$SEL" > gpurun_out/${TAG}_pytest_multi_n${N}.log 2>&1; el "pytest multi rc=$?"; tail -4 gpurun_out/${TAG}_pytest_multi_n${N}.log
run peer HMG_DEBUG_CFG=1 -- --steps 5 --warmup 3 --no-e2e --no-cpu-baseline
run peer_graph HMG_GRAPH=2 -- --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-also
run nccl HMG_PEER=0 -- --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-also
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${TAG}_bench_n${N}_*.log')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'value', round(d['value'],3), 'ms', round(d['ms_per_step'],3), 'comm', d['config'].get('comm'), 'ax', round(d['ax']['value'],1),
              'single', (d.get('single_gpu') or {}).get('value'), 'parity', d.get('parity_vs_single_gpu'), 'launches', d['gpu_launches'])
    except Exception as ex:
        print(f,'unreadable',ex)
PY
el done
