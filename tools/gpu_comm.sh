#!/bin/bash
N=${1:-8}; TAG=${2:-r02v}; C=${3:-32}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
t0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - t0 ))s] $*"; }
el start
for mode in 1 0; do
  HMG_PEER=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) \
      tools/comm_bench.py $C 6 > gpurun_out/${TAG}_comm_n${N}_peer$mode.log 2> gpurun_out/${TAG}_comm_n${N}_peer$mode.err
  el "comm bench peer=$mode rc=$?"; tail -1 gpurun_out/${TAG}_comm_n${N}_peer$mode.log | cut -c1-600
done
el done
