#!/usr/bin/env python
"""Per-call cost of the two multi-GPU communication patterns of a V-cycle, device-timed (hmg_time_op, max over ranks):
op 17 = a dot product summed over the ranks on level 1 (a few entries per rank: pure latency of the scalar sum),
op 18 = the cut-cell exchange alone (pack, transfer, unpack) on every level.  Launch with torchrun, one rank per GPU:

    python -m torch.distributed.run --nproc-per-node N tools/comm_bench.py [cells levels]      (HMG_PEER=0: the NCCL path)
"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import hmgb200 as hmg


def main():
    c = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    levels = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mesh, sigma = hmg.inputs.checkerboard_problem(3, c)
    owner = hmg.inputs.spatial_partition(mesh, world)
    idbuf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        raw = (ctypes.c_ubyte * 128)()
        hmg._lib.check(hmg.load().hmg_nccl_unique_id(raw))
        idbuf = torch.tensor(list(raw), dtype=torch.uint8)
    idbuf = idbuf.cuda()
    dist.broadcast(idbuf, 0)
    g = hmg.ImplicitFineGrid(mesh, levels, sigma, lam=1.0, device=local, owner_rank=owner, rank=rank, nranks=world,
                             nccl_id=bytes(idbuf.cpu().tolist()))
    for l in range(1, levels + 1):
        st = g.state(l)
        st.p.fill(1.0)
        st.Ap.fill(1.0)

    def timed(op, level, reps):
        g.time_op(op, level, 0, 5)
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([g.time_op(op, level, 0, reps) / reps], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) * 1e3          # microseconds per call

    out = {"ranks": world, "comm": g.comm_mode(), "cells": c, "grids": levels,
           "scalar_sum_us": timed(17, 1, 200),
           "cut_exchange_us": {str(l): timed(18, l, 100) for l in range(1, levels + 1)}}
    if rank == 0:
        print(json.dumps(out))
    g.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
