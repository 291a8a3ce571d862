#!/usr/bin/env python
"""bench.py -- fine-grid GDOF/s of the hot path (one step = one multigrid V-cycle, plus the
matrix-free global product A*x) on synthetic checkerboard inputs, with the HBM roofline of the
dominant kernel and the CPU baseline timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload auto|C1..C4] [--cells c]

One JSON line is printed by rank 0.  See DESIGN.md ("Measurement") for the byte model.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

# dim, cells per side, grids (= refinements + 1); BASELINE.md section 4
WORKLOADS = {
    "C1": dict(dim=2, c=48, levels=5, name="C1: 2D Tri64 checkerboard c=48 refinements=4 (README example shape)"),
    # c = 256: BASELINE.md section 4 (65 025 interior base nodes: the dense coarse inverse takes 34 GB and the 64-bit
    # cuSOLVER path; "coarse_solver_setup_s" in the output is that factorisation)
    "C2": dict(dim=2, c=256, levels=8, name="C2: 2D Tri64 checkerboard c=256 refinements=7"),
    "C3": dict(dim=3, c=20, levels=5, name="C3: 3D Tet64 checkerboard c=20 refinements=4"),
    "C4": dict(dim=3, c=32, levels=6, name="C4: 3D Tet64 checkerboard c=32 refinements=5"),
}
SMOOTHING_STEPS = 3       # top level; 2 below (reference behaviour, src/multigrid.jl:109)


def nf_of(dim, level):
    m = 1 << (level - 1)
    return (m + 1) * (m + 2) // 2 if dim == 2 else (m + 1) * (m + 2) * (m + 3) // 6


def vcycle_bytes_per_dof(dim, levels, s_top=SMOOTHING_STEPS, s_inner=2):
    """Compulsory-traffic model of one V-cycle, bytes per finest stored DOF:
    B_V = 8 (20 s_L + 7) + 8 sum_{k=2}^{L-1} rho_k (20 s_k + 8)."""
    nfL = nf_of(dim, levels)
    b = 8.0 * (20 * s_top + 7)
    for k in range(2, levels):
        b += 8.0 * nf_of(dim, k) / nfL * (20 * s_inner + 8)
    return b


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                self.samples.append([f.strip() for f in out.stdout.strip().split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = max(mx, float(s[1]))
                for n, v in zip(names, s[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def build_inputs(w, rank=0, nranks=1):
    import hmgb200 as hmg
    mesh, sigma = hmg.inputs.checkerboard_problem(w["dim"], w["c"], field=w.get("field", "checkerboard"), seed=1)
    return mesh, sigma


# ------------------------------------------------------------------------------------------------
# CPU arm: the threaded restatement of the reference on a bounded sample of the same workload
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(w, steps, warmup, target_seconds=45.0):
    """Times V-cycles (and the global product A*x) of the CPU restatement on a sample: same
    dim / grids / operator, fewer base cells, sized for about `target_seconds` of CPU work."""
    import hmgb200 as hmg
    from oracle.mesh import Mesh as OMesh
    from oracle.fem import build_local_diffusion_operators, build_local_mass_matrices, assemble_matrix
    from oracle.interfaces import list_boundary_nodes_edges_faces, list_interior_nodes
    from oracle.implicit import ImplicitFineGrid, ZeroDirichletConstraint, broadcast_interfaces, apply_constraint, local_rhs
    from oracle.operators import L2PlusDivAGrad
    from oracle.multigrid import LevelState, BaseLevel
    from oracle.cref import CpuReference

    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        cores = os.cpu_count() or 1
    dim, levels = w["dim"], w["levels"]
    nf = nf_of(dim, levels)
    # cost model of the restatement, measured on the GPU boxes' 16 host cores in round 2: 1.0 s per V-cycle of 1.35e7
    # stored DOFs in 3D, i.e. ~0.7e-6 effective core-seconds per stored DOF per V-cycle (the interface sums are serial,
    # so it scales sub-linearly with the core count); 2D is cheaper
    dof_budget = target_seconds / max(1, steps + warmup) / 0.8e-6 * min(cores, 64) ** 0.8
    per_cell = nf * (2 if dim == 2 else 6)
    c = int(max(2, min(w["c"], round((dof_budget / per_cell) ** (1.0 / dim)))))
    mesh, sigma = hmg.inputs.checkerboard_problem(dim, c, seed=1)
    base = OMesh(mesh.nodes, mesh.elements)
    imp = ImplicitFineGrid(base, levels)
    z = ZeroDirichletConstraint(*list_boundary_nodes_edges_faces(base))
    ops = [L2PlusDivAGrad(d, m, z, 1.0, sigma) for d, m in
           zip(build_local_diffusion_operators(imp.reference), build_local_mass_matrices(imp.reference))]
    ref = CpuReference(imp, ops, nthreads=cores)
    states = [LevelState(imp, l) for l in range(1, levels + 1)]
    top = states[-1]
    rng = np.random.default_rng(7)
    top.x[:, :] = rng.random(top.x.shape)
    broadcast_interfaces(top.x, imp, levels)
    apply_constraint(top.x, levels, z, imp)
    local_rhs(top.b, imp)
    interior = list_interior_nodes(base)
    A = assemble_matrix(base, sigma=sigma, lam=1.0)[interior][:, interior]
    bl = BaseLevel(A, base.nnodes, interior)
    dofs = nf * base.nelements
    for _ in range(warmup):
        ref.vcycle(bl, states, levels, SMOOTHING_STEPS)
    t0 = time.perf_counter()
    for _ in range(steps):
        ref.vcycle(bl, states, levels, SMOOTHING_STEPS)
    t_v = (time.perf_counter() - t0) / steps
    top.p[:, :] = top.x
    ref.global_product(top, levels)
    t0 = time.perf_counter()
    reps = max(1, min(10, int(2.0 / max(t_v / 10, 1e-3))))
    for _ in range(reps):
        ref.global_product(top, levels)
    t_ax = (time.perf_counter() - t0) / reps
    return dict(vcycle_gdofs=dofs / t_v / 1e9, ax_gdofs=dofs / t_ax / 1e9, ms_per_step=t_v * 1e3, cores=cores,
                sample=f"{w['name'].split(':')[0]} shape with c={c} cells per side ({base.nelements} coarse elements, "
                       f"{dofs} stored DOFs), {steps} V-cycles after {warmup} warm-up",
                dofs=dofs)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def host_memory_available():
    try:
        import psutil
        return int(psutil.virtual_memory().available)
    except Exception:
        return 1 << 40


def traffic_per_launch(name, dofs_local):
    """dram__bytes_read.sum + dram__bytes_write.sum of the finest-level apply kernel from the committed
    `ncu --set full` capture (profiles/apply_traffic.json: measured bytes per stored DOF per launch)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "apply_traffic.json")))
        return float(t[name]["dram_bytes_per_dof"]) * dofs_local, t[name].get("source")
    except Exception:
        return None, None


def measure(w, args, torch, hmg, dist, rank, world, device, steps, warmup, with_e2e, with_clocks, with_history=False):
    """One workload on this process' GPU (its share of the coarse elements when world > 1)."""
    dim, levels = w["dim"], w["levels"]
    mesh, sigma = build_inputs(w)
    if world > 1:
        owner = hmg.inputs.spatial_partition(mesh, world)
        idbuf = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            raw = (__import__("ctypes").c_ubyte * 128)()
            hmg._lib.check(hmg.load().hmg_nccl_unique_id(raw))
            idbuf = torch.tensor(list(raw), dtype=torch.uint8)
        idbuf = idbuf.cuda()
        dist.broadcast(idbuf, 0)
        nccl_id = bytes(idbuf.cpu().tolist())
        g = hmg.ImplicitFineGrid(mesh, levels, sigma, lam=1.0, device=device, owner_rank=owner, rank=rank,
                                 nranks=world, nccl_id=nccl_id)
    else:
        g = hmg.ImplicitFineGrid(mesh, levels, sigma, lam=1.0, device=device)
    nf = g.nf(levels)
    ne_local = g.ne_local
    dofs_total = nf * mesh.nelements
    dofs_local = nf * ne_local

    # inputs: x0 ~ U(0,1) then interface-sum + zero Dirichlet; b = local functional of 1 (un-summed).  The value of an
    # entry depends on the GLOBAL element index only (one Philox stream per block of 4096 elements), so every
    # partition of the mesh starts from the same vector -- which is what "parity_vs_single_gpu" compares.
    st = g.state(levels)
    l2g = g.local_elements() if world > 1 else np.arange(ne_local, dtype=np.int64)
    hx = torch.empty((ne_local, nf), dtype=torch.float64, pin_memory=True)      # column-major Nf x Ne
    BLK = 4096
    blocks = l2g // BLK
    starts = np.flatnonzero(np.r_[True, blocks[1:] != blocks[:-1]])
    ends = np.r_[starts[1:], len(l2g)]
    for a, z in zip(starts, ends):
        gb = int(blocks[a])
        n_in_block = min(BLK, mesh.nelements - gb * BLK)
        blk = np.random.Generator(np.random.Philox(key=1234, counter=[0, 0, 0, gb])).random((n_in_block, nf))
        hx[a:z] = torch.from_numpy(blk[l2g[a:z] - gb * BLK])
    X = hx.numpy().T
    st.x.set(X)
    bval = 1.0 / (nf * (2 if dim == 2 else 6))
    st.b.fill(bval)
    hmg.broadcast_interfaces(st.x, g, levels)
    hmg.apply_constraint(st.x, levels, g)
    st.p.copy_from(st.x)
    t_setup = time.perf_counter()
    bl = hmg.BaseLevel(g)
    t_setup = time.perf_counter() - t_setup

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        g.synchronize()

    def max_over_ranks(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # --- residual history of the first V-cycles (the parity observable; compared across partitions at N > 1) ----
    history = None
    if with_history:
        history = [float(v) for v in hmg.vcycles(g, bl, levels, SMOOTHING_STEPS, 3)]
        st.x.set(X)
        hmg.broadcast_interfaces(st.x, g, levels)
        hmg.apply_constraint(st.x, levels, g)

    # --- V-cycle, inputs resident in HBM -------------------------------------------------------
    for _ in range(warmup):
        hmg.vcycle(g, bl, levels, SMOOTHING_STEPS)
    barrier()
    launches_w = g.launch_count()
    sampler = ClockSampler(device) if with_clocks else None
    if sampler:
        sampler.start()
    barrier()
    ms = g.time_op(1, levels, SMOOTHING_STEPS, steps)         # CUDA events on the library's stream
    barrier()
    ms = max_over_ranks(ms)
    launches = g.launch_count() - launches_w
    vcycle_gdofs = dofs_total * steps / (ms * 1e-3) / 1e9

    # --- A*x (global product), and the dominant kernel alone ------------------------------------
    reps = 20
    g.time_op(0, levels, 0, 3)
    barrier()
    ms_ax = max_over_ranks(g.time_op(0, levels, 0, reps)) / reps
    g.time_op(3, levels, 0, 3)
    barrier()
    ms_apply = max_over_ranks(g.time_op(3, levels, 0, reps)) / reps
    clocks = sampler.summary() if sampler else None
    peak, peak_src = measured_peak()
    apply_bytes = 16.0 * dofs_local                            # read p once, write Ap once
    achieved = apply_bytes / (ms_apply * 1e-3) / 1e9
    traffic, traffic_src = traffic_per_launch(w["key"], dofs_local)

    # --- end to end through the public API with host buffers ------------------------------------
    e2e = None
    if with_e2e:
        hb = torch.empty((ne_local, nf), dtype=torch.float64, pin_memory=True)
        hout = torch.empty((ne_local, nf), dtype=torch.float64, pin_memory=True)
        hb.fill_(bval)
        B, OUT = hb.numpy().T, hout.numpy().T

        def e2e_step():
            st.x.set(X)
            st.b.set(B)
            hmg.vcycle(g, bl, levels, SMOOTHING_STEPS)
            st.x.get(OUT)
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, min(steps, 3))
        for _ in range(n_e2e):
            e2e_step()
        barrier()
        t_e2e = max_over_ranks((time.perf_counter() - t0) / n_e2e)
        e2e = {"value": dofs_total / t_e2e / 1e9, "unit": "GDOF/s", "h2d_bytes_per_step": 2 * 8 * dofs_local,
               "d2h_bytes_per_step": 8 * dofs_local, "steps": n_e2e,
               "what": "upload x and b from pinned host memory through hmg_upload, one V-cycle, download x"}
        del hb, hout
    bv = vcycle_bytes_per_dof(dim, levels)
    out = {
        "value": vcycle_gdofs, "ms_per_step": ms / steps,
        "config": {"workload": w["name"], "dim": dim, "grids": levels, "coarse_elements": int(mesh.nelements),
                   "stored_dofs": int(dofs_total), "smoothing_steps": SMOOTHING_STEPS,
                   "field": "checkerboard sigma in {1,9} per axis, seed 1", "lambda": 1.0,
                   "l2": f"inputs larger than L2 ({8 * dofs_local / 1e6:.0f} MB per vector per GPU vs 126 MB)",
                   "partition": "spatial blocks of whole cells, strong scaling" if world > 1 else "single GPU",
                   "comm": g.comm_mode(),
                   "coarse_solver_setup_s": t_setup},
        "ax": {"value": dofs_total / (ms_ax * 1e-3) / 1e9, "unit": "GDOF/s", "ms": ms_ax,
               "what": "Ap = broadcast(constraint(A p)) on the finest level, 16 B per stored DOF",
               "hbm_frac_of_measured": 16.0 * dofs_local / (ms_ax * 1e-3) / 1e9 / peak,
               "hbm_frac_of_nominal_8TBs": 16.0 * dofs_local / (ms_ax * 1e-3) / 1e9 / 8000.0},
        "vcycle_model": {"bytes_per_dof": bv, "achieved_gbs": bv * dofs_local / (ms / steps * 1e-3) / 1e9,
                         "hbm_frac_of_measured": bv * dofs_local / (ms / steps * 1e-3) / 1e9 / peak,
                         "hbm_frac_of_nominal_8TBs": bv * dofs_local / (ms / steps * 1e-3) / 1e9 / 8000.0},
        "roofline": {"kernel": "apply_kernel (local operator apply, finest level)", "bound": "hbm",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": apply_bytes, "ms_per_launch": ms_apply},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "residual_history": history,
    }
    g.close()
    del hx
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="auto")
    ap.add_argument("--cells", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the secondary workload (N = 1) / the single-GPU run (N > 1)")
    args = ap.parse_args()
    warmup = max(3, args.warmup)
    steps = max(1, args.steps)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # One workload at every N, so that value_N / (N x value_1) is a strong-scaling efficiency: C4 (BASELINE.json
    # configs[3]: 3D Tet64, refinements = 5, 32^3 cells) -- it fits one GPU, is the largest configuration and the one
    # north_star states the 0.8 target on.  configs[1] (C2: 2D Tri64, refinements = 7, 256^2 cells) is measured in the
    # same run at N = 1 and reported under "also".  At N > 1 rank 0 additionally runs the SAME workload from the SAME
    # inputs on its own GPU ("single_gpu") and the residual histories are compared ("parity_vs_single_gpu").
    name = "C4" if args.workload == "auto" else args.workload
    w = dict(WORKLOADS[name], key=name)
    if args.cells:
        w["c"] = args.cells
        w["name"] = w["name"].replace(f"c={WORKLOADS[name]['c']}", f"c={args.cells}")
    dim, levels = w["dim"], w["levels"]

    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_run(w, steps, warmup)
        line = {
            "impl": "reference", "metric": "fine-grid GDOF/s, multigrid V-cycle (stored finest-level DOFs per second)",
            "value": r["vcycle_gdofs"], "unit": "GDOF/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["name"], "smoothing_steps": SMOOTHING_STEPS, "sample": r["sample"]},
            "ax": {"value": r["ax_gdofs"], "unit": "GDOF/s"},
            "cpu_baseline": {"value": r["vcycle_gdofs"], "unit": "GDOF/s", "cores": r["cores"], "kind": "port",
                             "sample": r["sample"],
                             "note": "threaded C restatement of the reference's CPU algorithm, not Julia"},
            "e2e": {"value": r["vcycle_gdofs"], "unit": "GDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return

    import torch
    import hmgb200 as hmg
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU fallback")
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    device = local_rank

    main_r = measure(w, args, torch, hmg, dist, rank, world, device, steps, warmup, not args.no_e2e, True,
                     with_history=world > 1)
    also = None
    single = None
    parity = None
    if world > 1 and not args.no_also:
        # the same workload, the same inputs, ONE GPU, in this very job (rank 0 alone, the others wait): the
        # denominator of the strong-scaling efficiency free of box-to-box variation, and the multi-GPU parity record
        if rank == 0:
            r1 = measure(w, args, torch, hmg, None, 0, 1, device, min(steps, 3), 3, False, False, with_history=True)
            single = {"workload": r1["config"]["workload"], "value": r1["value"], "unit": "GDOF/s",
                      "ms_per_step": r1["ms_per_step"], "ax": r1["ax"]["value"],
                      "residual_history": r1["residual_history"]}
            hp, h1 = np.array(main_r["residual_history"]), np.array(r1["residual_history"])
            parity = {"max_rel_diff": float(np.max(np.abs(hp - h1) / h1)), "tolerance": 1e-10, "cycles": len(h1),
                      "what": "residual norm after each of the first V-cycles, partitioned context vs one GPU, same inputs"}
        dist.barrier()
    if world == 1 and args.workload == "auto" and not args.no_also:
        w2 = dict(WORKLOADS["C2"], key="C2")
        # the 2D workload needs ~140 GB of device memory during the coarse factorisation and 9 GB of pinned host memory
        if 1.5 * 8 * nf_of(2, 8) * 2 * w2["c"] ** 2 > host_memory_available():
            w2["c"] = 192
            w2["name"] = w2["name"].replace("c=256", "c=192 (host memory too small for c=256)")
        r2 = measure(w2, args, torch, hmg, dist, rank, world, device, min(steps, 5), 3, False, False)
        also = {"C2": {k: r2[k] for k in ("value", "ms_per_step", "config", "ax", "vcycle_model", "roofline", "gpu_launches")}}
        also["C2"]["unit"] = "GDOF/s"

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(w, 3, 1, target_seconds=20.0)
        cpu = {"value": r["vcycle_gdofs"], "unit": "GDOF/s", "cores": r["cores"], "kind": "port", "sample": r["sample"],
               "ax_gdofs": r["ax_gdofs"], "note": "threaded C restatement of the reference's CPU algorithm, not Julia"}

    if rank == 0:
        line = {
            "metric": "fine-grid GDOF/s, multigrid V-cycle (stored finest-level DOFs per second)",
            "value": main_r["value"], "unit": "GDOF/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": main_r["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": main_r["config"], "ax": main_r["ax"], "vcycle_model": main_r["vcycle_model"],
            "roofline": main_r["roofline"], "clocks": main_r["clocks"], "e2e": main_r["e2e"],
            "gpu_launches": main_r["gpu_launches"], "cpu_baseline": cpu, "also": also, "single_gpu": single,
            "parity_vs_single_gpu": parity,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
