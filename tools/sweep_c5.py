#!/usr/bin/env python
"""BASELINE.json configs[4] (C5): matrix-free A*x and V-cycle throughput on a random piecewise coefficient
field (recipe of tools/generate_st1_field.jl, alpha = 1, p = 1.5, seed 2) as a function of the number of
stored DOFs.  The field comes from the library's cuFFT generator (hmg_generate_field, noise drawn on the device); the
convergence of three V-cycles is recorded beside the timings.  One JSON line per problem; device-timed with CUDA events
on the library's stream.

    python tools/sweep_c5.py [--max-gb 60] > profiles/rNN_c5_sweep.jsonl
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import hmgb200 as hmg
from bench import vcycle_bytes_per_dof, measured_peak, SMOOTHING_STEPS

CASES = [(2, 8, 5), (2, 16, 6), (2, 32, 7), (2, 64, 8), (2, 128, 8), (3, 4, 4), (3, 8, 5), (3, 16, 5), (3, 24, 6), (3, 32, 6)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-gb", type=float, default=60.0)
    args = ap.parse_args()
    peak, _ = measured_peak()
    for dim, c, levels in CASES:
        nf = hmg.inputs.nf_of_level(dim, levels)
        ne = 2 * c ** 2 if dim == 2 else 6 * c ** 3
        if 9.5 * 8e-9 * nf * ne > args.max_gb:
            continue
        mesh, _ = hmg.inputs.checkerboard_problem(dim, c)
        cells = hmg.inputs.random_field_cells_device(dim, c, seed=2, alpha=1.0, p=1.5)       # cuFFT, on the device
        sigma = hmg.inputs.conductivity_per_element(mesh, cells, (c / 2.0 + 1.0,) * dim)
        g = hmg.ImplicitFineGrid(mesh, levels, sigma, lam=1.0)
        st = g.state(levels)
        rng = np.random.default_rng(3)
        host = np.empty((nf, ne), order="F")
        step = max(1, (1 << 25) // nf)
        for c0 in range(0, ne, step):
            host[:, c0:c0 + step] = rng.random((nf, min(step, ne - c0)))
        st.x.set(host)
        del host
        hmg.broadcast_interfaces(st.x, g, levels)
        hmg.apply_constraint(st.x, levels, g)
        st.p.copy_from(st.x)
        hmg.rhs_a_xi_grad_v(st.b, g, np.ones(dim) / dim ** 0.5)
        bl = hmg.BaseLevel(g)
        hist = [float(v) for v in hmg.vcycles(g, bl, levels, SMOOTHING_STEPS, 3)]
        dofs = nf * ne
        reps = 20 if dofs < 2e8 else 10
        g.time_op(0, levels, 0, 3)
        ms_ax = g.time_op(0, levels, 0, reps) / reps
        g.time_op(1, levels, SMOOTHING_STEPS, 2)
        ms_v = g.time_op(1, levels, SMOOTHING_STEPS, 5) / 5
        bv = vcycle_bytes_per_dof(dim, levels)
        print(json.dumps({
            "dim": dim, "cells_per_side": c, "grids": levels, "coarse_elements": ne, "stored_dofs": dofs,
            "sigma_min_max": [float(sigma.min()), float(sigma.max())], "field": "hmg_generate_field (cuFFT), seed 2",
            "residual_after_vcycles_1_2_3": hist,
            "ax_ms": ms_ax, "ax_gdofs": dofs / ms_ax / 1e6, "ax_hbm_frac": 16.0 * dofs / ms_ax / 1e6 / peak,
            "vcycle_ms": ms_v, "vcycle_gdofs": dofs / ms_v / 1e6, "vcycle_hbm_frac": bv * dofs / ms_v / 1e6 / peak}), flush=True)
        g.close()


if __name__ == "__main__":
    main()
