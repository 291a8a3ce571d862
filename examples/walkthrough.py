#!/usr/bin/env python
"""The multigrid walk-through of the reference's documentation (docs/src/index.md:162-305) over the device API:
base mesh of n x n cells, random checkerboard conductivities, `levels` implicit grids, right-hand side f = 1 built
locally, random initial guess made consistent on the interfaces and zero on the boundary, then V-cycles with the
logged residual norm(zero_out_all_but_one!(r)) after each of them.

    python examples/walkthrough.py [--dim 2] [-n 32] [--levels 3] [--steps 1] [--cycles 100] [--seed 0] [--save LEVEL]

Needs a CUDA device (there is no CPU fallback).  Not part of the test-suite: every call it makes is covered by
tests/test_gpu_*.py."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hmgb200 as hmg  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=2)
    ap.add_argument("-n", type=int, default=32)
    ap.add_argument("--levels", type=int, default=3)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--cycles", type=int, default=100)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--save", type=int, default=0, help="export x on the nodes of this level to walkthrough_x.vtu")
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    # base = hypercube(elementtype, n); a = conductivity_per_element(base, generate_conductivity(base, n))
    base = hmg.inputs.hypercube(a.dim, a.n)                                   # domain [1, n + 1]^dim
    cells = np.where(rng.random((a.n,) * a.dim + (a.dim,)) < 0.5, 1.0, 9.0)
    cond = hmg.inputs.conductivity_per_element(base, cells, (0.0,) * a.dim)
    lam = 1.0
    # implicit = ImplicitFineGrid(base, levels) + constraint + level_operators + level_states, all on the device
    g = hmg.ImplicitFineGrid(base, a.levels, cond, lam=lam)
    try:
        # F = cholesky(assemble_checkerboard(base, a, lam)[interior, interior]); base_level = BaseLevel(...)
        base_level = hmg.BaseLevel(g)
        finest = g.state(a.levels)
        hmg.local_rhs(finest.b, g)                                            # integrate v dx locally
        nf = g.nf(a.levels)
        finest.x.set(np.asfortranarray(rng.random((nf, base.nelements))))     # rand!(x): local values
        hmg.broadcast_interfaces(finest.x, g, a.levels)                       # sum boundaries
        hmg.apply_constraint(finest.x, a.levels, g)                           # impose b.c.
        print(f"Implicit grid: base mesh has {base.nnodes} nodes and {base.nelements} elements; finest level "
              f"({a.levels}) has {nf} nodes per element; at most {nf * base.nelements} unknowns.")
        for i in range(1, a.cycles + 1):
            res = hmg.vcycle(g, base_level, a.levels, a.steps, resnorm=True)
            print(f"After cycle {i}: norm(finest_level.r) = {res!r}")
        if a.save:
            print("saved", hmg.vtk.export_unknown(g, finest.x, 0, a.save, "walkthrough_x"))
    finally:
        g.close()


if __name__ == "__main__":
    main()
