#!/usr/bin/env python
"""Generates tests/golden/*.json with the CPU oracle (the reference itself cannot run here: no Julia).

The fixtures pin the ORACLE against drift (tests/test_oracle_golden.py re-runs the small cases on the CPU) and let
the GPU tests compare the device driver with the oracle at sizes where re-running the oracle inside the test would
take minutes -- in particular BASELINE.json configs[0], the README example
`checkerboard_homogenization(3, Tri64, refinements=4, tolerance=1e-3, save=nothing)`.

Inputs are seeded (numpy default_rng): sigma per unit cell in {1, 9} with p = 1/2, x0 ~ U(0, 1), xi = ones/sqrt(dim);
the reference draws the same quantities from Julia's unseeded global RNG.

    python tests/golden/make_golden.py            # rewrites the fixtures
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import driver as od          # noqa: E402

# name -> (n, dim, refinements, tolerance, seed)
CASES = {
    "homogenization_tri_n1_r3": (1, 2, 3, 1e-5, 42),
    "homogenization_tet_n0_r2": (0, 3, 2, 1e-4, 42),
    "homogenization_C1_tri_n3_r4": (3, 2, 4, 1e-3, 7),      # BASELINE.json configs[0]
}


def inputs(n, dim, refinements, seed):
    rng = np.random.default_rng(seed)
    radius = od.compute_box_radius(0, n) + od.compute_boundary_layer(1.0, n)
    cells = np.where(rng.random((2 * radius,) * dim + (dim,)) < 0.5, 1.0, 9.0)
    base, _ = od.make_base(dim, n)
    m = 1 << refinements
    nf = (m + 1) * (m + 2) // 2 if dim == 2 else (m + 1) * (m + 2) * (m + 3) // 6
    x0 = np.asfortranarray(rng.random((nf, base.nelements)))
    return cells, x0, base


def main():
    for name, (n, dim, refinements, tol, seed) in CASES.items():
        cells, x0, base = inputs(n, dim, refinements, seed)
        t = time.time()
        sigma, hist = od.checkerboard_homogenization(n, dim, refinements=refinements, tolerance=tol, sigma_cells=cells, x0=x0)
        out = {"n": n, "dim": dim, "refinements": refinements, "tolerance": tol, "seed": seed,
               "coarse_elements": int(base.nelements), "sigma": sigma,
               "history": [[list(map(float, h)) for h in step] for step in hist],
               "oracle_seconds": round(time.time() - t, 1)}
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(out, f, indent=1)
        print(name, sigma, [len(s) for s in hist], out["oracle_seconds"], "s")


if __name__ == "__main__":
    main()
