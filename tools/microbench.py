"""Device-timed micro-benchmarks of the library operations (development helper).

    python tools/microbench.py DIM CELLS LEVELS [reps]
ops: 0 fused global product, 3 apply only (y = A x + Dirichlet), 4 interface kernel only, 2 mul! (y += A x),
1 V-cycle.  Prints ms, GDOF/s and the 16 B/DOF HBM fraction of the measured peak."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import hmgb200 as hmg


def main():
    dim, c, levels = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 20
    peak = 6541.8
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    mesh, sigma = hmg.inputs.checkerboard_problem(dim, c)
    if os.environ.get("HMG_MB_SORT"):      # experiment: elements grouped by their type inside the cell
        k = 2 if dim == 2 else 6
        order = np.arange(mesh.nelements).reshape(-1, k).T.reshape(-1)
        mesh = hmg.Mesh(mesh.nodes, mesh.elements[order])
        sigma = np.ascontiguousarray(sigma[order])
    g = hmg.ImplicitFineGrid(mesh, levels, sigma, lam=1.0)
    nf = g.nf(levels)
    dofs = nf * mesh.nelements
    rng = np.random.default_rng(0)
    st = g.state(levels)
    x = np.asfortranarray(rng.random((nf, mesh.nelements)))
    st.x.set(x); st.b.set(x); st.p.set(x)
    hmg.broadcast_interfaces(st.p, g, levels)
    out = {"dim": dim, "c": c, "levels": levels, "dofs": dofs, "env": {k: v for k, v in os.environ.items() if k.startswith("HMG_")}}
    # (op, name, algorithmic bytes per stored DOF)
    ops = ((3, "apply", 16), (11, "apply_dot", 16), (4, "interface", 16), (13, "interface_pairs", 16), (14, "interface_multi", 16), (0, "global_product", 16), (5, "residual", 24),
           (2, "mul", 24), (12, "fused_p_product", 32), (6, "cg_update", 48), (16, "x_update", 24), (7, "p_update", 24), (8, "copy_dot", 16), (9, "restrict", 9), (10, "interp", 17))
    for op, name, bpd in ops:
        if levels < 2 and op in (9, 10):
            continue
        try:
            g.time_op(op, levels, 0, 3)
        except Exception as ex:
            out[name] = str(ex)[:60]
            continue
        ms = g.time_op(op, levels, 0, reps) / reps
        out[name] = {"ms": round(ms, 4), "gdofs": round(dofs / ms / 1e6, 2), "bytes_per_dof": bpd,
                     "hbm_frac": round(bpd * dofs / ms / 1e6 / peak, 3)}
    if len(sys.argv) > 5 and sys.argv[5] == "v":
        bl = hmg.BaseLevel(g)
        g.time_op(1, levels, 3, 2)
        ms = g.time_op(1, levels, 3, 5) / 5
        out["vcycle"] = {"ms": round(ms, 3), "gdofs": round(dofs / ms / 1e6, 3)}
    print(json.dumps(out))
    g.close()


if __name__ == "__main__":
    main()
