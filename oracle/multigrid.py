"""Geometric multigrid V-cycle with a CG smoother.

Oracle (test infrastructure only).  Restates src/multigrid.jl, including its
quirks: dots over ALL stored entries (interface duplicates over-counted, :54,64,67),
levels below the top always use steps = 2 (:109), the restricted residual is the
local un-summed one (:102-105).
"""
import numpy as np
import scipy.sparse.linalg as spla

from .implicit import (new_state, broadcast_interfaces, apply_constraint, copy_to_base,
                       distribute)
from .operators import mul, local_residual


class LevelState:
    """src/multigrid.jl:7-25 -- five (Nf, Ne) matrices."""

    def __init__(self, implicit, level):
        self.x = new_state(implicit, level)
        self.b = new_state(implicit, level)
        self.r = new_state(implicit, level)
        self.p = new_state(implicit, level)
        self.Ap = new_state(implicit, level)


class BaseLevel:
    """src/multigrid.jl:30-41 -- direct solver on the interior of the base mesh."""

    def __init__(self, A_interior, total_nodes, interior_nodes):
        self.solve = spla.splu(A_interior.tocsc()).solve      # stands in for CHOLMOD
        self.b = np.zeros(total_nodes)
        self.b_interior = np.zeros(len(interior_nodes))
        self.interior_nodes = interior_nodes


def dot(a, b):
    return float(np.dot(a.ravel(order="K"), b.ravel(order="K")))


def smoothing_steps(steps, implicit, ops, curr, k):
    """src/multigrid.jl:46-71 -- ``steps`` CG iterations started from the current x."""
    local_residual(implicit, ops, curr, k)
    broadcast_interfaces(curr.r, implicit, k)
    curr.p[:, :] = curr.r
    rsqrprev = dot(curr.r, curr.r)
    for _ in range(steps):
        curr.Ap[:, :] = 0.0
        mul(1.0, implicit.base, ops, curr.p, curr.Ap)
        apply_constraint(curr.Ap, k, ops.constraint, implicit)
        broadcast_interfaces(curr.Ap, implicit, k)
        alpha = rsqrprev / dot(curr.p, curr.Ap)
        curr.x += alpha * curr.p
        curr.r += (-alpha) * curr.Ap
        rsqr = dot(curr.r, curr.r)
        curr.p[:, :] = curr.r + (rsqr / rsqrprev) * curr.p
        rsqrprev = rsqr


def vcycle(implicit, base, ops, levels, k, steps=2):
    """src/multigrid.jl:73-119.  ``k`` is 1-based; ``ops``/``levels`` are 0-based lists."""
    if k == 1:
        l1 = levels[0]
        broadcast_interfaces(l1.b, implicit, 1)
        copy_to_base(base.b, l1.b, implicit)
        base.b_interior[:] = base.b[base.interior_nodes]
        tmp = base.solve(base.b_interior)
        base.b[:] = 0.0
        base.b[base.interior_nodes] = tmp
        distribute(l1.x, base.b, implicit)
        return
    curr = levels[k - 1]
    nxt = levels[k - 2]
    P = implicit.reference.interops[k - 2]
    smoothing_steps(steps, implicit, ops[k - 1], curr, k)
    local_residual(implicit, ops[k - 1], curr, k)
    nxt.b[:, :] = P.T @ curr.r                 # restrict_to!, src/interpolation.jl:64-74
    nxt.x[:, :] = 0.0
    vcycle(implicit, base, ops, levels, k - 1)  # steps NOT forwarded (:109)
    curr.x += P @ nxt.x                         # interpolate_and_sum_to!, :52-62
    smoothing_steps(steps, implicit, ops[k - 1], curr, k)
