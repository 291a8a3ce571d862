"""Local linear operators on the implicit fine grid: build + apply.

Oracle (test infrastructure only).  Restates src/build_local_operators.jl:6-32 and
src/apply_local_operators.jl.
"""
import numpy as np

from .fem import Geometry
from .implicit import apply_constraint


class SimpleDiffusion:
    """src/build_local_operators.jl:15-19 -- L = -a * Laplace, with its constraint ``bc``."""

    def __init__(self, ops, bc, a):
        self.ops = ops            # ops[k][l] CSC, one level
        self.bc = bc
        self.a = a


class L2PlusDivAGrad:
    """src/build_local_operators.jl:26-32 -- L = lambda - div(sigma grad); mutable lambda/constraint."""

    def __init__(self, diffusion_terms, mass, constraint, lam, sigmas):
        self.diffusion_terms = diffusion_terms   # ops[k][l] CSC, one level
        self.mass = mass
        self.constraint = constraint
        self.lam = lam
        self.sigmas = np.asarray(sigmas, dtype=np.float64)   # (Ne, dim)


def element_tensors(base, A):
    """Per coarse element: P = Jinv' * (sigma .* Jinv) and |det J|.

    src/apply_local_operators.jl:59-63 (SimpleDiffusion: P = Jinv' * Jinv) and :101-105
    (L2PlusDivAGrad), with Jinv = inv(J') (src/cell_values.jl:113)."""
    g = Geometry(base)
    Jinv = g.inv_jac
    if isinstance(A, SimpleDiffusion):
        P = np.transpose(Jinv, (0, 2, 1)) @ Jinv
    else:
        P = np.transpose(Jinv, (0, 2, 1)) @ (A.sigmas[:, :, None] * Jinv)
    return P, g.det


def mul(alpha, base, A, x, y):
    """y <- alpha * A * x + y, column-local: no interface sum, no constraint.

    src/apply_local_operators.jl:40-72 (SimpleDiffusion) and :85-120 (L2PlusDivAGrad);
    the inner kernel is my_A_mul_B!, :125-133."""
    dim = base.dim
    P, det = element_tensors(base, A)
    if isinstance(A, SimpleDiffusion):
        for i in range(dim):
            for j in range(dim):
                coef = alpha * P[:, i, j] * det * A.a                    # :69
                y += A.ops[i][j] @ (x * coef[None, :])
    else:
        for i in range(dim):
            for j in range(dim):
                coef = alpha * det * P[:, i, j]                          # :112
                y += A.diffusion_terms[i][j] @ (x * coef[None, :])
        coef = alpha * A.lam * det                                       # :116-118
        if np.any(coef != 0):
            y += A.mass @ (x * coef[None, :])
    return y


def local_residual(implicit, A, curr, k):
    """src/apply_local_operators.jl:7-27 -- r = b - A x (local), then the constraint."""
    curr.r[:, :] = curr.b
    mul(-1.0, implicit.base, A, curr.x, curr.r)
    z = A.bc if isinstance(A, SimpleDiffusion) else A.constraint
    apply_constraint(curr.r, k, z, implicit)
    return curr.r
