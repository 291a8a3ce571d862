"""The threaded C restatement (CPU baseline) agrees with the numpy oracle."""
import numpy as np
import pytest

import hmgb200 as hmg
from oracle.mesh import Mesh as OMesh
from oracle.fem import build_local_diffusion_operators, build_local_mass_matrices, assemble_matrix
from oracle.interfaces import list_boundary_nodes_edges_faces, list_interior_nodes
from oracle.implicit import ImplicitFineGrid, ZeroDirichletConstraint, broadcast_interfaces, apply_constraint, \
    zero_out_all_but_one, local_rhs
from oracle.operators import L2PlusDivAGrad, mul
from oracle.multigrid import LevelState, BaseLevel, vcycle
from oracle.cref import CpuReference


def setup(dim, c, levels):
    mesh, sigma = hmg.inputs.checkerboard_problem(dim, c)
    base = OMesh(mesh.nodes, mesh.elements)
    imp = ImplicitFineGrid(base, levels)
    z = ZeroDirichletConstraint(*list_boundary_nodes_edges_faces(base))
    ops = [L2PlusDivAGrad(d, m, z, 0.9, sigma) for d, m in
           zip(build_local_diffusion_operators(imp.reference), build_local_mass_matrices(imp.reference))]
    return base, imp, z, ops, sigma


@pytest.mark.parametrize("dim,c,levels", [(2, 4, 4), (3, 2, 3)])
def test_cref_primitives_and_vcycle_match_numpy_oracle(dim, c, levels):
    base, imp, z, ops, sigma = setup(dim, c, levels)
    ref = CpuReference(imp, ops, nthreads=3)
    rng = np.random.default_rng(0)
    x = np.asfortranarray(rng.random((imp.nf(levels), base.nelements)))
    y = np.asfortranarray(rng.random(x.shape))
    y2 = y.copy(order="F")
    ref.mul(-0.4, levels, x, y)
    mul(-0.4, base, ops[-1], x, y2)
    assert np.max(np.abs(y - y2)) <= 1e-13 * np.max(np.abs(y2))
    a, b = x.copy(order="F"), x.copy(order="F")
    assert np.array_equal(ref.broadcast_interfaces(a, levels), broadcast_interfaces(b, imp, levels))
    assert np.array_equal(ref.apply_constraint(a, levels, z), apply_constraint(b, levels, z, imp))
    assert np.array_equal(ref.zero_out_all_but_one(a, levels), zero_out_all_but_one(b, imp, levels))

    interior = list_interior_nodes(base)
    A = assemble_matrix(base, sigma=sigma, lam=0.9)[interior][:, interior]
    s1 = [LevelState(imp, l) for l in range(1, levels + 1)]
    s2 = [LevelState(imp, l) for l in range(1, levels + 1)]
    for s in (s1, s2):
        s[-1].x[:, :] = x
        broadcast_interfaces(s[-1].x, imp, levels)
        apply_constraint(s[-1].x, levels, z, imp)
        local_rhs(s[-1].b, imp)
    bl1 = BaseLevel(A, base.nnodes, interior)
    bl2 = BaseLevel(A, base.nnodes, interior)
    for _ in range(3):
        ref.vcycle(bl1, s1, levels, 3)
        vcycle(imp, bl2, ops, s2, levels, 3)
        n1 = np.linalg.norm(zero_out_all_but_one(s1[-1].r, imp, levels))
        n2 = np.linalg.norm(zero_out_all_but_one(s2[-1].r, imp, levels))
        assert abs(n1 - n2) <= 1e-11 * n2
    assert np.max(np.abs(s1[-1].x - s2[-1].x)) <= 1e-11 * np.max(np.abs(s2[-1].x))
