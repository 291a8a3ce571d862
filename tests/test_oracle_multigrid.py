"""The oracle's multigrid (parity-unpinned by the reference) is validated by itself:
V-cycles converge to the direct solution of the explicitly assembled fine problem, and the restated
driver reproduces the docstring's sample of the reference within sampling noise."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla
from scipy.spatial import cKDTree

from oracle.mesh import hypercube, refine_uniformly, sort_element_nodes
from oracle.fem import (assemble_matrix, assemble_vector, build_local_diffusion_operators,
                        build_local_mass_matrices)
from oracle.interfaces import list_boundary_nodes_edges_faces, list_interior_nodes
from oracle.implicit import (ImplicitFineGrid, ZeroDirichletConstraint, broadcast_interfaces, apply_constraint,
                             zero_out_all_but_one, construct_full_grid, local_rhs)
from oracle.operators import L2PlusDivAGrad
from oracle.multigrid import LevelState, BaseLevel, vcycle
from oracle.driver import conductivity_per_element, checkerboard_homogenization, make_base
from oracle.reference_element import refined_element


@pytest.mark.parametrize("dim,c,levels,cycles", [(2, 4, 3, 12), (3, 2, 3, 8)])
def test_vcycles_converge_to_direct_fine_solution(dim, c, levels, cycles):
    base = hypercube(dim, c)
    rng = np.random.default_rng(1)
    cells = np.where(rng.random((c,) * dim + (dim,)) < 0.5, 1.0, 9.0)
    cond = conductivity_per_element(base, cells, (0.0,) * dim)
    lam = 1.0
    implicit = ImplicitFineGrid(base, levels)
    z = ZeroDirichletConstraint(*list_boundary_nodes_edges_faces(base))
    ops = [L2PlusDivAGrad(d, m, z, lam, cond) for d, m in
           zip(build_local_diffusion_operators(implicit.reference), build_local_mass_matrices(implicit.reference))]
    states = [LevelState(implicit, i + 1) for i in range(levels)]
    top = states[-1]
    top.x[:, :] = rng.random(top.x.shape)
    broadcast_interfaces(top.x, implicit, levels)
    apply_constraint(top.x, levels, z, implicit)
    local_rhs(top.b, implicit)
    interior = list_interior_nodes(base)
    A = assemble_matrix(base, sigma=cond, lam=lam)
    bl = BaseLevel(A[interior][:, interior], base.nnodes, interior)
    # direct solution on the explicitly refined mesh (children of element e are contiguous)
    fine = refine_uniformly(base, times=levels - 1)
    fine.elements = sort_element_nodes(fine.elements)
    cond_f = np.repeat(cond, (2 ** dim) ** (levels - 1), axis=0)
    Af = assemble_matrix(fine, sigma=cond_f, lam=lam)
    bf = assemble_vector(fine)
    intf = list_interior_nodes(fine)
    xf = np.zeros(fine.nnodes)
    xf[intf] = spla.spsolve(Af[intf][:, intf].tocsc(), bf[intf])
    _, mp = cKDTree(fine.nodes).query(construct_full_grid(implicit, levels).nodes)
    errs, res = [], []
    for _ in range(cycles):
        vcycle(implicit, bl, ops, states, levels, 3)
        r = zero_out_all_but_one(top.r.copy(order="F"), implicit, levels)
        res.append(np.linalg.norm(r))
        errs.append(np.max(np.abs(top.x.ravel(order="F") - xf[mp])))
    assert errs[-1] < 1e-4 * errs[0]
    assert res[-1] < 1e-4 * res[0]
    assert all(b < 0.5 * a for a, b in zip(errs[:-1], errs[1:]))     # contraction every cycle


def test_driver_reproduces_the_docstring_sample_within_sampling_noise():
    """src/examples/homogenized_coefficients.jl:156-158: checkerboard_homogenization(5, Tri64,
    refinements = 1, tolerance = 1e-5) returned 1.6163911... on one unseeded random field; the domain has
    112^2 cells, so another random field gives the same value to about a percent."""
    n, dim, refs = 5, 2, 1
    base, R = make_base(dim, n)
    rng = np.random.default_rng(1)
    cells = np.where(rng.random((2 * R,) * dim + (dim,)) < 0.5, 1.0, 9.0)
    nf = refined_element(refs + 1, dim).levels[-1].nnodes
    x0 = rng.random((nf, base.nelements))
    sigma, hist = checkerboard_homogenization(n, dim, refinements=refs, tolerance=1e-5, sigma_cells=cells, x0=x0)
    assert len(hist) == 2                       # two outer steps run for n = 5 (one domain shrink)
    assert abs(sigma - 1.6163911040833774) < 0.03
