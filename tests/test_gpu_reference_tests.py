"""The reference's known-answer test for A*x, run through the CUDA path.

test/test_operator.jl:9-73: the implicit product on a 5-tet cube refined once equals the product
with the matrix assembled on the explicitly refined mesh.  SimpleDiffusion(a = 1) is
L2PlusDivAGrad with sigma = 1 and lambda = 0."""
import numpy as np
import pytest
from scipy.spatial import cKDTree

import hmgb200 as hmg
from oracle.mesh import refine_uniformly, sort_element_nodes, cube5_mesh
from oracle.fem import assemble_matrix
from oracle.implicit import ImplicitFineGrid as OImplicit, construct_full_grid, broadcast_interfaces

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("levels", [3, 5])
def test_operator_matches_assembled_matrix_on_gpu(levels):
    base = refine_uniformly(cube5_mesh(), times=1)
    base.elements = sort_element_nodes(base.elements)
    oimp = OImplicit(base, levels)
    rng = np.random.default_rng(11)
    local_x = np.asfortranarray(rng.random((oimp.nf(levels), base.nelements)))
    broadcast_interfaces(local_x, oimp, levels)
    g = hmg.ImplicitFineGrid(hmg.Mesh(base.nodes, base.elements), levels, np.ones((base.nelements, 3)), lam=0.0)
    st = g.state(levels)
    st.x.set(local_x)
    st.r.fill(0.0)
    hmg.mul(1.0, g, st.x, st.r)
    hmg.broadcast_interfaces(st.r, g, levels)
    local_y = st.r.get()
    total_fine = refine_uniformly(base, times=levels - 1)
    total_A = assemble_matrix(total_fine)
    dist, mapping = cKDTree(total_fine.nodes).query(construct_full_grid(oimp, levels).nodes)
    assert np.all(dist < 1e-4)
    total_x = np.zeros(total_fine.nnodes)
    total_x[mapping] = local_x.ravel(order="F")
    total_y = total_A @ total_x
    # the reference asserts <= 20 eps for values O(1); the stencil form sums in another order
    assert np.max(np.abs(total_y[mapping] - local_y.ravel(order="F"))) <= 200 * np.finfo(float).eps
    g.close()


def test_create_rejects_unsorted_elements():
    mesh, sigma = hmg.inputs.checkerboard_problem(2, 2)
    bad = hmg.Mesh(mesh.nodes, mesh.elements[:, ::-1])
    with pytest.raises(AssertionError):
        hmg.ImplicitFineGrid(bad, 2, sigma)
