"""Row N4 of SURVEY.md 8f on the host: refined_mesh / construct_full_grid of the product (library tables) against
the oracle's restatement of src/multilevel_reference.jl:41-61 and src/implicit_fine_grid.jl:41-78, and the .vtu
files export_domain / write_vtu produce (src/examples/homogenized_coefficients.jl:71-87, src/utils.jl:14-19)."""
import numpy as np
import pytest

import hmgb200 as hmg
from oracle.mesh import Mesh as OMesh
from oracle.implicit import ImplicitFineGrid as OImplicit, construct_full_grid as o_full_grid


@pytest.mark.parametrize("dim,levels", [(2, 1), (2, 4), (2, 6), (3, 1), (3, 3), (3, 5)])
def test_refined_mesh_matches_oracle(dim, levels):
    mesh, _ = hmg.inputs.checkerboard_problem(dim, 1)
    oimp = OImplicit(OMesh(mesh.nodes, mesh.elements), levels)
    for level in range(1, levels + 1):
        ref = hmg.vtk.refined_mesh(dim, levels, level)
        oref = oimp.refined_mesh(level)
        assert np.array_equal(ref.nodes, oref.nodes)            # dyadic coordinates: exact
        assert np.array_equal(ref.elements, oref.elements)      # same elements, same order, index-sorted


@pytest.mark.parametrize("dim,c,levels,level", [(2, 3, 4, 3), (3, 2, 3, 2), (3, 2, 3, 1)])
def test_construct_full_grid_matches_oracle(dim, c, levels, level):
    mesh, _ = hmg.inputs.checkerboard_problem(dim, c, ordered=True)
    # a sheared copy: the affine map must use the element's own Jacobian, not the unit cell's
    shear = np.eye(dim) + 0.25 * np.triu(np.ones((dim, dim)), 1)
    mesh = hmg.Mesh(mesh.nodes @ shear.T, mesh.elements)
    full = hmg.vtk.construct_full_grid(mesh, levels, level)
    ofull = o_full_grid(OImplicit(OMesh(mesh.nodes, mesh.elements), levels), level)
    assert np.array_equal(full.elements, ofull.elements)
    assert np.allclose(full.nodes, ofull.nodes, rtol=0, atol=1e-14)
    assert full.nnodes == mesh.nelements * hmg.inputs.nf_of_level(dim, level)


@pytest.mark.parametrize("dim", [2, 3])
def test_vtu_round_trip_and_export_domain(dim, tmp_path):
    mesh, sigma = hmg.inputs.checkerboard_problem(dim, 3)
    path = hmg.vtk.export_domain(mesh, sigma, str(tmp_path / "checkerboard"))
    assert path.endswith("checkerboard.vtu")
    back, pd, cd = hmg.vtk.read_vtu(path)
    assert np.array_equal(back.nodes, mesh.nodes) and np.array_equal(back.elements, mesh.elements)
    assert pd == {} and list(cd) == ["a"]
    assert np.array_equal(cd["a"], sigma)                       # dim components per element
    text = open(path).read()
    assert 'type="UnstructuredGrid"' in text and 'header_type="UInt64"' in text
    assert f'NumberOfPoints="{mesh.nnodes}"' in text and f'NumberOfCells="{mesh.nelements}"' in text
    # scalar point data on the explicit grid of a level, as export_unknown writes it
    full = hmg.vtk.construct_full_grid(mesh, 3, 2)
    v = np.random.default_rng(0).random(full.nnodes)
    p2 = hmg.vtk.write_vtu(str(tmp_path / "ahom_0.vtu"), full, point_data={"v": v})
    back, pd, cd = hmg.vtk.read_vtu(p2)
    assert np.array_equal(pd["v"], v) and back.nelements == full.nelements
    with pytest.raises(ValueError):
        hmg.vtk.write_vtu(str(tmp_path / "bad"), full, point_data={"v": v[:-1]})


@pytest.mark.parametrize("dim,c,levels", [(2, 3, 5), (3, 2, 4)])
def test_local_rhs_host_matches_oracle(dim, c, levels):
    """local_rhs! (src/implicit_fine_grid.jl:391-409) as the product's host mirror computes it (from the library's own
    refined reference mesh) against the oracle, on a sheared and scaled mesh (|det J| differs from 1)."""
    from oracle.implicit import local_rhs as o_local_rhs
    mesh, _ = hmg.inputs.checkerboard_problem(dim, c)
    shear = 1.5 * (np.eye(dim) + 0.25 * np.triu(np.ones((dim, dim)), 1))
    mesh = hmg.Mesh(mesh.nodes @ shear.T, mesh.elements)
    oimp = OImplicit(OMesh(mesh.nodes, mesh.elements), levels)
    expect = o_local_rhs(np.zeros((oimp.nf(levels), mesh.nelements), order="F"), oimp)
    got = hmg.local_rhs_host(mesh, levels)
    assert got.shape == expect.shape and got.flags.f_contiguous
    assert np.allclose(got, expect, rtol=1e-13, atol=0)
    # the functional of f = 1 sums to the volume of the domain, every element counted once
    assert abs(got.sum() - (c ** dim) * abs(np.linalg.det(shear))) <= 1e-12 * got.sum()
