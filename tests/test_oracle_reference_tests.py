"""The reference's own tests, restated against the oracle (this is what pins it).

Each test cites the reference test file it restates (paths relative to the
reference repository root).  Indices are 0-based here, 1-based there.
"""
import numpy as np
import pytest

from oracle.mesh import (Mesh, refine_uniformly, hypercube, sort_element_nodes, cube5_mesh,
                         edge_graph)
from oracle.sorting import (sort_bitonic, radix_sort, remove_singletons, remove_repeated_pairs,
                            remove_duplicates, left_minus_right, complement, binary_search)
from oracle.reference_element import refined_element, nodes_on_ref_faces, nodes_on_ref_edges
from oracle.interfaces import compress, list_boundary_nodes_edges_faces, list_interior_nodes
from oracle.fem import build_local_diffusion_operators, assemble_matrix, Geometry
from oracle.implicit import (ImplicitFineGrid, ZeroDirichletConstraint, broadcast_interfaces,
                             construct_full_grid, distribute, new_state)
from oracle.operators import SimpleDiffusion, mul


def test_bitonic():
    """test/bitonic.jl -- every permutation of 1..4 items sorts."""
    import itertools
    for n in (1, 2, 3, 4):
        for p in itertools.permutations(range(n)):
            assert sort_bitonic(p) == tuple(range(n))


def test_counting_sort():
    """test/counting_sort.jl -- radix sort of tuples equals a lexicographic sort, stably."""
    rng = np.random.default_rng(0)
    v = [tuple(int(a) for a in rng.integers(0, 7, size=3)) + (i,) for i in range(200)]
    out = radix_sort(v, key=lambda t: t[:3])
    assert [t[:3] for t in out] == sorted(t[:3] for t in v)
    for a, b in zip(out[:-1], out[1:]):
        if a[:3] == b[:3]:
            assert a[3] < b[3]          # stable


def test_tricks():
    """test/tricks.jl:5-36 -- every vector of the reference's file, verbatim (the values are data, not indices)."""
    # "Remove duplicates from sorted vector" (test/tricks.jl:5-9)
    assert remove_duplicates([]) == []
    assert remove_duplicates([2]) == [2]
    assert remove_duplicates([1, 1, 2, 3, 4, 4, 4, 5]) == [1, 2, 3, 4, 5]
    # "Remove singletons from sorted vector" (test/tricks.jl:11-18)
    assert remove_singletons([]) == []
    assert remove_singletons([2]) == []
    assert remove_singletons([2, 3]) == []
    assert remove_singletons([1, 1]) == [1, 1]
    assert remove_singletons([1, 2, 2, 4, 4, 4, 5]) == [2, 2, 4, 4, 4]
    assert remove_singletons([1, 1, 2, 3, 4, 4, 4, 5]) == [1, 1, 4, 4, 4]
    # "Remove sorted array from sorted array" (test/tricks.jl:20-27)
    assert left_minus_right([], []) == []
    assert left_minus_right([2], []) == [2]
    assert left_minus_right([], [1]) == []
    assert left_minus_right([1, 2, 3], [4, 5, 6]) == [1, 2, 3]
    assert left_minus_right([1, 3, 6, 9], [3, 9]) == [1, 6]
    assert left_minus_right([1, 2, 3, 4], [1, 2, 3, 4]) == []
    # "Remove repeated pairs array" (test/tricks.jl:29-36)
    assert remove_repeated_pairs([]) == []
    assert remove_repeated_pairs([1]) == [1]
    assert remove_repeated_pairs([1, 1]) == []
    assert remove_repeated_pairs([1, 1, 2]) == [2]
    assert remove_repeated_pairs([1, 1, 2, 3, 3, 4]) == [2, 4]
    assert remove_repeated_pairs([1, 1, 2, 3, 3, 4, 4]) == [2]


def test_tricks_more():
    """further cases of the same helpers plus complement / binary_search (src/sorting_tricks.jl:197-217, src/utils.jl)."""
    assert remove_duplicates([1, 1, 2, 3, 3, 3, 4]) == [1, 2, 3, 4]
    assert remove_singletons([1, 2, 2, 3, 4, 4, 4, 5]) == [2, 2, 4, 4, 4]
    assert remove_repeated_pairs([1, 2, 2]) == [1]
    assert left_minus_right([1, 2, 3, 4, 5, 6], [2, 4, 7]) == [1, 3, 5, 6]
    assert list(complement([1, 3, 4], 6)) == [0, 2, 5]
    assert list(complement([], 3)) == [0, 1, 2]
    v = [1, 3, 3, 5, 8]
    assert binary_search(v, 3, 0, 4) == 1
    assert binary_search(v, 8, 0, 4) == 4


def test_generated_grids():
    """test/generated_grids.jl:4-10."""
    mesh = hypercube(3, 20)
    assert np.all(np.diff(mesh.elements, axis=1) > 0)
    assert mesh.nnodes == 21 ** 3
    assert mesh.nelements == 6 * 20 ** 3
    # every tet has positive volume 1/6 and the tets tile the cube
    g = Geometry(mesh)
    assert np.allclose(g.det, 1.0)
    tri = hypercube(2, 7)
    assert tri.nnodes == 64 and tri.nelements == 98
    assert np.allclose(Geometry(tri).det, 1.0)


def test_refined_reference_element():
    """test/refined_reference_element.jl:5-37."""
    N = 8
    tets = refined_element(N, 3)
    assert tets.levels[0].nnodes == 4
    assert tets.levels[1].nnodes == 10
    f = nodes_on_ref_faces(tets.levels[0])
    assert [list(map(int, a)) for a in f] == [[0, 1, 2], [0, 1, 3], [0, 2, 3], [1, 2, 3]]
    for i in range(1, N + 1):
        for nodes in nodes_on_ref_faces(tets.levels[i - 1]):
            assert len(nodes) == sum(range(1, 2 ** (i - 1) + 2))
    e = nodes_on_ref_edges(tets.levels[0])
    assert [list(map(int, a)) for a in e] == [[0, 1], [0, 2], [0, 3], [1, 2], [1, 3], [2, 3]]
    for i in range(1, N + 1):
        for nodes in nodes_on_ref_edges(tets.levels[i - 1]):
            assert len(nodes) == 2 ** (i - 1) + 1


def test_reference_sizes():
    """nodes / elements / edges per level."""
    tri = refined_element(8, 2)
    assert [l.nnodes for l in tri.levels] == [3, 6, 15, 45, 153, 561, 2145, 8385]
    assert [l.nelements for l in tri.levels] == [4 ** k for k in range(8)]
    tet = refined_element(6, 3)
    assert [l.nnodes for l in tet.levels] == [4, 10, 35, 165, 969, 6545]
    assert [l.nelements for l in tet.levels] == [8 ** k for k in range(6)]
    assert [edge_graph(l).nedges for l in tet.levels] == [6, 25, 130, 804, 5576, 41360]
    # hierarchical numbering: level k nodes are the first rows of every finer level
    for a, b in zip(tet.levels[:-1], tet.levels[1:]):
        assert np.array_equal(a.nodes, b.nodes[:a.nnodes])


def test_sparse_cell_to_element():
    """test/sparse_cell_to_element.jl:4-27 (0-based: offsets shift by one)."""
    m = compress([(1, 2), (1, 2), (2, 3), (2, 3)], np.array([1, 2, 3, 5]), np.array([2, 3, 4, 6]))
    assert list(m.offset) == [0, 2, 4]
    assert m.cells.tolist() == [[1, 2], [2, 3]]
    assert list(m.element) == [1, 2, 3, 5] and list(m.local_id) == [2, 3, 4, 6]
    m = compress([(1, 2), (2, 3)], np.array([1, 3]), np.array([2, 4]))
    assert list(m.offset) == [0, 1, 2]
    assert m.cells.tolist() == [[1, 2], [2, 3]]


@pytest.mark.parametrize("times,refs", [(2, 4), (3, 3)])
def test_implicit_grid_interfaces_match(times, refs):
    """test/implicit_grid.jl:8-93 -- the k-th local node of every owner of an interface
    node / edge / face is the same physical point (the reference uses times=3, refs=5;
    (3,3) and (2,4) keep the CPU suite short while covering the same code)."""
    coarse = refine_uniformly(cube5_mesh(), times=times)
    coarse.elements = sort_element_nodes(coarse.elements)
    implicit = ImplicitFineGrid(coarse, refs)
    g = Geometry(coarse)
    for level in range(1, refs + 1):
        ref_mesh = implicit.refined_mesh(level)
        for rows, cols, cell, first in implicit._groups(level, implicit.interfaces):
            xs = np.einsum("eij,ekj->eki", g.J[cols], ref_mesh.nodes[rows]) + g.shift[cols][:, None, :]
            firsts = xs[np.nonzero(first)[0]][cell]
            assert np.allclose(xs, firsts, rtol=0, atol=1e-12)


def test_interpolation_reproduces_linears():
    """test/interpolation.jl:8-35 (total_levels = 5 instead of 6)."""
    total_levels = 5
    coarse = cube5_mesh()
    coarse.elements = sort_element_nodes(coarse.elements)
    implicit = ImplicitFineGrid(coarse, total_levels)
    direction = np.random.default_rng(3).standard_normal(3)
    xs = 10.0 + coarse.nodes @ direction
    ys = new_state(implicit, 1)
    distribute(ys, xs, implicit)
    for level in range(2, total_levels + 1):
        ys = implicit.reference.interops[level - 2] @ ys
        full = construct_full_grid(implicit, level)
        assert np.allclose(10.0 + full.nodes @ direction, ys.ravel(order="F"), rtol=1e-13, atol=1e-13)


def _example_operator(levels, times):
    """test/test_operator.jl:9-69."""
    base = refine_uniformly(cube5_mesh(), times=times)
    base.elements = sort_element_nodes(base.elements)
    implicit = ImplicitFineGrid(base, levels)
    rng = np.random.default_rng(11)
    local_x = np.asfortranarray(rng.random((implicit.nf(levels), base.nelements)))
    broadcast_interfaces(local_x, implicit, levels)
    local_y = new_state(implicit, levels)
    constraint = ZeroDirichletConstraint(*list_boundary_nodes_edges_faces(base))
    local_A = SimpleDiffusion(build_local_diffusion_operators(implicit.reference)[levels - 1], constraint, 1.0)
    repeated = construct_full_grid(implicit, levels)
    total_fine = refine_uniformly(base, times=levels - 1)
    total_A = assemble_matrix(total_fine)
    # geometric node matching (the reference does an O(n^2) search with tol 1e-4)
    from scipy.spatial import cKDTree
    dist, mapping = cKDTree(total_fine.nodes).query(repeated.nodes)
    assert np.all(dist < 1e-4)
    total_x = np.zeros(total_fine.nnodes)
    total_x[mapping] = local_x.ravel(order="F")
    mul(1.0, base, local_A, local_x, local_y)
    broadcast_interfaces(local_y, implicit, levels)
    total_y = total_A @ total_x
    return np.max(np.abs(total_y[mapping] - local_y.ravel(order="F")))


def test_operator_matches_assembled_matrix():
    """test/test_operator.jl:68 -- implicit A*x == assembled A*x within 20 eps (levels=4 here)."""
    assert _example_operator(levels=4, times=1) <= 20 * np.finfo(float).eps


@pytest.mark.slow
def test_operator_matches_assembled_matrix_reference_size():
    """test/test_operator.jl exactly as written: levels = 5, base refined once."""
    assert _example_operator(levels=5, times=1) <= 20 * np.finfo(float).eps


def test_list_faces():
    """test/list_faces.jl:6-27 (not part of runtests.jl, but it pins the maps the zero Dirichlet constraint is built
    from, src/interface.jl:207-284): one tetrahedron has 4 boundary faces, 6 edges, 4 nodes; after two refinements
    every face is split in 16, and the boundary nodes are the face lattices minus the doubly counted edges and the
    triply counted vertices.  Also: the library's host tables (csrc/topology.cpp) agree cell for cell."""
    import ctypes as C
    import hmgb200 as hmg
    nodes = np.array([(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)], dtype=np.float64)
    mesh = Mesh(nodes, np.array([[0, 1, 2, 3]]))
    n, e, f = list_boundary_nodes_edges_faces(mesh)
    assert (f.ncells, e.ncells, n.ncells) == (4, 6, 4)
    mesh = refine_uniformly(mesh, times=2)
    mesh.elements = sort_element_nodes(mesh.elements)
    n, e, f = list_boundary_nodes_edges_faces(mesh)
    assert f.ncells == 4 * 16
    assert e.ncells == 2 * 16 * 3                       # 4 * 16 * 3 edges, each counted twice
    assert n.ncells == sum(range(1, 6)) * 4 - 6 * 3 - 2 * 4
    interior = list_interior_nodes(mesh)
    assert len(interior) == mesh.nnodes - n.ncells == 1   # the lattice of m = 4 has one interior point
    # the product's own boundary detection on the same mesh
    lib = hmg.load()
    el1 = np.ascontiguousarray(mesh.elements + 1, dtype=np.int64)
    cmask = np.zeros(mesh.nelements, dtype=np.uint16)
    flag = np.zeros(mesh.nnodes, dtype=np.uint8)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    hmg._lib.check_host(lib.hmg_host_boundary(3, mesh.nelements, mesh.nnodes, vp(el1), vp(cmask), vp(flag)))
    assert np.array_equal(np.nonzero(flag)[0], interior)
    # every boundary face of the oracle shows up as the face class bit of its owning element
    face_bit = [hmg.load().hmg_host_class_of(3, 0, lf) for lf in range(4)]
    cell = np.repeat(np.arange(f.ncells), np.diff(f.offset))
    for el, lid in zip(f.element[cell], f.local_id[cell]):
        assert cmask[el] & (1 << face_bit[lid])
    assert sum(bin(int(c) & sum(1 << b for b in face_bit)).count("1") for c in cmask) == f.ncells
