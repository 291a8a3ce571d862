"""CPU tests of the library's host-side setup against the oracle (no GPU needed).

The stencil tables, lattice numbering, interface pairing, interpolation structure, interface
maps and Dirichlet classes that the kernels consume are expanded by the library's host-only
introspection entry points (include/hmg_introspect.h) and compared with the oracle's explicit
sparse operators and maps.
"""
import ctypes as C

import numpy as np
import pytest

import hmgb200 as hmg
from hmgb200 import _lib as L
from oracle.mesh import Mesh as OMesh, refine_uniformly, sort_element_nodes, cube5_mesh, hypercube
from oracle.reference_element import refined_element
from oracle.fem import build_local_diffusion_operators_level, mass_matrix, Geometry
from oracle.interfaces import interfaces, list_boundary_nodes_edges_faces, list_interior_nodes
from oracle.implicit import ImplicitFineGrid as OImplicit

lib = L.load()
vp = lambda a: a.ctypes.data_as(C.c_void_p)


def host_reference(dim, nlevels, level):
    sizes = np.zeros(8, dtype=np.int64)
    L.check_host(lib.hmg_host_reference(dim, nlevels, level, vp(sizes), None, None, None))
    h2l = np.zeros(sizes[1], dtype=np.int32)
    mt = C.c_double()
    L.check_host(lib.hmg_host_reference(dim, nlevels, level, vp(sizes), vp(h2l), None, C.byref(mt)))
    return sizes, h2l, mt.value


@pytest.mark.parametrize("dim,nlevels", [(2, 6), (3, 5)])
def test_lattice_numbering_is_a_permutation_matching_node_coordinates(dim, nlevels):
    ref = refined_element(nlevels, dim)
    for level in range(1, nlevels + 1):
        sizes, h2l, _ = host_reference(dim, nlevels, level)
        mesh = ref.levels[level - 1]
        m = 2 ** (level - 1)
        assert sizes[0] == m and sizes[1] == mesh.nnodes
        assert sorted(h2l) == list(range(mesh.nnodes))
        # 2D: lexicographic order of the lattice coordinates (i, j), j fastest;
        # 3D: diagonal-plane order (i + j, i, k), k fastest (csrc/lattice.hpp)
        lat = np.rint(mesh.nodes * m).astype(int)
        if dim == 2:
            order = np.lexsort((lat[:, 1], lat[:, 0]))
        else:
            order = np.lexsort((lat[:, 2], lat[:, 0], lat[:, 0] + lat[:, 1]))
        expect = np.empty(mesh.nnodes, dtype=int)
        expect[order] = np.arange(mesh.nnodes)
        assert np.array_equal(h2l, expect)


@pytest.mark.parametrize("dim,nlevels", [(2, 6), (3, 5)])
def test_stencil_tables_reproduce_the_reference_operators(dim, nlevels):
    """sum_kl |J| P_kl ops[k,l] + lambda |J| mass (src/apply_local_operators.jl:105-118) as a dense
    matrix == the library's class-table stencil expanded through the kernels' offset functions."""
    ref = refined_element(nlevels, dim)
    rng = np.random.default_rng(5)
    for level in range(1, nlevels + 1):
        mesh = ref.levels[level - 1]
        nf = mesh.nnodes
        ops = build_local_diffusion_operators_level(mesh)
        mass = mass_matrix(mesh)
        Pm = rng.standard_normal((dim, dim))
        Pm = Pm @ Pm.T
        lam_detj = 0.7
        expect = lam_detj * mass.toarray()
        coef = []
        for k in range(dim):
            for l in range(k, dim):
                coef.append(Pm[k, l])
            for l in range(dim):
                expect += Pm[k, l] * ops[k][l].toarray()
        coef = np.array(coef + [lam_detj])
        dense = np.zeros((nf, nf), order="F")
        L.check_host(lib.hmg_host_local_matrix(dim, nlevels, level, vp(coef), vp(dense)))
        scale = np.abs(expect).max()
        assert np.abs(dense - expect).max() <= 1e-14 * scale
        _, _, mt = host_reference(dim, nlevels, level)
        assert abs(mt - mass.sum()) <= 1e-15


@pytest.mark.parametrize("dim,nlevels", [(2, 8), (3, 6)], ids=["tri-L8", "tet-L6"])
def test_benchmarked_depths_numbering_and_stencil(dim, nlevels):
    """The finest level of the BENCHMARKED hierarchies (C2: 2D with 8 grids, C4: 3D with 6 grids): lattice numbering and
    the expanded stencil against the oracle's explicit operators, compared in sparse form (the dense matrix of 3D level 6
    has 6545^2 entries)."""
    import scipy.sparse as sp
    ref = refined_element(nlevels, dim)
    level = nlevels
    mesh = ref.levels[level - 1]
    nf = mesh.nnodes
    m = 2 ** (level - 1)
    sizes, h2l, mt = host_reference(dim, nlevels, level)
    assert sizes[0] == m and sizes[1] == nf
    lat = np.rint(mesh.nodes * m).astype(int)
    order = np.lexsort((lat[:, 1], lat[:, 0])) if dim == 2 else np.lexsort((lat[:, 2], lat[:, 0], lat[:, 0] + lat[:, 1]))
    expect_perm = np.empty(nf, dtype=int)
    expect_perm[order] = np.arange(nf)
    assert np.array_equal(h2l, expect_perm)
    ops = build_local_diffusion_operators_level(mesh)
    mass = mass_matrix(mesh)
    rng = np.random.default_rng(6)
    Pm = rng.standard_normal((dim, dim))
    Pm = Pm @ Pm.T
    lam_detj = 0.7
    expect = lam_detj * mass.tocsr()
    coef = []
    for k in range(dim):
        for l in range(k, dim):
            coef.append(Pm[k, l])
        for l in range(dim):
            expect = expect + Pm[k, l] * ops[k][l].tocsr()
    coef = np.array(coef + [lam_detj])
    dense = np.zeros((nf, nf), order="F")
    L.check_host(lib.hmg_host_local_matrix(dim, nlevels, level, vp(coef), vp(dense)))
    got = sp.csr_matrix(dense)
    del dense
    diff = (got - expect)
    scale = np.abs(expect.data).max()
    assert (np.abs(diff.data).max() if diff.nnz else 0.0) <= 1e-14 * scale
    assert abs(mt - mass.sum()) <= 1e-15
    P = ref.interops[level - 2]
    dense = np.zeros(P.shape, order="F")
    L.check_host(lib.hmg_host_transfer_matrix(dim, nlevels, level, vp(dense)))
    assert (sp.csr_matrix(dense) != P.tocsr()).nnz == 0


@pytest.mark.parametrize("dim,nlevels", [(2, 6), (3, 5)])
def test_transfer_structure_matches_interpolation_operator(dim, nlevels):
    ref = refined_element(nlevels, dim)
    for level in range(2, nlevels + 1):
        P = ref.interops[level - 2].toarray()
        dense = np.zeros(P.shape, order="F")
        L.check_host(lib.hmg_host_transfer_matrix(dim, nlevels, level, vp(dense)))
        assert np.array_equal(dense, P)


@pytest.mark.parametrize("dim,nlevels", [(2, 6), (3, 5)])
def test_interface_pairing_matches_local_numbering(dim, nlevels):
    """The q-th paired node of every local face / edge / vertex is the q-th entry of the
    reference's ascending lists (src/multilevel_reference.jl:125-203)."""
    ref = refined_element(nlevels, dim)
    for level in range(1, nlevels + 1):
        nb = ref.numbering[level - 1]
        kinds = [(1, nb.edges_interior), (2, [[n] for n in nb.nodes])]
        if dim == 3:
            kinds.append((0, nb.faces_interior))
        for kind, lists in kinds:
            perm = None
            for lid, expect in enumerate(lists):
                cnt = C.c_int64()
                L.check_host(lib.hmg_host_interface_rows(dim, nlevels, level, kind, lid, None, C.byref(cnt)))
                assert cnt.value == len(expect)
                rows = np.zeros(max(1, cnt.value), dtype=np.int32)
                L.check_host(lib.hmg_host_interface_rows(dim, nlevels, level, kind, lid, vp(rows), C.byref(cnt)))
                rows = [int(v) for v in rows[:cnt.value]]
                expect = [int(v) for v in expect]
                # the library may enumerate the nodes of a cell in another order than the reference
                # (edges: monotone along the edge), but it must be the SAME re-ordering for every
                # local cell, so that k-th <-> k-th pairing of owners is preserved
                if perm is None:
                    assert sorted(rows) == expect
                    perm = [expect.index(r) for r in rows]
                assert rows == [expect[q] for q in perm]


def _meshes():
    out = []
    b = refine_uniformly(cube5_mesh(), times=2)
    b.elements = sort_element_nodes(b.elements)
    out.append(b)
    out.append(hypercube(3, 3))
    out.append(hypercube(2, 5))
    t = refine_uniformly(hypercube(2, 2), times=1)
    t.elements = sort_element_nodes(t.elements)
    out.append(t)
    # irregular topologies: a third of the elements removed at random (holes, re-entrant corners, faces and edges
    # that end up on the boundary from the inside, nodes no element uses any more), the rest in shuffled order
    rng = np.random.default_rng(11)
    for dim, c in ((3, 3), (2, 6)):
        full = hypercube(dim, c)
        keep = rng.permutation(full.nelements)[: (2 * full.nelements) // 3]
        out.append(type(full)(full.nodes, full.elements[keep]))
    return out


MESH_IDS = ["cube5x2", "hyper3", "hyper2", "tri-refined", "tet-holes-shuffled", "tri-holes-shuffled"]


def host_map(mesh, kind):
    el1 = np.ascontiguousarray(mesh.elements + 1, dtype=np.int64)
    nc, nent = C.c_int64(), C.c_int64()
    L.check_host(lib.hmg_host_topology(mesh.dim, mesh.nelements, mesh.nnodes, vp(el1), kind, C.byref(nc), C.byref(nent),
                                       None, None, None))
    off = np.zeros(nc.value + 1, dtype=np.int64)
    el = np.zeros(nent.value, dtype=np.int64)
    lid = np.zeros(nent.value, dtype=np.int64)
    L.check_host(lib.hmg_host_topology(mesh.dim, mesh.nelements, mesh.nnodes, vp(el1), kind, C.byref(nc), C.byref(nent),
                                       vp(off), vp(el), vp(lid)))
    return off, el, lid


@pytest.mark.parametrize("mesh", _meshes(), ids=MESH_IDS)
def test_interface_maps_match_oracle(mesh):
    inter = interfaces(mesh)
    kinds = [(1, inter.edges), (2, inter.nodes), (3, inter.all_nodes)]
    if mesh.dim == 3:
        kinds.append((0, inter.faces))
    for kind, om in kinds:
        off, el, lid = host_map(mesh, kind)
        if kind == 3:
            # the library's all-nodes CSR is indexed by node id; unused nodes have empty rows
            sizes = np.diff(off)
            assert np.array_equal(np.nonzero(sizes)[0], om.cells[:, 0])
            off = np.concatenate([[0], np.cumsum(sizes[sizes > 0])])
        assert np.array_equal(off, om.offset)
        assert np.array_equal(el, om.element)
        assert np.array_equal(lid, om.local_id)


@pytest.mark.parametrize("mesh", _meshes(), ids=MESH_IDS)
def test_dirichlet_classes_and_interior_nodes_match_oracle(mesh):
    dim = mesh.dim
    el1 = np.ascontiguousarray(mesh.elements + 1, dtype=np.int64)
    cmask = np.zeros(mesh.nelements, dtype=np.uint16)
    interior = np.zeros(mesh.nnodes, dtype=np.uint8)
    L.check_host(lib.hmg_host_boundary(dim, mesh.nelements, mesh.nnodes, vp(el1), vp(cmask), vp(interior)))
    assert np.array_equal(np.nonzero(interior)[0], list_interior_nodes(mesh))
    nodes, edges, faces = list_boundary_nodes_edges_faces(mesh)
    expect = np.zeros(mesh.nelements, dtype=np.uint16)
    for kind, om in ((2, nodes), (1, edges), (0, faces)):
        for e, l in zip(om.element, om.local_id):
            expect[e] |= np.uint16(1 << lib.hmg_host_class_of(dim, kind, int(l)))
    assert np.array_equal(cmask, expect)


def test_element_coefficients_match_oracle():
    rng = np.random.default_rng(9)
    for mesh in _meshes():
        dim = mesh.dim
        # perturb the nodes so that J is generic
        nodes = mesh.nodes + 0.05 * rng.standard_normal(mesh.nodes.shape)
        sig = rng.uniform(1, 9, size=(mesh.nelements, dim))
        g = Geometry(OMesh(nodes, mesh.elements))
        P = np.transpose(g.inv_jac, (0, 2, 1)) @ (sig[:, :, None] * g.inv_jac)
        stride = 8 if dim == 3 else 4
        coef = np.zeros((mesh.nelements, stride))
        el1 = np.ascontiguousarray(mesh.elements + 1, dtype=np.int64)
        nd = np.ascontiguousarray(nodes)
        L.check_host(lib.hmg_host_element_coefficients(dim, mesh.nelements, mesh.nnodes, vp(nd), vp(el1), vp(sig),
                                                       vp(coef), stride))
        c = 0
        for k in range(dim):
            for l in range(k, dim):
                assert np.allclose(coef[:, c], g.det * P[:, k, l], rtol=1e-12, atol=1e-12)
                c += 1
        assert np.allclose(coef[:, c], g.det, rtol=1e-14)


@pytest.mark.parametrize("dim,nlevels,seg_shift", [(2, 8, 4), (2, 6, 5), (2, 4, 2), (3, 6, 4), (3, 4, 4)])
def test_apply_sweeps_reproduce_the_local_operator(dim, nlevels, seg_shift):
    """The task enumeration and line sweeps of the apply kernel (csrc/apply_core.cuh: class dispatch,
    compile-time segment weights, register sliding windows, row geometry in the diagonal-plane order),
    executed on the host for one element, equal the dense local operator
    sum_kl |J| P_kl ops[k,l] + lambda |J| mass (src/apply_local_operators.jl:105-118) times x."""
    rng = np.random.default_rng(11)
    for level in range(1, nlevels + 1):
        sizes, h2l, _ = host_reference(dim, nlevels, level)
        nf = int(sizes[1])
        Pm = rng.standard_normal((dim, dim))
        Pm = Pm @ Pm.T
        coef = np.array([Pm[k, l] for k in range(dim) for l in range(k, dim)] + [0.6])
        dense = np.zeros((nf, nf), order="F")
        L.check_host(lib.hmg_host_local_matrix(dim, nlevels, level, vp(coef), vp(dense)))
        x_h = rng.standard_normal(nf)
        x_l = np.empty(nf)
        x_l[h2l] = x_h                      # lattice order
        y_l = np.full(nf, np.nan)
        info = np.zeros(4, dtype=np.int64)
        L.check_host(lib.hmg_host_apply_sweep(dim, nlevels, level, seg_shift, vp(coef), vp(x_l), vp(y_l), vp(info)))
        expect = dense @ x_h
        assert np.abs(y_l[h2l] - expect).max() <= 1e-13 * np.abs(expect).max(), (level, info)
        # the row window of every task fits the ring, the ring fits one SM
        assert 0 < info[1] + 64 <= info[2] and info[3] <= 227 * 1024, info
