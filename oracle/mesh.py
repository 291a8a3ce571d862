"""Explicit simplex meshes, edge graph, uniform refinement, hypercube generators.

Oracle (test infrastructure only).  Restates src/grid.jl, src/sparse_graph.jl,
src/tri/refine.jl, src/tet/refine.jl, src/tri/generate_grid.jl,
src/tet/generate_grid.jl.  0-based indices; orderings identical to the reference.
"""
import numpy as np

TET_FACES = ((0, 1, 2), (0, 1, 3), (0, 2, 3), (1, 2, 3))      # src/grid.jl:89
TET_EDGES = ((0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3))  # src/grid.jl:90
TRI_EDGES = ((0, 1), (0, 2), (1, 2))                          # src/grid.jl:91


class Mesh:
    """src/grid.jl:19-22.  nodes: (Nn, dim) float64; elements: (Ne, dim+1) int64."""

    def __init__(self, nodes, elements):
        self.nodes = np.ascontiguousarray(nodes, dtype=np.float64)
        self.elements = np.ascontiguousarray(elements, dtype=np.int64)
        assert self.nodes.ndim == 2 and self.elements.ndim == 2
        assert self.elements.shape[1] == self.nodes.shape[1] + 1

    @property
    def dim(self):
        return self.nodes.shape[1]

    @property
    def nnodes(self):
        return self.nodes.shape[0]

    @property
    def nelements(self):
        return self.elements.shape[0]


class SparseGraph:
    """src/sparse_graph.jl:4-7 -- CSR of edges (from < to), ``adj`` sorted per row."""

    def __init__(self, ptr, adj):
        self.ptr = ptr
        self.adj = adj
        n = len(ptr) - 1
        # flat (from, to) key for vectorised edge_index
        frm = np.repeat(np.arange(n, dtype=np.int64), np.diff(ptr))
        self._n = n
        self._key = frm * n + adj
        self.frm = frm

    @property
    def nedges(self):
        return len(self.adj)

    def edge_index(self, n1, n2):
        """src/sparse_graph.jl:14-15 -- natural index of the sorted edge (n1 < n2)."""
        key = np.asarray(n1, dtype=np.int64) * self._n + np.asarray(n2, dtype=np.int64)
        idx = np.searchsorted(self._key, key)
        assert np.all(self._key[idx] == key)
        return idx


def edge_graph(mesh):
    """src/sparse_graph.jl:20-48 -- sorted, de-duplicated upper-triangular adjacency.

    The order of ``adj`` (ascending ``from``, then ascending ``to``) defines the
    numbering of the midpoint nodes created by refine_uniformly."""
    N = mesh.elements.shape[1]
    pairs = []
    for i in range(N):
        for j in range(i + 1, N):
            a = mesh.elements[:, i]
            b = mesh.elements[:, j]
            pairs.append(np.stack([np.minimum(a, b), np.maximum(a, b)], axis=1))
    pairs = np.unique(np.concatenate(pairs, axis=0), axis=0)  # lexicographic + unique
    ptr = np.zeros(mesh.nnodes + 1, dtype=np.int64)
    np.add.at(ptr, pairs[:, 0] + 1, 1)
    ptr = np.cumsum(ptr)
    return SparseGraph(ptr, pairs[:, 1].copy())


def sort_element_nodes(elements):
    """src/sorting_tricks.jl:34-39."""
    return np.sort(elements, axis=1)


def _refine_nodes(mesh, graph):
    mid = (mesh.nodes[graph.frm] + mesh.nodes[graph.adj]) / 2
    return np.concatenate([mesh.nodes, mid], axis=0)


def _refine_tri(mesh, graph):
    """src/tri/refine.jl:5-43 -- red refinement, children index-sorted on creation."""
    Nn = mesh.nnodes
    t = mesh.elements
    a = graph.edge_index(t[:, 0], t[:, 1]) + Nn
    b = graph.edge_index(t[:, 0], t[:, 2]) + Nn
    c = graph.edge_index(t[:, 1], t[:, 2]) + Nn
    kids = np.stack([
        np.stack([t[:, 0], a, b], axis=1),
        np.stack([t[:, 1], c, a], axis=1),
        np.stack([t[:, 2], b, c], axis=1),
        np.stack([a, c, b], axis=1),
    ], axis=1)                                   # (Nt, 4, 3)
    kids = np.sort(kids, axis=2).reshape(-1, 3)  # sort_bitonic on each child
    return Mesh(_refine_nodes(mesh, graph), kids)


_BEY = ((0, 4, 5, 6), (4, 1, 7, 8), (5, 7, 2, 9), (6, 8, 9, 3),
        (4, 5, 6, 8), (4, 5, 7, 8), (5, 6, 8, 9), (5, 7, 8, 9))   # src/tet/refine.jl:46-47


def _refine_tet(mesh, graph):
    """src/tet/refine.jl:5-54 -- Bey refinement; children keep Bey's vertex order."""
    Nn = mesh.nnodes
    t = mesh.elements
    parts = [t[:, 0], t[:, 1], t[:, 2], t[:, 3]]
    for i in range(4):
        for j in range(i + 1, 4):
            lo = np.minimum(t[:, i], t[:, j])
            hi = np.maximum(t[:, i], t[:, j])
            parts.append(graph.edge_index(lo, hi) + Nn)
    parts = np.stack(parts, axis=1)              # (Nt, 10)
    kids = np.stack([parts[:, list(q)] for q in _BEY], axis=1).reshape(-1, 4)
    return Mesh(_refine_nodes(mesh, graph), kids)


def refine_uniformly(mesh, graph=None, times=1):
    """src/grid.jl:59-64 (times=) and the per-type methods."""
    if graph is not None:
        return _refine_tri(mesh, graph) if mesh.dim == 2 else _refine_tet(mesh, graph)
    for _ in range(times):
        g = edge_graph(mesh)
        mesh = _refine_tri(mesh, g) if mesh.dim == 2 else _refine_tet(mesh, g)
    return mesh


def hypercube(dim, n, scale=1.0, origin=None, sorted_=True):
    """src/tri/generate_grid.jl:6-35 and src/tet/generate_grid.jl:6-45.

    Note the reference numbers nodes x-outer / z-inner but looks them up through a
    column-major ``reshape(1:Nn, n+1, ...)``, i.e. with the axes reversed; this is
    restated literally."""
    if origin is None:
        origin = (1.0,) * dim
    origin = np.asarray(origin, dtype=np.float64)
    n1 = n + 1
    if dim == 2:
        xs, ys = np.meshgrid(np.arange(n1), np.arange(n1), indexing="ij")
        nodes = np.stack([scale * xs.ravel() + origin[0], scale * ys.ravel() + origin[1]], axis=1)
        nn = lambda x, y: x + y * n1                          # column-major reshape
        x, y = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
        x = x.ravel()
        y = y.ravel()                                         # y inner
        c1, c2, c3, c4 = nn(x, y), nn(x + 1, y), nn(x, y + 1), nn(x + 1, y + 1)
        el = np.stack([np.stack([c1, c2, c3], 1), np.stack([c2, c3, c4], 1)], axis=1).reshape(-1, 3)
    else:
        xs, ys, zs = np.meshgrid(np.arange(n1), np.arange(n1), np.arange(n1), indexing="ij")
        nodes = np.stack([scale * xs.ravel() + origin[0], scale * ys.ravel() + origin[1],
                          scale * zs.ravel() + origin[2]], axis=1)
        nn = lambda x, y, z: x + y * n1 + z * n1 * n1
        x, y, z = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
        x, y, z = x.ravel(), y.ravel(), z.ravel()
        c1, c2, c3, c4 = nn(x, y, z), nn(x + 1, y, z), nn(x, y + 1, z), nn(x + 1, y + 1, z)
        c5, c6, c7, c8 = nn(x, y, z + 1), nn(x + 1, y, z + 1), nn(x, y + 1, z + 1), nn(x + 1, y + 1, z + 1)
        tets = [(c1, c2, c3, c7), (c1, c2, c5, c7), (c2, c4, c3, c7),
                (c2, c4, c7, c8), (c2, c6, c5, c7), (c2, c6, c7, c8)]
        el = np.stack([np.stack(t, 1) for t in tets], axis=1).reshape(-1, 4)
    if sorted_:
        el = sort_element_nodes(el)
    return Mesh(nodes, el)


def affine_map(mesh, el):
    """src/grid.jl:120-135 -- J = [p2-p1 ... ], shift = p1 (el: vertex ids)."""
    p = mesh.nodes[np.asarray(el)]
    J = (p[1:] - p[0]).T
    return J, p[0]


def cube5_mesh():
    """The 5-tet unit cube used by the reference's tests (test/test_operator.jl:12-13)."""
    nodes = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (1, 1, 0), (0, 0, 1), (1, 0, 1), (0, 1, 1), (1, 1, 1)]
    elements = np.array([(1, 2, 3, 5), (2, 3, 4, 8), (3, 5, 7, 8), (2, 5, 6, 8), (2, 3, 5, 8)]) - 1
    return Mesh(np.array(nodes, dtype=np.float64), elements)
