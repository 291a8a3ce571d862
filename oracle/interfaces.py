"""Base-mesh interface maps and boundary lists.

Oracle (test infrastructure only).  Restates src/interface.jl.  Vectorised with a
stable lexicographic sort, which is what the reference's LSD counting sort is
(src/sorting_tricks.jl:44-76): within every cell the owners appear in ascending
element index.
"""
import numpy as np

from .mesh import TET_FACES, TET_EDGES, TRI_EDGES


class SparseCellToElementMap:
    """src/interface.jl:31-35 -- CSR: cell i owns values[offset[i]:offset[i+1]].

    ``values`` is split into ``element`` and ``local_id`` arrays (ElementId,
    src/interface.jl:6-9)."""

    def __init__(self, offset, cells, element, local_id):
        self.offset = np.asarray(offset, dtype=np.int64)
        self.cells = np.asarray(cells, dtype=np.int64)
        self.element = np.asarray(element, dtype=np.int64)
        self.local_id = np.asarray(local_id, dtype=np.int64)

    @property
    def ncells(self):
        return len(self.offset) - 1 if len(self.offset) else 0

    def valrange(self, i):
        """src/interface.jl:47."""
        return range(self.offset[i], self.offset[i + 1])


def empty_map(N):
    """src/interface.jl:38-42."""
    return SparseCellToElementMap(np.zeros(0, np.int64), np.zeros((0, N), np.int64),
                                  np.zeros(0, np.int64), np.zeros(0, np.int64))


def compress(cells, element, local_id):
    """src/interface.jl:317-351 -- group a sorted cell list into CSR form."""
    cells = np.asarray(cells, dtype=np.int64).reshape(len(element), -1)
    n = len(element)
    if n == 0:
        return SparseCellToElementMap(np.zeros(1, np.int64), cells, element, local_id)
    new = np.ones(n, dtype=bool)
    new[1:] = np.any(cells[1:] != cells[:-1], axis=1)
    starts = np.nonzero(new)[0]
    offset = np.concatenate([starts, [n]])
    return SparseCellToElementMap(offset, cells[starts], element, local_id)


def _list_cells_with_element(mesh, local_cells):
    """src/interface.jl:124-197 -- (cell nodes, element, local id), element-major order."""
    el = mesh.elements
    ne = el.shape[0]
    nl = len(local_cells)
    cells = np.stack([el[:, list(c)] for c in local_cells], axis=1).reshape(ne * nl, -1)
    element = np.repeat(np.arange(ne, dtype=np.int64), nl)
    local_id = np.tile(np.arange(nl, dtype=np.int64), ne)
    return cells, element, local_id


def _stable_sort(cells, element, local_id):
    """radix_sort!(list, nnodes, N), src/interface.jl:84,98,111."""
    keys = tuple(cells[:, k] for k in range(cells.shape[1] - 1, -1, -1))
    order = np.lexsort(keys)  # stable; last key is primary
    return cells[order], element[order], local_id[order]


def _group_sizes(cells):
    n = cells.shape[0]
    new = np.ones(n, dtype=bool)
    new[1:] = np.any(cells[1:] != cells[:-1], axis=1)
    gid = np.cumsum(new) - 1
    sizes = np.bincount(gid)
    return gid, sizes


def _remove_singletons(cells, element, local_id):
    """src/sorting_tricks.jl:130-154."""
    gid, sizes = _group_sizes(cells)
    keep = sizes[gid] > 1
    return cells[keep], element[keep], local_id[keep]


def _remove_repeated_pairs(cells, element, local_id):
    """src/sorting_tricks.jl:222-248 -- a group of g equal cells leaves g mod 2 (its last)."""
    gid, sizes = _group_sizes(cells)
    n = cells.shape[0]
    last = np.ones(n, dtype=bool)
    last[:-1] = gid[1:] != gid[:-1]
    keep = last & (sizes[gid] % 2 == 1)
    return cells[keep], element[keep], local_id[keep]


def _local_cells(dim):
    nodes = tuple((i,) for i in range(dim + 1))
    edges = TRI_EDGES if dim == 2 else TET_EDGES
    return nodes, edges


class Interfaces:
    """src/interface.jl:55-60."""

    def __init__(self, all_nodes, nodes, edges, faces):
        self.all_nodes = all_nodes
        self.nodes = nodes
        self.edges = edges
        self.faces = faces


def interfaces(mesh):
    """src/interface.jl:65-117 -- cells shared by >= 2 elements (+ all_nodes)."""
    lnodes, ledges = _local_cells(mesh.dim)
    c, e, l = _stable_sort(*_list_cells_with_element(mesh, lnodes))
    all_nodes = compress(c, e, l)
    nodes = compress(*_remove_singletons(c, e, l))
    edges = compress(*_remove_singletons(*_stable_sort(*_list_cells_with_element(mesh, ledges))))
    if mesh.dim == 3:
        faces = compress(*_remove_singletons(*_stable_sort(*_list_cells_with_element(mesh, TET_FACES))))
    else:
        faces = empty_map(3)
    return Interfaces(all_nodes, nodes, edges, faces)


def _intersect(cells, element, local_id, wanted):
    """src/interface.jl:291-309 -- keep entries whose cell occurs in sorted ``wanted``."""
    nn = int(max(cells.max(initial=0), wanted.max(initial=0))) + 1

    def key(a):
        k = np.zeros(a.shape[0], dtype=np.int64)
        for j in range(a.shape[1]):
            k = k * nn + a[:, j]
        return k
    keep = np.isin(key(cells), key(wanted))
    return cells[keep], element[keep], local_id[keep]


def list_boundary_nodes_edges_faces(mesh):
    """src/interface.jl:207-284 -- boundary cells with ALL their owners."""
    lnodes, ledges = _local_cells(mesh.dim)
    if mesh.dim == 3:
        fc, fe, fl = _remove_repeated_pairs(*_stable_sort(*_list_cells_with_element(mesh, TET_FACES)))
        be = np.concatenate([fc[:, [0, 1]], fc[:, [0, 2]], fc[:, [1, 2]]], axis=0)
        be = np.unique(be, axis=0)
        ec, ee, el_ = _intersect(*_stable_sort(*_list_cells_with_element(mesh, ledges)), be)
        faces = compress(fc, fe, fl)
    else:
        ec, ee, el_ = _remove_repeated_pairs(*_stable_sort(*_list_cells_with_element(mesh, ledges)))
        be = ec
        faces = empty_map(3)
    bn = np.unique(be.ravel()).reshape(-1, 1)
    nc, ne, nl = _intersect(*_stable_sort(*_list_cells_with_element(mesh, lnodes)), bn)
    return compress(nc, ne, nl), compress(ec, ee, el_), faces


def list_interior_nodes(mesh):
    """src/grid.jl:176-202 -- nodes not on any boundary face (sorted)."""
    local_faces = TET_FACES if mesh.dim == 3 else TRI_EDGES
    fc, _, _ = _remove_repeated_pairs(*_stable_sort(*_list_cells_with_element(mesh, local_faces)))
    boundary = np.unique(fc.ravel())
    mask = np.ones(mesh.nnodes, dtype=bool)
    mask[boundary] = False
    return np.nonzero(mask)[0].astype(np.int64)
