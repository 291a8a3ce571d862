"""Implicit fine grid and its distributed-vector primitives.

Oracle (test infrastructure only).  Restates src/implicit_fine_grid.jl.  State
vectors are (Nf, Ne) Fortran-ordered float64 arrays: column e = coarse element e,
contiguous -- the memory layout of the reference's Matrix{Float64}.
"""
import numpy as np

from .mesh import Mesh
from .reference_element import refined_element
from .interfaces import interfaces
from .fem import Geometry, assemble_vector


class ZeroDirichletConstraint:
    """src/implicit_fine_grid.jl:80-84."""

    def __init__(self, nodes, edges, faces):
        self.nodes = nodes
        self.edges = edges
        self.faces = faces


class ImplicitFineGrid:
    """src/implicit_fine_grid.jl:6-24."""

    def __init__(self, base, levels):
        assert np.all(np.diff(base.elements, axis=1) > 0), "base elements must be sorted"  # :14
        self.levels = levels
        self.reference = refined_element(levels, base.dim)
        self.interfaces = interfaces(base)
        self.base = base
        self._cache = {}

    @property
    def dim(self):
        return self.base.dim

    def refined_mesh(self, level):
        """1-based level, as in the reference (src/implicit_fine_grid.jl:22)."""
        return self.reference.levels[level - 1]

    def local_numbering(self, level):
        return self.reference.numbering[level - 1]

    def nf(self, level):
        return self.refined_mesh(level).nnodes

    # -- flat index lists -------------------------------------------------
    def _entries(self, smap, lists, level, key):
        """For every (cell, owner) entry of ``smap``: rows = local node list of the owner's
        local cell, col = owner element, cell = cell id.  Cached per (key, level)."""
        ck = (key, level, id(smap))
        if ck not in self._cache:
            if smap.ncells == 0 or len(lists[0]) == 0:
                self._cache[ck] = None
            else:
                table = np.asarray(lists, dtype=np.int64)           # (nlocal, k)
                rows = table[smap.local_id]                          # (entries, k)
                cols = smap.element
                cell = np.repeat(np.arange(smap.ncells), np.diff(smap.offset))
                first = np.zeros(len(cols), dtype=bool)
                first[smap.offset[:-1]] = True
                self._cache[ck] = (rows, cols, cell, first)
            self._cache[("keepalive",) + ck] = smap   # id() stays unique while cached
        return self._cache[ck]

    def _groups(self, level, maps):
        """(rows, cols, cell, first) for faces / edges / nodes of a set of maps."""
        numbering = self.local_numbering(level)
        out = []
        if self.dim == 3:
            out.append(self._entries(maps.faces, numbering.faces_interior, level, "f"))
        out.append(self._entries(maps.edges, numbering.edges_interior, level, "e"))
        out.append(self._entries(maps.nodes, [[n] for n in numbering.nodes], level, "n"))
        return [g for g in out if g is not None]


def new_state(implicit, level):
    return np.zeros((implicit.nf(level), implicit.base.nelements), order="F")


def broadcast_interfaces(x, implicit, level):
    """src/implicit_fine_grid.jl:209-328 -- sum over owners (ascending element index),
    write the sum back to every owner; k-th node of one owner <-> k-th node of the others."""
    for rows, cols, cell, _ in implicit._groups(level, implicit.interfaces):
        vals = x[rows, cols[:, None]]
        buf = np.zeros((cell[-1] + 1, rows.shape[1]))
        np.add.at(buf, cell, vals)              # sequential in entry order = owner order
        x[rows, cols[:, None]] = buf[cell]
    return x


def apply_constraint(x, level, z, implicit):
    """src/implicit_fine_grid.jl:94-139 -- zero every stored copy of a boundary node."""
    for rows, cols, _, _ in implicit._groups(level, z):
        x[rows, cols[:, None]] = 0.0
    return x


def zero_out_all_but_one(x, implicit, level):
    """src/implicit_fine_grid.jl:334-386 -- keep only the first owner's copy."""
    for rows, cols, _, first in implicit._groups(level, implicit.interfaces):
        x[rows[~first], cols[~first, None]] = 0.0
    return x


def copy_to_base(u, v, implicit):
    """src/implicit_fine_grid.jl:148-171 -- first owner's value of every base node."""
    m = implicit.interfaces.all_nodes
    numbering = implicit.local_numbering(1)
    first = m.offset[:-1]
    local = np.asarray(numbering.nodes)[m.local_id[first]]
    u[m.cells[:, 0]] = v[local, m.element[first]]
    return u


def distribute(v, u, implicit):
    """src/implicit_fine_grid.jl:178-202 -- copy base values to every owner."""
    m = implicit.interfaces.all_nodes
    numbering = implicit.local_numbering(1)
    cell = np.repeat(np.arange(m.ncells), np.diff(m.offset))
    local = np.asarray(numbering.nodes)[m.local_id]
    v[local, m.element] = u[m.cells[cell, 0]]
    return v


def construct_full_grid(implicit, level):
    """src/implicit_fine_grid.jl:41-78 -- explicit mesh with interface nodes repeated."""
    base = implicit.base
    ref = implicit.refined_mesh(level)
    g = Geometry(base)
    nodes = np.einsum("eij,nj->eni", g.J, ref.nodes) + g.shift[:, None, :]
    nn = ref.nnodes
    elements = ref.elements[None, :, :] + (np.arange(base.nelements) * nn)[:, None, None]
    return Mesh(nodes.reshape(-1, base.dim), elements.reshape(-1, base.dim + 1))


def local_rhs(b, implicit):
    """src/implicit_fine_grid.jl:391-409 -- functional int v, un-summed."""
    fine = implicit.refined_mesh(implicit.levels)
    b_ref = assemble_vector(fine)
    g = Geometry(implicit.base)
    b[:, :] = b_ref[:, None] * g.det[None, :]
    return b
