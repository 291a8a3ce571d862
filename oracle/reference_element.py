"""Refined reference simplex: levels, local numbering, interpolation operators.

Oracle (test infrastructure only).  Restates src/multilevel_reference.jl and
src/interpolation.jl:7-50.
"""
import numpy as np
import scipy.sparse as sp

from .mesh import Mesh, edge_graph, refine_uniformly, sort_element_nodes
from .sorting import left_minus_right

EPS = np.finfo(np.float64).eps


def reference_element(dim):
    """src/multilevel_reference.jl:3-13."""
    if dim == 2:
        return Mesh(np.array([(0, 0), (1, 0), (0, 1)], dtype=np.float64), np.array([[0, 1, 2]]))
    return Mesh(np.array([(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)], dtype=np.float64),
                np.array([[0, 1, 2, 3]]))


class ReferenceNumbering:
    """src/multilevel_reference.jl:19-25 -- ascending local-index lists."""

    def __init__(self, faces, faces_interior, edges, edges_interior, nodes):
        self.faces = faces
        self.faces_interior = faces_interior
        self.edges = edges
        self.edges_interior = edges_interior
        self.nodes = nodes


def nodes_on_ref_faces(m):
    """src/multilevel_reference.jl:63-70 (Tets)."""
    x = m.nodes
    return [
        list(np.nonzero(x[:, 2] == 0)[0]),
        list(np.nonzero(x[:, 1] == 0)[0]),
        list(np.nonzero(x[:, 0] == 0)[0]),
        list(np.nonzero(x[:, 0] + x[:, 1] + x[:, 2] >= 1 - 10 * EPS)[0]),
    ]


def _is_on_edge(a, b, x):
    """src/multilevel_reference.jl:83-101 -- IsOnEdge(a, b)(x), tolerance 1e-7."""
    diff = b - a
    unit = diff / np.linalg.norm(diff)
    vec = x - a
    proj = vec @ unit
    return np.abs(proj * proj - np.einsum("ij,ij->i", vec, vec)) < 1e-7


def nodes_on_ref_edges(m):
    """src/multilevel_reference.jl:72-78 (Tris) and :103-116 (Tets)."""
    x = m.nodes
    if m.dim == 2:
        return [
            list(np.nonzero(x[:, 1] == 0)[0]),
            list(np.nonzero(x[:, 0] == 0)[0]),
            list(np.nonzero(x[:, 0] + x[:, 1] >= 1 - 10 * EPS)[0]),
        ]
    ref = reference_element(3).nodes
    pairs = ((0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3))
    return [list(np.nonzero(_is_on_edge(ref[a], ref[b], x))[0]) for a, b in pairs]


def get_local_numbering(m):
    """src/multilevel_reference.jl:125-180 (Tets) and :182-203 (Tris)."""
    if m.dim == 2:
        edge_to_nodes = nodes_on_ref_edges(m)
        nodes_to_nodes = [0, 1, 2]
        interior = [left_minus_right(e, nodes_to_nodes) for e in edge_to_nodes]
        return ReferenceNumbering([[]], [[]], edge_to_nodes, interior, nodes_to_nodes)

    face_to_nodes = nodes_on_ref_faces(m)
    ref = reference_element(3).nodes
    x = m.nodes

    def filt(a, b, face):
        idx = np.asarray(face, dtype=np.int64)
        return list(idx[_is_on_edge(ref[a], ref[b], x[idx])])

    edge_to_nodes = [
        filt(0, 1, face_to_nodes[0]),
        filt(0, 2, face_to_nodes[0]),
        filt(0, 3, face_to_nodes[1]),
        filt(1, 2, face_to_nodes[3]),
        filt(1, 3, face_to_nodes[3]),
        filt(2, 3, face_to_nodes[3]),
    ]
    nodes_to_nodes = [0, 1, 2, 3]
    fi = [list(f) for f in face_to_nodes]
    for f, es in ((0, (0, 1, 3)), (1, (0, 2, 4)), (2, (1, 2, 5)), (3, (3, 4, 5))):
        for e in es:
            fi[f] = left_minus_right(fi[f], edge_to_nodes[e])
    ei = [left_minus_right(e, nodes_to_nodes) for e in edge_to_nodes]
    return ReferenceNumbering(face_to_nodes, fi, edge_to_nodes, ei, nodes_to_nodes)


def interpolation_operator(mesh, graph=None):
    """src/interpolation.jl:7-50 -- P = [I; 1/2 1/2] of size (Nn+Ne) x Nn (CSC)."""
    if graph is None:
        graph = edge_graph(mesh)
    Nn = mesh.nnodes
    Ne = graph.nedges
    rows = np.concatenate([np.arange(Nn), np.repeat(np.arange(Nn, Nn + Ne), 2)])
    cols = np.concatenate([np.arange(Nn), np.stack([graph.frm, graph.adj], axis=1).ravel()])
    vals = np.concatenate([np.ones(Nn), np.full(2 * Ne, 0.5)])
    P = sp.csc_matrix((vals, (rows, cols)), shape=(Nn + Ne, Nn))
    P.sort_indices()
    return P


class MultilevelReference:
    """src/multilevel_reference.jl:32-36."""

    def __init__(self, levels, numbering, interops):
        self.levels = levels
        self.numbering = numbering
        self.interops = interops

    @property
    def dim(self):
        return self.levels[0].dim


def refined_element(n, dim):
    """src/multilevel_reference.jl:41-61."""
    levels = [reference_element(dim)]
    numbering = [get_local_numbering(levels[0])]
    interops = []
    for i in range(n - 1):
        graph = edge_graph(levels[i])
        levels.append(refine_uniformly(levels[i], graph))
        numbering.append(get_local_numbering(levels[i + 1]))
        interops.append(interpolation_operator(levels[i], graph))
    for mesh in levels:
        mesh.elements = sort_element_nodes(mesh.elements)
    return MultilevelReference(levels, numbering, interops)
