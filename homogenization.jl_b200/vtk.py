"""VTK export of the base mesh and of a coarse-level slice of the unknown.

Mirrors ``export_domain`` / ``export_unknown`` (src/examples/homogenized_coefficients.jl:71-87), which go
through ``vtk_grid(filename, mesh)`` (src/utils.jl:14-19, WriteVTK.jl with ``compress = 0``) and
``construct_full_grid`` (src/implicit_fine_grid.jl:41-78).  WriteVTK.jl is a third-party package; what is
restated here is the file *format* it produces for these calls: a VTK XML ``UnstructuredGrid`` (``.vtu``)
with Float64 points padded to three components, Int64 connectivity / offsets, UInt8 cell types
(5 = triangle, 10 = tetrahedron) and uncompressed data arrays (written inline, base64, UInt64 block header).

The unknown never crosses PCIe as a whole: ``export_unknown`` downloads only the first Nf(level) rows of every
column (``hmg_download_rows``), which are exactly the nodes of the coarser level (hierarchical row order).
"""
import base64
import ctypes as C
import xml.etree.ElementTree as ET

import numpy as np

from . import _lib
from .api import Mesh

VTK_TRIANGLE, VTK_TETRA = 5, 10
_VTK_TYPES = {np.dtype(np.float64): "Float64", np.dtype(np.int64): "Int64", np.dtype(np.uint8): "UInt8",
              np.dtype(np.int32): "Int32", np.dtype(np.float32): "Float32"}
_NP_TYPES = {v: k for k, v in _VTK_TYPES.items()}


def refined_mesh(dim, levels, level):
    """refined_mesh(implicit, level) (src/implicit_fine_grid.jl:24): the refined reference element of ``level`` out of
    ``levels`` grids -- nodes in hierarchical row order, fine elements index-sorted (src/multilevel_reference.jl:41-61).
    Host-only (no context needed)."""
    lib = _lib.load()
    nel = C.c_int64()
    _lib.check_host(lib.hmg_host_refined_mesh(dim, levels, level, None, None, C.byref(nel)))
    sizes = (C.c_int64 * 8)()
    _lib.check_host(lib.hmg_host_reference(dim, levels, level, sizes, None, None, None))
    nf = int(sizes[1])
    nodes = np.empty((nf, dim), dtype=np.float64)
    elems = np.empty((int(nel.value), dim + 1), dtype=np.int64)
    _lib.check_host(lib.hmg_host_refined_mesh(dim, levels, level, nodes.ctypes.data_as(C.c_void_p),
                                              elems.ctypes.data_as(C.c_void_p), C.byref(nel)))
    return Mesh(nodes, elems - 1)


def construct_full_grid(base, levels, level):
    """construct_full_grid(implicit, level) (src/implicit_fine_grid.jl:41-78): the explicit mesh of ``level`` with the
    interface nodes repeated -- node (e, i) = J_e ref_i + p_1(e), fine elements offset by e * Nf(level)."""
    ref = refined_mesh(base.dim, levels, level)
    p = base.nodes[base.elements]                                   # (Ne, dim+1, dim)
    J = np.transpose(p[:, 1:, :] - p[:, :1, :], (0, 2, 1))          # columns p_k - p_1 (src/grid.jl:120-135)
    nodes = np.einsum("eij,nj->eni", J, ref.nodes) + p[:, :1, :]
    nn = ref.nnodes
    elements = ref.elements[None, :, :] + (np.arange(base.nelements, dtype=np.int64) * nn)[:, None, None]
    return Mesh(nodes.reshape(-1, base.dim), elements.reshape(-1, base.dim + 1))


def _data_array(parent, name, arr, ncomp=None):
    arr = np.ascontiguousarray(arr)
    el = ET.SubElement(parent, "DataArray", type=_VTK_TYPES[arr.dtype], Name=name, format="binary")
    if ncomp is not None:
        el.set("NumberOfComponents", str(ncomp))
    raw = arr.tobytes()
    el.text = base64.b64encode(np.uint64(len(raw)).tobytes() + raw).decode("ascii")
    return el


def write_vtu(filename, mesh, point_data=None, cell_data=None):
    """vtk_grid(filename, mesh) do vtk ... end (src/utils.jl:14-19): writes ``filename`` (``.vtu`` appended when
    missing) and returns the path.  point_data / cell_data: name -> array; vectorial data is (n, ncomp)."""
    path = filename if filename.endswith(".vtu") else filename + ".vtu"
    nv = mesh.dim + 1
    root = ET.Element("VTKFile", type="UnstructuredGrid", version="1.0", byte_order="LittleEndian", header_type="UInt64")
    grid = ET.SubElement(root, "UnstructuredGrid")
    piece = ET.SubElement(grid, "Piece", NumberOfPoints=str(mesh.nnodes), NumberOfCells=str(mesh.nelements))
    pts = np.zeros((mesh.nnodes, 3), dtype=np.float64)
    pts[:, :mesh.dim] = mesh.nodes
    _data_array(ET.SubElement(piece, "Points"), "Points", pts, 3)
    cells = ET.SubElement(piece, "Cells")
    _data_array(cells, "connectivity", mesh.elements.astype(np.int64).reshape(-1))
    _data_array(cells, "offsets", np.arange(1, mesh.nelements + 1, dtype=np.int64) * nv)
    _data_array(cells, "types", np.full(mesh.nelements, VTK_TRIANGLE if mesh.dim == 2 else VTK_TETRA, dtype=np.uint8))
    for tag, data, count in (("PointData", point_data, mesh.nnodes), ("CellData", cell_data, mesh.nelements)):
        sec = ET.SubElement(piece, tag)
        for name, arr in (data or {}).items():
            arr = np.asarray(arr, dtype=np.float64)
            if arr.shape[0] != count:
                raise ValueError(f"{tag} '{name}': expected {count} entries, got {arr.shape[0]}")
            _data_array(sec, name, arr, arr.shape[1] if arr.ndim == 2 else None)
    ET.ElementTree(root).write(path, xml_declaration=True, encoding="utf-8")
    return path


def read_vtu(path):
    """Reads back what ``write_vtu`` wrote: (mesh, point_data, cell_data).  Used by the tests and for round trips."""
    root = ET.parse(path).getroot()
    piece = root.find("UnstructuredGrid/Piece")

    def decode(el):
        raw = base64.b64decode(el.text)
        n = int(np.frombuffer(raw[:8], dtype=np.uint64)[0])
        arr = np.frombuffer(raw[8:8 + n], dtype=_NP_TYPES[el.get("type")]).copy()
        nc = el.get("NumberOfComponents")
        return arr.reshape(-1, int(nc)) if nc else arr

    pts = decode(piece.find("Points/DataArray"))
    arrays = {el.get("Name"): decode(el) for el in piece.findall("Cells/DataArray")}
    vtk_type = int(arrays["types"][0]) if len(arrays["types"]) else VTK_TRIANGLE
    dim = 2 if vtk_type == VTK_TRIANGLE else 3
    mesh = Mesh(pts[:, :dim], arrays["connectivity"].reshape(-1, dim + 1))
    pd = {el.get("Name"): decode(el) for el in piece.findall("PointData/DataArray")}
    cd = {el.get("Name"): decode(el) for el in piece.findall("CellData/DataArray")}
    return mesh, pd, cd


def export_domain(base, cond, filename="checkerboard"):
    """export_domain(base, cond) (src/examples/homogenized_coefficients.jl:71-79): the base mesh with the per-element
    conductivities as cell data "a" (dim components per element)."""
    return write_vtu(filename, base, cell_data={"a": np.asarray(cond, dtype=np.float64)})


def export_unknown(implicit, x, k, level, filename=None):
    """export_unknown(base, implicit, x, k, level) (:81-87): the explicit grid of ``level`` with
    x[1 : nnodes(refined_mesh(implicit, level)), :][:] as point data "v", written to ``ahom_<k>.vtu``.
    ``x`` is a DeviceMatrix of the finest level; only the first Nf(level) rows of every column are downloaded."""
    if not 1 <= level <= implicit.levels:
        raise ValueError("level out of range")
    base = implicit.base
    if implicit.ne_local != base.nelements:                            # partitioned context: this rank's columns
        base = Mesh(base.nodes, base.elements[implicit.local_elements()])
    full = construct_full_grid(base, implicit.levels, level)
    vals = x.get_rows(implicit.nf(level))                              # (Nf(level), Ne), column-major
    return write_vtu(filename or f"ahom_{k}", full, point_data={"v": vals.reshape(-1, order="F")})
