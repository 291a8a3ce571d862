"""ctypes wrapper of oracle/c/hmg_ref_cpu.c -- the threaded CPU restatement of the reference.

TEST INFRASTRUCTURE / CPU BASELINE ONLY ("restatement, not Julia").  The loop structure is the
reference's (dim^2+1 CSC scatter-SpMVs per element with cyclic thread distribution, serial
interface sums, unfused BLAS-1 passes); the orchestration below follows src/multigrid.jl:46-119
line by line.  Validated against the numpy oracle in tests/test_oracle_cref.py.
"""
import ctypes as C
import os

import numpy as np

from .fem import Geometry
from .operators import element_tensors
from .implicit import copy_to_base, distribute

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "c", "libhmg_ref_cpu.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            raise ImportError(f"{_LIB} missing: run `make -C oracle/c`")
        _lib = C.CDLL(_LIB)
        _lib.ref_dot.restype = C.c_double
        _lib.ref_max_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class _Csc:
    def __init__(self, A):
        A = A.tocsc()
        A.sort_indices()
        self.colptr = np.ascontiguousarray(A.indptr, dtype=np.int64)
        self.rowval = np.ascontiguousarray(A.indices, dtype=np.int64)
        self.nzval = np.ascontiguousarray(A.data, dtype=np.float64)
        self.shape = A.shape


class CpuReference:
    """The hot path on the CPU with ``nthreads`` threads (JULIA_NUM_THREADS analogue)."""

    def __init__(self, implicit, ops, nthreads=None):
        self.lib = lib()
        self.implicit = implicit
        self.ops = ops
        self.nthreads = int(nthreads or os.cpu_count())
        dim = implicit.dim
        self.dim = dim
        self.level_ops = []
        for op in ops:
            mats = [_Csc(op.diffusion_terms[k][l]) for k in range(dim) for l in range(dim)]
            arr = lambda field: (C.c_void_p * (dim * dim))(*[_p(getattr(m, field)) for m in mats])
            self.level_ops.append(dict(mats=mats, colptr=arr("colptr"), rowval=arr("rowval"), nzval=arr("nzval"),
                                       mass=_Csc(op.mass)))
        self.interops = [_Csc(P) for P in implicit.reference.interops]
        self.refresh_geometry()
        self._cells = {}

    def refresh_geometry(self):
        P, det = element_tensors(self.implicit.base, self.ops[-1])
        self.P = np.ascontiguousarray(P)
        self.det = np.ascontiguousarray(det)

    # -- primitives --------------------------------------------------------------------------
    def mul(self, alpha, level, x, y):
        lo = self.level_ops[level - 1]
        nf, ne = x.shape
        self.lib.ref_mul(C.c_double(alpha), self.dim, C.c_int64(ne), C.c_int64(nf), _p(self.P), _p(self.det),
                         C.c_double(self.ops[level - 1].lam), lo["colptr"], lo["rowval"], lo["nzval"],
                         _p(lo["mass"].colptr), _p(lo["mass"].rowval), _p(lo["mass"].nzval), _p(x), _p(y),
                         self.nthreads)
        return y

    def _families(self, level, maps, tag):
        key = (tag, level, id(maps))
        if key not in self._cells:
            nb = self.implicit.local_numbering(level)
            fams = []
            if self.dim == 3:
                fams.append((maps.faces, nb.faces_interior))
            fams.append((maps.edges, nb.edges_interior))
            fams.append((maps.nodes, [[n] for n in nb.nodes]))
            out = []
            for smap, lists in fams:
                if smap.ncells == 0 or len(lists[0]) == 0:
                    continue
                rows = np.ascontiguousarray(np.asarray(lists, dtype=np.int64))
                out.append((smap, rows, np.zeros(rows.shape[1])))
            self._cells[key] = (out, maps)
        return self._cells[key][0]

    def _cells_op(self, mode, x, level, maps, tag):
        nf = x.shape[0]
        for smap, rows, buf in self._families(level, maps, tag):
            self.lib.ref_cells(mode, C.c_int64(smap.ncells), _p(smap.offset), _p(smap.element), _p(smap.local_id),
                               C.c_int64(rows.shape[1]), _p(rows), C.c_int64(nf), _p(x), _p(buf))
        return x

    def broadcast_interfaces(self, x, level):
        return self._cells_op(0, x, level, self.implicit.interfaces, "i")

    def apply_constraint(self, x, level, z):
        return self._cells_op(1, x, level, z, "z")

    def zero_out_all_but_one(self, x, level):
        return self._cells_op(2, x, level, self.implicit.interfaces, "i")

    def dot(self, a, b):
        return float(self.lib.ref_dot(C.c_int64(a.size), _p(a), _p(b), self.nthreads))

    def axpy(self, alpha, x, y):
        self.lib.ref_axpy(C.c_int64(x.size), C.c_double(alpha), _p(x), _p(y), self.nthreads)

    def copy(self, dst, src):
        self.lib.ref_copy(C.c_int64(src.size), _p(src), _p(dst), self.nthreads)

    def fill(self, dst, v):
        self.lib.ref_fill(C.c_int64(dst.size), C.c_double(v), _p(dst), self.nthreads)

    # -- src/apply_local_operators.jl:18-27, src/multigrid.jl:46-119 ---------------------------
    def local_residual(self, curr, k):
        self.copy(curr.r, curr.b)
        self.mul(-1.0, k, curr.x, curr.r)
        self.apply_constraint(curr.r, k, self.ops[k - 1].constraint)

    def global_product(self, curr, k):
        """Ap = broadcast(constraint(A p)) -- the benchmarked A*x (src/multigrid.jl:58-61)."""
        self.fill(curr.Ap, 0.0)
        self.mul(1.0, k, curr.p, curr.Ap)
        self.apply_constraint(curr.Ap, k, self.ops[k - 1].constraint)
        self.broadcast_interfaces(curr.Ap, k)

    def smoothing_steps(self, steps, curr, k):
        self.local_residual(curr, k)
        self.broadcast_interfaces(curr.r, k)
        self.copy(curr.p, curr.r)
        rsqrprev = self.dot(curr.r, curr.r)
        for _ in range(steps):
            self.global_product(curr, k)
            alpha = rsqrprev / self.dot(curr.p, curr.Ap)
            self.axpy(alpha, curr.p, curr.x)
            self.axpy(-alpha, curr.Ap, curr.r)
            rsqr = self.dot(curr.r, curr.r)
            self.lib.ref_xpby(C.c_int64(curr.p.size), _p(curr.r), C.c_double(rsqr / rsqrprev), _p(curr.p))
            rsqrprev = rsqr

    def vcycle(self, base, levels, k, steps=2):
        if k == 1:
            l1 = levels[0]
            self.broadcast_interfaces(l1.b, 1)
            copy_to_base(base.b, l1.b, self.implicit)
            base.b_interior[:] = base.b[base.interior_nodes]
            tmp = base.solve(base.b_interior)
            base.b[:] = 0.0
            base.b[base.interior_nodes] = tmp
            distribute(l1.x, base.b, self.implicit)
            return
        curr, nxt = levels[k - 1], levels[k - 2]
        P = self.interops[k - 2]
        ne = curr.x.shape[1]
        self.smoothing_steps(steps, curr, k)
        self.local_residual(curr, k)
        self.lib.ref_restrict(C.c_int64(ne), C.c_int64(P.shape[0]), C.c_int64(P.shape[1]), _p(P.colptr), _p(P.rowval),
                              _p(P.nzval), _p(curr.r), _p(nxt.b), self.nthreads)
        self.fill(nxt.x, 0.0)
        self.vcycle(base, levels, k - 1)
        self.lib.ref_interpolate(C.c_int64(ne), C.c_int64(P.shape[0]), C.c_int64(P.shape[1]), _p(P.colptr),
                                 _p(P.rowval), _p(P.nzval), _p(nxt.x), _p(curr.x), self.nthreads)
        self.smoothing_steps(steps, curr, k)
