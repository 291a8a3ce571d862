"""Host-side mirror of the reference's Julia API for the hot path, over the C ABI.

Names, argument meaning and error behaviour follow the reference (file:line relative to the
reference repository root); matrices live on the device and are addressed through
``DeviceMatrix`` handles, the analogue of a ``LevelState{Float64,DeviceMatrix}`` in Julia
(src/multigrid.jl:7).  Indices are 0-based on this side and converted to the ABI's 1-based
Int64 at the boundary.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import VEC_IDS, check


class Mesh:
    """src/grid.jl:19-22 -- nodes (Nn, dim) float64, elements (Ne, dim+1) int64, 0-based."""

    def __init__(self, nodes, elements):
        self.nodes = np.ascontiguousarray(nodes, dtype=np.float64)
        self.elements = np.ascontiguousarray(elements, dtype=np.int64)
        if self.nodes.ndim != 2 or self.elements.ndim != 2 or self.elements.shape[1] != self.nodes.shape[1] + 1:
            raise ValueError("Mesh: nodes must be (Nn, dim) and elements (Ne, dim+1)")

    @property
    def dim(self):
        return self.nodes.shape[1]

    @property
    def nnodes(self):
        return self.nodes.shape[0]

    @property
    def nelements(self):
        return self.elements.shape[0]


class DeviceMatrix:
    """An Nf(level) x Ne state matrix resident on the GPU (reference layout on upload/download)."""

    def __init__(self, implicit, level, name):
        self.implicit = implicit
        self.level = level
        self.name = name
        self.which = VEC_IDS[name]

    @property
    def shape(self):
        return (self.implicit.nf(self.level), self.implicit.ne_local)

    def set(self, host):
        """copyto!(device, host): host is (Nf, Ne_local), hierarchical row order."""
        host = np.asfortranarray(host, dtype=np.float64)
        if host.shape != self.shape:
            raise ValueError(f"expected shape {self.shape}, got {host.shape}")
        check(self.implicit.lib.hmg_upload(self.implicit.ctx, self.level, self.which,
                                           host.ctypes.data_as(C.c_void_p), host.shape[0]))
        return self

    def get(self, out=None):
        """Array(device): returns an (Nf, Ne_local) Fortran-ordered host array."""
        if out is None:
            out = np.empty(self.shape, dtype=np.float64, order="F")
        assert out.flags.f_contiguous and out.shape == self.shape
        check(self.implicit.lib.hmg_download(self.implicit.ctx, self.level, self.which,
                                             out.ctypes.data_as(C.c_void_p), out.shape[0]))
        return out

    def get_rows(self, nrows):
        """Array(device[1:nrows, :]): the first ``nrows`` hierarchical rows of every column -- with
        nrows = Nf(k) the values on the nodes of the coarser level k, the slice export_unknown takes
        (src/examples/homogenized_coefficients.jl:84).  Only nrows x Ne_local doubles are transferred."""
        nrows = int(nrows)
        if not 0 <= nrows <= self.shape[0]:
            raise ValueError(f"nrows must be in 0..{self.shape[0]}")
        out = np.empty((nrows, self.shape[1]), dtype=np.float64, order="F")
        check(self.implicit.lib.hmg_download_rows(self.implicit.ctx, self.level, self.which, nrows,
                                                  out.ctypes.data_as(C.c_void_p), max(nrows, 1)))
        return out

    def copy_columns_from(self, other):
        """self = other[:, OneTo(Ne_local(self))] on the device, across contexts: shrink_level_state
        (src/examples/homogenized_coefficients.jl:54-60) without a host round trip.  ``self`` lives in a
        context created on an element prefix of ``other``'s base mesh."""
        if other.level != self.level:
            raise ValueError("copy_columns_from: levels differ")
        check(self.implicit.lib.hmg_copy_columns_from(self.implicit.ctx, self.level, self.which,
                                                      other.implicit.ctx, other.which))
        return self

    def fill(self, value):
        check(self.implicit.lib.hmg_fill(self.implicit.ctx, self.level, self.which, float(value)))
        return self

    def copy_from(self, other):
        assert other.level == self.level
        check(self.implicit.lib.hmg_copy(self.implicit.ctx, self.level, self.which, other.which))
        return self


class LevelState:
    """src/multigrid.jl:7-25 -- x, b, r, p, Ap of one level (plus the scratch v, w)."""

    def __init__(self, implicit, level):
        for name in VEC_IDS:
            setattr(self, name, DeviceMatrix(implicit, level, name))


class ImplicitFineGrid:
    """ImplicitFineGrid(base, levels) (src/implicit_fine_grid.jl:6-18) together with the zero
    Dirichlet constraint of the base mesh (src/interface.jl:207-284), the operator
    L2PlusDivAGrad(diff, mass, constraint, lambda, sigma) on every level
    (src/build_local_operators.jl:26-32) and the level states (src/multigrid.jl:18-25) -- all of
    which the library builds and keeps on the device."""

    def __init__(self, base, levels, sigma, lam=1.0, device=0, owner_rank=None, rank=0, nranks=1, nccl_id=None):
        self.lib = _lib.load()
        self.base = base
        self.levels = int(levels)
        self.dim = base.dim
        sigma = np.ascontiguousarray(sigma, dtype=np.float64)
        if sigma.shape != (base.nelements, base.dim):
            raise ValueError("sigma must be (Ne, dim)")
        if not np.all(np.diff(base.elements, axis=1) > 0):
            raise AssertionError("base elements must be sorted (src/implicit_fine_grid.jl:14)")
        elems1 = np.ascontiguousarray(base.elements + 1, dtype=np.int64)
        ctx = C.c_void_p()
        if owner_rank is None:
            check(self.lib.hmg_create(base.dim, self.levels, base.nelements, base.nnodes,
                                      base.nodes.ctypes.data_as(C.c_void_p), elems1.ctypes.data_as(C.c_void_p),
                                      sigma.ctypes.data_as(C.c_void_p), float(lam), int(device), C.byref(ctx)))
        else:
            owner_rank = np.ascontiguousarray(owner_rank, dtype=np.int32)
            check(self.lib.hmg_create_partitioned(
                base.dim, self.levels, base.nelements, base.nnodes,
                base.nodes.ctypes.data_as(C.c_void_p), elems1.ctypes.data_as(C.c_void_p),
                sigma.ctypes.data_as(C.c_void_p), float(lam), int(device), int(rank), int(nranks),
                owner_rank.ctypes.data_as(C.c_void_p), nccl_id, C.byref(ctx)))
        self.ctx = ctx
        self.lam = float(lam)
        self.ne_local = int(self.lib.hmg_ne_local(self.ctx))
        self.states = [LevelState(self, l) for l in range(1, self.levels + 1)]

    @classmethod
    def simple_diffusion(cls, base, levels, a=1.0, **kw):
        """The grid of SimpleDiffusion(ops, bc, a), L = -a * Laplacian with the zero Dirichlet constraint
        (src/build_local_operators.jl:19-23, product src/apply_local_operators.jl:40-72: P = Jinv' * Jinv scaled by
        a * |J|): the same operator as L2PlusDivAGrad with sigma = (a, ..., a) and lambda = 0, whose mass term the
        kernels skip exactly like the reference does (src/apply_local_operators.jl:116)."""
        sigma = np.full((base.nelements, base.dim), float(a))
        return cls(base, levels, sigma, lam=0.0, **kw)

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.hmg_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def nf(self, level):
        return int(self.lib.hmg_nf(self.ctx, level))

    def ld(self, level):
        return int(self.lib.hmg_ld(self.ctx, level))

    def state(self, level):
        """1-based level, as in the reference."""
        return self.states[level - 1]

    def refined_mesh(self, level):
        """refined_mesh(implicit, level) (src/implicit_fine_grid.jl:24): the refined reference element, host data."""
        from .vtk import refined_mesh
        return refined_mesh(self.dim, self.levels, level)

    def construct_full_grid(self, level):
        """construct_full_grid(implicit, level) (src/implicit_fine_grid.jl:41-78), host data."""
        from .vtk import construct_full_grid
        return construct_full_grid(self.base, self.levels, level)

    def local_elements(self):
        out = np.empty(self.ne_local, dtype=np.int64)
        check(self.lib.hmg_local_elements(self.ctx, out.ctypes.data_as(C.c_void_p)))
        return out - 1

    def comm_mode(self):
        """'single', 'nccl' or 'peer' (hmg_comm_mode): how the ranks of a partitioned context exchange data."""
        return {0: "single", 1: "nccl", 2: "peer"}[int(self.lib.hmg_comm_mode(self.ctx))]

    def set_lambda(self, lam):
        """operator.λ = λ (src/examples/homogenized_coefficients.jl:331)."""
        check(self.lib.hmg_set_lambda(self.ctx, float(lam)))
        self.lam = float(lam)

    def set_sigma(self, sigma):
        sigma = np.ascontiguousarray(sigma, dtype=np.float64)
        check(self.lib.hmg_set_sigma(self.ctx, sigma.ctypes.data_as(C.c_void_p)))

    def hier_to_lattice(self, level):
        out = np.empty(self.nf(level), dtype=np.int32)
        check(self.lib.hmg_hier_to_lattice(self.ctx, level, out.ctypes.data_as(C.c_void_p)))
        return out

    def synchronize(self):
        check(self.lib.hmg_synchronize(self.ctx))

    def launch_count(self):
        return int(self.lib.hmg_launch_count(self.ctx))

    def time_op(self, op, level, steps=3, reps=1):
        ms = C.c_float()
        check(self.lib.hmg_time_op(self.ctx, int(op), int(level), int(steps), int(reps), C.byref(ms)))
        return float(ms.value)


# ---- the reference's generic functions, device methods ---------------------------------------
def mul(alpha, implicit, x, y):
    """mul!(α, base, A, x, y): y ← α A x + y, column-local (src/apply_local_operators.jl:85-120)."""
    check(implicit.lib.hmg_mul(implicit.ctx, x.level, float(alpha), x.which, y.which))
    return y


def apply_global(implicit, x, y):
    """The fused global product of src/multigrid.jl:58-61: y = broadcast(constraint(A x))."""
    check(implicit.lib.hmg_apply_global(implicit.ctx, x.level, x.which, y.which))
    return y


def apply_constraint(x, level, implicit):
    """apply_constraint!(x, level, z, implicit) (src/implicit_fine_grid.jl:94-139)."""
    check(implicit.lib.hmg_apply_constraint(implicit.ctx, level, x.which))
    return x


def broadcast_interfaces(x, implicit, level):
    """broadcast_interfaces!(x, implicit, level) (src/implicit_fine_grid.jl:209-328)."""
    check(implicit.lib.hmg_broadcast_interfaces(implicit.ctx, level, x.which))
    return x


def zero_out_all_but_one(x, implicit, level):
    """zero_out_all_but_one!(x, implicit, level) (src/implicit_fine_grid.jl:334-386)."""
    check(implicit.lib.hmg_zero_out_all_but_one(implicit.ctx, level, x.which))
    return x


def local_residual(implicit, k):
    """local_residual!(implicit, A, curr, k) (src/apply_local_operators.jl:18-27)."""
    check(implicit.lib.hmg_local_residual(implicit.ctx, k))


def restrict_to(implicit, k):
    """restrict_to!(levels[k-1].b, P, levels[k].r) (src/interpolation.jl:64-74)."""
    check(implicit.lib.hmg_restrict(implicit.ctx, k))


def interpolate_and_sum_to(implicit, k):
    """interpolate_and_sum_to!(levels[k].x, P, levels[k-1].x) (src/interpolation.jl:52-62)."""
    check(implicit.lib.hmg_interpolate_add(implicit.ctx, k))


def smoothing_steps(steps, implicit, k):
    """smoothing_steps!(steps, implicit, ops, curr, k) (src/multigrid.jl:46-71)."""
    check(implicit.lib.hmg_smoothing_steps(implicit.ctx, k, int(steps)))


def dot(implicit, a, b):
    """dot(a, b) over all stored entries (src/multigrid.jl:54)."""
    out = C.c_double()
    check(implicit.lib.hmg_dot(implicit.ctx, a.level, a.which, b.which, C.byref(out)))
    return float(out.value)


def axpy(implicit, alpha, x, y):
    check(implicit.lib.hmg_axpy(implicit.ctx, x.level, float(alpha), x.which, y.which))
    return y


def copy_to_base(implicit, v):
    """copy_to_base!(u, v, implicit) (src/implicit_fine_grid.jl:148-171); returns u."""
    u = np.zeros(implicit.base.nnodes)
    check(implicit.lib.hmg_copy_to_base(implicit.ctx, v.which, u.ctypes.data_as(C.c_void_p)))
    return u


def distribute(implicit, v, u):
    """distribute!(v, u, implicit) (src/implicit_fine_grid.jl:178-202)."""
    u = np.ascontiguousarray(u, dtype=np.float64)
    check(implicit.lib.hmg_distribute(implicit.ctx, v.which, u.ctypes.data_as(C.c_void_p)))
    return v


def local_rhs_host(base, levels):
    """The matrix local_rhs!(b, implicit) fills (src/implicit_fine_grid.jl:391-409): b[:, e] = b_ref * |det J_e| with
    b_ref = assemble_vector(refined_mesh(implicit, levels), identity), i.e. the integral of every P1 hat function over
    the refined reference element (each fine element gives volume / (dim + 1) to its vertices).  O(Nf * Ne) host work,
    done once per problem; returns an (Nf, Ne) Fortran-ordered array."""
    from .vtk import refined_mesh
    dim = base.dim
    ref = refined_mesh(dim, levels, levels)
    p = ref.nodes[ref.elements]                                          # (nel, dim+1, dim)
    vol = np.abs(np.linalg.det(p[:, 1:, :] - p[:, :1, :])) / (2.0 if dim == 2 else 6.0)
    b_ref = np.zeros(ref.nnodes)
    np.add.at(b_ref, ref.elements.ravel(), np.repeat(vol / (dim + 1), dim + 1))
    q = base.nodes[base.elements]
    det = np.abs(np.linalg.det(q[:, 1:, :] - q[:, :1, :]))
    return np.asfortranarray(b_ref[:, None] * det[None, :])


def local_rhs(b, implicit):
    """local_rhs!(b, implicit): the un-summed functional of f = 1 on the finest level, computed on the host and uploaded."""
    if b.level != implicit.levels:
        raise ValueError("local_rhs!: b must be a finest-level matrix")
    full = local_rhs_host(implicit.base, implicit.levels)
    if implicit.ne_local != implicit.base.nelements:
        full = np.asfortranarray(full[:, implicit.local_elements()])
    return b.set(full)


# ---- driver functionals on the finest level (src/examples/homogenized_coefficients.jl) -----------
def _xi(implicit, xi):
    xi = np.ascontiguousarray(xi, dtype=np.float64)
    if xi.shape != (implicit.dim,):
        raise ValueError("xi must have dim entries")
    return xi


def rhs_a_xi_grad_v(b, implicit, xi):
    """rhs_aξ∇v!(b, ∂ϕ∂xᵢs, implicit, σs, ξ) (:449-474): b[i, e] = dot(∫∇ϕ_i, -|J| J⁻¹(σ_e .* ξ)), un-summed."""
    xi = _xi(implicit, xi)
    check(implicit.lib.hmg_rhs_axi_grad(implicit.ctx, xi.ctypes.data_as(C.c_void_p), b.which))
    return b


def integrate_first_term(v0, implicit, nsubset, xi):
    """integrate_first_term(v₀, ∂ϕ∂xᵢs, implicit, 1:nsubset, ops, σs, ξ) (:592-632)."""
    xi = _xi(implicit, xi)
    out = C.c_double()
    check(implicit.lib.hmg_integrate_first_term(implicit.ctx, v0.which, xi.ctypes.data_as(C.c_void_p), int(nsubset),
                                                C.byref(out)))
    return float(out.value)


def integrate_terms(vk, vkm1, implicit, nsubset):
    """integrate_terms(vₖ, vₖ₋₁, implicit, 1:nsubset, ops) (:634-667)."""
    out = C.c_double()
    check(implicit.lib.hmg_integrate_terms(implicit.ctx, vk.which, vkm1.which, int(nsubset), C.byref(out)))
    return float(out.value)


def integrate_area(implicit, nsubset):
    """integrate_area(ops, implicit, 1:nsubset) (:673-689)."""
    out = C.c_double()
    check(implicit.lib.hmg_integrate_area(implicit.ctx, int(nsubset), C.byref(out)))
    return float(out.value)


def next_rhs(b, x, implicit):
    """next_rhs!(b, x, implicit, ops) (:695-713): b = λ |J| M x, local."""
    check(implicit.lib.hmg_next_rhs(implicit.ctx, b.which, x.which))
    return b


class BaseLevel:
    """BaseLevel(Float64, F, nnodes, interior) (src/multigrid.jl:30-41).  Instead of a CHOLMOD
    factor the library takes the sparse matrix A[interior, interior] itself (scipy CSC), or
    assembles it from the base mesh when ``A_interior`` is None
    (src/examples/homogenized_coefficients.jl:259-261)."""

    def __init__(self, implicit, A_interior=None, interior_nodes=None):
        self.implicit = implicit
        if A_interior is None:
            check(implicit.lib.hmg_assemble_coarse(implicit.ctx))
        else:
            A = A_interior.tocsc()
            A.sort_indices()
            colptr = np.ascontiguousarray(A.indptr, dtype=np.int64) + 1
            rowval = np.ascontiguousarray(A.indices, dtype=np.int64) + 1
            nzval = np.ascontiguousarray(A.data, dtype=np.float64)
            interior = np.ascontiguousarray(interior_nodes, dtype=np.int64) + 1
            check(implicit.lib.hmg_set_coarse_matrix(
                implicit.ctx, A.shape[0], colptr.ctypes.data_as(C.c_void_p), rowval.ctypes.data_as(C.c_void_p),
                nzval.ctypes.data_as(C.c_void_p), interior.ctypes.data_as(C.c_void_p)))


def vcycle(implicit, base_level, k, steps=2, resnorm=False):
    """vcycle!(implicit, base, ops, levels, k, steps) (src/multigrid.jl:73-119).  With
    ``resnorm`` the logged residual norm(zero_out_all_but_one!(r)) of
    src/examples/homogenized_coefficients.jl:286-287 is returned (r is zeroed in place like there)."""
    assert base_level.implicit is implicit
    if resnorm:
        out = C.c_double()
        check(implicit.lib.hmg_vcycle(implicit.ctx, int(k), int(steps), C.byref(out)))
        return float(out.value)
    check(implicit.lib.hmg_vcycle(implicit.ctx, int(k), int(steps), None))
    return None


def vcycles(implicit, base_level, k, steps, ncycles, resnorms=True):
    """``ncycles`` V-cycles without host synchronisation in between; returns the residual history."""
    out = np.zeros(ncycles) if resnorms else None
    check(implicit.lib.hmg_vcycles(implicit.ctx, int(k), int(steps), int(ncycles),
                                   out.ctypes.data_as(C.c_void_p) if resnorms else None))
    return out
