"""P1 element values, quadrature, explicit assembly, local operator construction.

Oracle (test infrastructure only).  Restates src/cell_values.jl, src/assembly.jl,
src/build_local_operators.jl and the assembly helpers of
src/examples/homogenized_coefficients.jl.  Vectorised over elements; the
quadrature loop is kept sequential so that sums are formed in the reference's
order.
"""
import numpy as np
import scipy.sparse as sp


def quad_rule(dim):
    """src/cell_values.jl:10-37 -- TriQuad3 / TetQuad4 (default_quad)."""
    if dim == 2:
        pts = np.array([(0.0, 0.5), (0.5, 0.0), (0.5, 0.5)])
        w = np.array([1 / 6, 1 / 6, 1 / 6])
    else:
        s5 = np.sqrt(5.0)
        a, b = (5 + 3 * s5) / 20, (5 - s5) / 20
        pts = np.array([(a, b, b), (b, a, b), (b, b, a), (b, b, b)])
        w = np.array([1 / 24] * 4)
    return pts, w


def basis_values(dim, pts):
    """src/cell_values.jl:40-51 -- phi_i(x_q); shape (nquad, N)."""
    first = 1.0 - pts.sum(axis=1, keepdims=True)
    return np.concatenate([first, pts], axis=1)


def ref_gradients(dim):
    """src/cell_values.jl:86 -- constant gradients of the P1 basis; shape (dim, N)."""
    return np.concatenate([-np.ones((dim, 1)), np.eye(dim)], axis=1)


class Geometry:
    """Per-element J, inv(J'), |det J| -- reinit!, src/cell_values.jl:104-127."""

    def __init__(self, mesh, elements=None):
        el = mesh.elements if elements is None else elements
        p = mesh.nodes[el]                                   # (Ne, N, dim)
        self.J = np.transpose(p[:, 1:, :] - p[:, :1, :], (0, 2, 1))   # columns p_k - p_1
        self.inv_jac = np.linalg.inv(np.transpose(self.J, (0, 2, 1)))  # inv(J')
        self.det = np.abs(np.linalg.det(self.J))
        self.shift = p[:, 0, :]
        dim = mesh.dim
        self.gradients = self.inv_jac @ ref_gradients(dim)    # (Ne, dim, N)


def _coo_to_csc(rows, cols, vals, n, dropzeros=True):
    A = sp.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsc()
    A.sum_duplicates()
    if dropzeros:
        A.eliminate_zeros()
    A.sort_indices()
    return A


def build_local_diffusion_operators_level(mesh):
    """src/build_local_operators.jl:51-105 -- ops[k][l][i,j] = int d_k phi_i d_l phi_j."""
    dim = mesh.dim
    N = dim + 1
    _, w = quad_rule(dim)
    g = Geometry(mesh)
    el = mesh.elements
    rows = np.repeat(el, N, axis=1).ravel()          # element[i] for (i, j) with j inner
    cols = np.tile(el, (1, N)).ravel()               # element[j]
    ops = [[None] * dim for _ in range(dim)]
    for k in range(dim):
        for l in range(dim):
            # A_locals[l,k][i,j] += w[qp] * grad_i[k] * grad_j[l]
            loc = np.zeros((el.shape[0], N, N))
            for q in range(len(w)):
                loc += w[q] * g.gradients[:, k, :, None] * g.gradients[:, l, None, :]
            vals = (loc * g.det[:, None, None]).reshape(-1)
            ops[k][l] = _coo_to_csc(rows, cols, vals, mesh.nnodes)
    return ops


def mass_matrix(mesh):
    """src/build_local_operators.jl:107-141."""
    dim = mesh.dim
    N = dim + 1
    pts, w = quad_rule(dim)
    phi = basis_values(dim, pts)
    g = Geometry(mesh)
    el = mesh.elements
    loc = np.zeros((N, N))
    for q in range(len(w)):
        loc += w[q] * np.outer(phi[q], phi[q])
    vals = (loc[None, :, :] * g.det[:, None, None]).reshape(-1)
    rows = np.repeat(el, N, axis=1).ravel()
    cols = np.tile(el, (1, N)).ravel()
    return _coo_to_csc(rows, cols, vals, mesh.nnodes)


def build_local_diffusion_operators(ref):
    """src/build_local_operators.jl:39-43."""
    return [build_local_diffusion_operators_level(level) for level in ref.levels]


def build_local_mass_matrices(ref):
    """src/build_local_operators.jl:45-49."""
    return [mass_matrix(level) for level in ref.levels]


def assemble_matrix(mesh, sigma=None, lam=0.0):
    """src/assembly.jl:4-60 with bf = dot (sigma None), or
    src/examples/homogenized_coefficients.jl:358-402 (assemble_checkerboard)."""
    dim = mesh.dim
    N = dim + 1
    pts, w = quad_rule(dim)
    phi = basis_values(dim, pts)
    g = Geometry(mesh)
    el = mesh.elements
    grad = g.gradients                                # (Ne, dim, N)
    if sigma is None:
        stiff = np.einsum("edi,edj->eij", grad, grad)
    else:
        stiff = np.einsum("edi,ed,edj->eij", grad, np.asarray(sigma), grad)
    loc = np.zeros((el.shape[0], N, N))
    for q in range(len(w)):
        loc += w[q] * (lam * np.outer(phi[q], phi[q])[None, :, :] + stiff)
    vals = (loc * g.det[:, None, None]).reshape(-1)
    rows = np.repeat(el, N, axis=1).ravel()
    cols = np.tile(el, (1, N)).ravel()
    return _coo_to_csc(rows, cols, vals, mesh.nnodes, dropzeros=False)


def assemble_vector(mesh):
    """src/assembly.jl:121-154 with functional = identity."""
    dim = mesh.dim
    pts, w = quad_rule(dim)
    phi = basis_values(dim, pts)
    g = Geometry(mesh)
    loc = np.zeros(dim + 1)
    for q in range(len(w)):
        loc += w[q] * phi[q]
    b = np.zeros(mesh.nnodes)
    np.add.at(b, mesh.elements.ravel(), (loc[None, :] * g.det[:, None]).ravel())
    return b


def partial_derivatives_functionals(mesh):
    """src/examples/homogenized_coefficients.jl:407-442 -- int d phi_i / d x_j; (Nn, dim)."""
    dim = mesh.dim
    _, w = quad_rule(dim)
    g = Geometry(mesh)
    loc = np.zeros_like(g.gradients)                  # (Ne, dim, N)
    for q in range(len(w)):
        loc += w[q] * g.gradients
    loc = loc * g.det[:, None, None]
    bs = np.zeros((mesh.nnodes, dim))
    # element-major, i inner: the reference's accumulation order (:418-438)
    np.add.at(bs, mesh.elements.ravel(), np.transpose(loc, (0, 2, 1)).reshape(-1, dim))
    return bs
