"""Application driver pieces needed for the sigma / residual histories.

Oracle (test infrastructure only).  Restates the parts of
src/examples/homogenized_coefficients.jl that surround the V-cycle.  All
randomness is supplied by the caller (the reference uses Julia's unseeded global
RNG, :246 and :485-488), so that the same arrays can be fed to the CUDA path.
"""
import math
import numpy as np

from .mesh import Mesh, hypercube
from .fem import (Geometry, assemble_matrix, partial_derivatives_functionals,
                  build_local_diffusion_operators, build_local_mass_matrices)
from .interfaces import list_boundary_nodes_edges_faces, list_interior_nodes
from .implicit import (ImplicitFineGrid, ZeroDirichletConstraint, broadcast_interfaces,
                       apply_constraint, zero_out_all_but_one)
from .operators import L2PlusDivAGrad
from .multigrid import LevelState, BaseLevel, vcycle


def compute_boundary_layer(lam, n):
    """:9"""
    return int(math.floor(4 * (n + 1) * lam ** -0.5))


def compute_box_radius(k, n, eps=0.0):
    """:10"""
    return int(math.floor(2.0 ** (n - k * (0.5 - eps))))


def infnorm(x):
    return np.max(np.abs(x), axis=-1)


def element_centers(mesh):
    """:15 -- mean of the vertex coordinates, summed left to right."""
    p = mesh.nodes[mesh.elements]
    s = p[:, 0, :].copy()
    for i in range(1, p.shape[1]):
        s = s + p[:, i, :]
    return s / p.shape[1]


def order_nodes_and_elements_by_magnitude(mesh):
    """:21-28 -- stable sorts by inf-norm (sortperm is stable; sort! on tuples is MergeSort)."""
    I = np.argsort(infnorm(mesh.nodes), kind="stable")
    J = np.empty_like(I)
    J[I] = np.arange(len(I))
    sorted_mesh = Mesh(mesh.nodes[I], np.sort(J[mesh.elements], axis=1))
    order = np.argsort(infnorm(element_centers(sorted_mesh)), kind="stable")
    sorted_mesh.elements = sorted_mesh.elements[order]
    return sorted_mesh


def find_elements_in_radius(mesh, radius):
    """:34-38 -- length of the element prefix within ``radius`` (elements sorted)."""
    return int(np.searchsorted(infnorm(element_centers(mesh)), radius, side="right"))


def find_nodes_in_radius(mesh, radius):
    """:44-48"""
    return int(np.searchsorted(infnorm(mesh.nodes), radius + 10 * np.finfo(float).eps, side="right"))


def unit_xi(dim):
    """:62-65 -- the default xi = ones / sqrt(dim)."""
    xi = np.ones(dim)
    return xi / np.linalg.norm(xi)


def conductivity_per_element(mesh, sigma_cells, offset):
    """:494-503 -- sigma_cells: (n,)*dim + (dim,) array of per-cell diagonal tensors."""
    idx = np.trunc(element_centers(mesh) + np.asarray(offset)).astype(np.int64) - 1
    return sigma_cells[tuple(idx[:, d] for d in range(mesh.dim))]


def rhs_a_xi_grad_v(b, dphi, implicit, sigmas, xi):
    """:449-474 -- b[i, e] = dot(dphi[i], -|J| Jinv' (sigma_e .* xi)); un-summed."""
    g = Geometry(implicit.base)
    P = -g.det[:, None] * np.einsum("eji,ej->ei", g.inv_jac, sigmas * xi[None, :])
    b[:, :] = dphi @ P.T
    return b


def integrate_area(ops, implicit, nsubset):
    """:673-689"""
    g = Geometry(implicit.base, implicit.base.elements[:nsubset])
    m_total = ops.mass.sum()
    area = 0.0
    for d in g.det:
        area += m_total * d
    return area


def integrate_first_term(v0, dphi, implicit, nsubset, ops, sigmas, xi):
    """:592-632"""
    g = Geometry(implicit.base, implicit.base.elements[:nsubset])
    P = -g.det[:, None] * np.einsum("eji,ej->ei", g.inv_jac, sigmas[:nsubset] * xi[None, :])
    v = v0[:, :nsubset]
    Mv = ops.mass @ v
    running = np.einsum("ie,ie->e", v, dphi @ P.T + Mv)
    total = 0.0
    for e in range(nsubset):
        total += running[e] * g.det[e]
    return total


def integrate_terms(vk, vkm1, implicit, nsubset, ops):
    """:634-667"""
    g = Geometry(implicit.base, implicit.base.elements[:nsubset])
    v = vk[:, :nsubset]
    Mv = ops.mass @ v
    running = np.einsum("ie,ie->e", v + vkm1[:, :nsubset], Mv)
    total = 0.0
    for e in range(nsubset):
        total += running[e] * g.det[e]
    return total


def next_rhs(b, x, implicit, ops):
    """:695-713 -- b = lambda |J| M x (local)."""
    g = Geometry(implicit.base)
    b[:, :] = 0.0
    b += ops.mass @ (x * (ops.lam * g.det)[None, :])
    return b


def make_base(dim, n):
    """:185-209 -- the initial domain of checkerboard_homogenization(n, ...)."""
    lam = 1.0
    box_radius = compute_box_radius(0, n)
    boundary_layer = compute_boundary_layer(lam, n)
    total_radius = box_radius + boundary_layer
    base = order_nodes_and_elements_by_magnitude(
        hypercube(dim, 2 * total_radius, origin=(-float(total_radius),) * dim))
    return base, total_radius


def checkerboard_homogenization(n, dim, refinements=2, smoothing_steps=3, tolerance=1e-4,
                                xi=None, sigma_cells=None, x0=None, max_cycles=1000, log=None):
    """:174-343.  ``sigma_cells``: per-unit-cell diagonal conductivities, shape
    (2R,)*dim + (dim,); ``x0``: (Nf, Ne) initial guess before interface-sum/constraint.
    Returns (sigma, history) with history = list of per-outer-step lists of
    (residual_norm, sigma + dsigma, |dsigma - dsigma_prev|) -- the @info of :287."""
    if xi is None:
        xi = unit_xi(dim)
    lam = 1.0
    sigma = 0.0
    box_radius = compute_box_radius(0, n)
    boundary_layer = compute_boundary_layer(lam, n)
    total_radius = box_radius + boundary_layer
    base, _ = make_base(dim, n)
    cond = conductivity_per_element(base, sigma_cells, (total_radius + 1.0,) * dim)
    total_grids = refinements + 1
    implicit = ImplicitFineGrid(base, total_grids)
    constraint = ZeroDirichletConstraint(*list_boundary_nodes_edges_faces(base))
    diff_terms = build_local_diffusion_operators(implicit.reference)
    mass_terms = build_local_mass_matrices(implicit.reference)
    level_operators = [L2PlusDivAGrad(d, m, constraint, lam, cond) for d, m in zip(diff_terms, mass_terms)]
    level_states = [LevelState(implicit, i + 1) for i in range(total_grids)]
    top = level_states[-1]
    top.x[:, :] = x0
    broadcast_interfaces(top.x, implicit, total_grids)
    apply_constraint(top.x, total_grids, constraint, implicit)
    dphi = partial_derivatives_functionals(implicit.refined_mesh(total_grids))
    rhs_a_xi_grad_v(top.b, dphi, implicit, cond, xi)
    v_prev = None
    history = []
    for k in range(n + 1):
        interior = list_interior_nodes(base)
        A = assemble_matrix(base, sigma=cond, lam=lam)
        base_level = BaseLevel(A[interior][:, interior], base.nnodes, interior)
        dsigma = 0.0
        dsigma_prev = 0.0
        hist = []
        for i in range(max_cycles):
            vcycle(implicit, base_level, level_operators, level_states, total_grids, smoothing_steps)
            nsub = find_elements_in_radius(base, box_radius)
            area = integrate_area(level_operators[-1], implicit, nsub)
            if k == 0:
                integral = integrate_first_term(top.x, dphi, implicit, nsub, level_operators[-1], cond, xi)
            else:
                integral = integrate_terms(top.x, v_prev, implicit, nsub, level_operators[-1])
            dsigma = 2.0 ** k * integral / area
            zero_out_all_but_one(top.r, implicit, total_grids)
            rn = float(np.linalg.norm(top.r.ravel(order="K")))
            hist.append((rn, sigma + dsigma, abs(dsigma - dsigma_prev)))
            if log:
                log(k, i + 1, *hist[-1])
            if abs(dsigma - dsigma_prev) < tolerance:
                break
            dsigma_prev = dsigma
        history.append(hist)
        sigma += dsigma
        lam /= 2
        box_radius = compute_box_radius(k + 1, n)
        boundary_layer = compute_boundary_layer(lam, n)
        if box_radius + boundary_layer > total_radius:
            break
        total_radius = box_radius + boundary_layer
        nn = find_nodes_in_radius(base, total_radius)
        ne = find_elements_in_radius(base, total_radius)
        base = Mesh(base.nodes[:nn], base.elements[:ne])
        cond = cond[:ne]          # level operators keep the full vector; only the prefix is read
        constraint = ZeroDirichletConstraint(*list_boundary_nodes_edges_faces(base))
        for st in level_states:
            for name in ("x", "b", "r", "p", "Ap"):
                setattr(st, name, np.asfortranarray(getattr(st, name)[:, :ne]))
        top = level_states[-1]
        implicit = ImplicitFineGrid(base, total_grids)
        apply_constraint(top.x, total_grids, constraint, implicit)
        v_prev = top.x.copy(order="F")
        for op in level_operators:
            op.lam = lam
            op.constraint = constraint
            op.sigmas = cond
        next_rhs(top.b, top.x, implicit, level_operators[-1])
    return sigma, history
