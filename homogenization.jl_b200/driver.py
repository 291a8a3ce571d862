"""Host-side mirror of ``checkerboard_homogenization`` (src/examples/homogenized_coefficients.jl:174-343)
-- the caller of the hot path -- over the device API: every V-cycle, integral and right-hand side runs
on the GPU; the host only keeps the outer loop, the radii and the domain shrink (a new context on the
element / node prefix; the solution's column prefix moves device to device, ``hmg_copy_columns_from``).

Randomness is the caller's: ``sigma_cells`` (per unit cell diagonal conductivities) and ``x0`` (initial
guess before interface sum and constraint) are inputs, because the reference draws them from Julia's
unseeded global RNG (:246, :485-488).
"""
import math

import numpy as np

from . import api, inputs, vtk


def compute_boundary_layer(lam, n):
    """:9"""
    return int(math.floor(4 * (n + 1) * lam ** -0.5))


def compute_box_radius(k, n, eps=0.0):
    """:10"""
    return int(math.floor(2.0 ** (n - k * (0.5 - eps))))


def _infnorm(a):
    return np.max(np.abs(a), axis=-1)


def find_elements_in_radius(mesh, radius):
    """:34-38 -- length of the element prefix within ``radius`` (elements sorted by magnitude)."""
    return int(np.searchsorted(_infnorm(inputs.element_centers(mesh)), radius, side="right"))


def find_nodes_in_radius(mesh, radius):
    """:44-48"""
    return int(np.searchsorted(_infnorm(mesh.nodes), radius + 10 * np.finfo(float).eps, side="right"))


def checkerboard_homogenization(n, dim, refinements=2, smoothing_steps=3, tolerance=1e-4, xi=None, sigma_cells=None,
                                x0=None, max_cycles=1000, log=None, device=0, save=None, save_prefix="", shrink="device"):
    """Returns (sigma, history); history[k] = [(residual norm, sigma + dsigma, |dsigma - dsigma_prev|), ...]
    per V-cycle of outer step k -- the @info line of :287.

    ``save`` (:149-152, :219, :302): a level 1..refinements+1 -- writes ``checkerboard.vtu`` (export_domain: the base
    mesh with the conductivities) and, after every outer step k, ``ahom_<k>.vtu`` (export_unknown: v_k on the nodes of
    that level).  ``shrink``: "device" moves the column prefix of x between the contexts on the GPU, "host" through
    a download / upload (kept for the comparison in the tests).  The reference's shrink_level_state (:54-60) slices all
    five vectors of every level; only the finest x carries information across the shrink (b is rebuilt by next_rhs!,
    r / p / Ap and the coarser levels are overwritten by the next V-cycle), so only that prefix is moved."""
    if shrink not in ("device", "host"):
        raise ValueError("shrink must be 'device' or 'host'")
    xi = np.ones(dim) / math.sqrt(dim) if xi is None else np.asarray(xi, dtype=np.float64)      # :62-65
    lam = 1.0
    sigma = 0.0
    box_radius = compute_box_radius(0, n)
    total_radius = box_radius + compute_boundary_layer(lam, n)
    base = inputs.order_by_magnitude(inputs.hypercube(dim, 2 * total_radius, origin=(-float(total_radius),) * dim))
    cond = inputs.conductivity_per_element(base, sigma_cells, (total_radius + 1.0,) * dim)
    grids = refinements + 1
    if save is not None:
        if not 1 <= save <= grids:
            raise ValueError("save must be a level in 1..refinements+1")
        vtk.export_domain(base, cond, save_prefix + "checkerboard")
    g = api.ImplicitFineGrid(base, grids, cond, lam=lam, device=device)
    old = None
    top = g.state(grids)
    top.x.set(x0)
    api.broadcast_interfaces(top.x, g, grids)
    api.apply_constraint(top.x, grids, g)
    api.rhs_a_xi_grad_v(top.b, g, xi)
    history = []
    try:
        for k in range(n + 1):
            bl = api.BaseLevel(g)                      # assemble_checkerboard + cholesky of :259-261, on GPU 0
            dsigma = dsigma_prev = 0.0
            hist = []
            nsub = find_elements_in_radius(base, box_radius)
            area = api.integrate_area(g, nsub)
            for i in range(max_cycles):
                rn = api.vcycle(g, bl, grids, smoothing_steps, resnorm=True)
                integral = (api.integrate_first_term(top.x, g, nsub, xi) if k == 0
                            else api.integrate_terms(top.x, top.v, g, nsub))
                dsigma = 2.0 ** k * integral / area
                hist.append((rn, sigma + dsigma, abs(dsigma - dsigma_prev)))
                if log:
                    log(k, i + 1, *hist[-1])
                if abs(dsigma - dsigma_prev) < tolerance:
                    break
                dsigma_prev = dsigma
            history.append(hist)
            if save is not None:
                vtk.export_unknown(g, top.x, k, save, save_prefix + f"ahom_{k}")
            sigma += dsigma
            lam /= 2
            box_radius = compute_box_radius(k + 1, n)
            boundary_layer = compute_boundary_layer(lam, n)
            if box_radius + boundary_layer > total_radius:
                break
            # shrink the domain: element / node prefixes of the magnitude-ordered mesh (:297-339)
            total_radius = box_radius + boundary_layer
            nn = find_nodes_in_radius(base, total_radius)
            ne = find_elements_in_radius(base, total_radius)
            base = api.Mesh(base.nodes[:nn], base.elements[:ne])
            cond = np.ascontiguousarray(cond[:ne])
            if shrink == "host":
                x_host = top.x.get()[:, :ne]
                g.close()
                g = api.ImplicitFineGrid(base, grids, cond, lam=lam, device=device)
                top = g.state(grids)
                top.x.set(np.asfortranarray(x_host))
            else:
                old, old_top = g, top
                g = api.ImplicitFineGrid(base, grids, cond, lam=lam, device=device)
                top = g.state(grids)
                top.x.copy_columns_from(old_top.x)         # shrink_level_state (:54-60) on the device
                old.close()
                old = None
            api.apply_constraint(top.x, grids, g)
            top.v.copy_from(top.x)                     # v_prev
            api.next_rhs(top.b, top.x, g)
    finally:
        g.close()
        if old is not None:
            old.close()
    return sigma, history
