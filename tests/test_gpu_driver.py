"""The callers on either side of the V-cycle, on the device: rhs_a_xi_grad_v!,
integrate_first_term, integrate_terms, integrate_area, next_rhs!
(src/examples/homogenized_coefficients.jl:449-474, 592-713) against the oracle, and the homogenized
coefficient of checkerboard_homogenization (:174-343) end to end: sigma within 1e-8 relative, the
residual history within 1e-10 relative per cycle."""
import numpy as np
import pytest

import hmgb200 as hmg
from parity_common import Pair, relerr
from oracle import driver as od
from oracle import implicit as oi
from oracle.fem import partial_derivatives_functionals

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dim,c,levels", [(2, 5, 4), (3, 3, 3), (3, 2, 5)], ids=["tri-c5-L4", "tet-c3-L3", "tet-c2-L5"])
def test_driver_functionals_match_oracle(dim, c, levels):
    pair = Pair(dim, c, levels, lam=0.6, ordered=True)
    try:
        L = levels
        xi = np.array([0.3, -0.8, 0.5][:dim])
        dphi = partial_derivatives_functionals(pair.oimp.refined_mesh(L))
        st = pair.g.state(L)
        ost = pair.ostates[-1]
        # rhs
        od.rhs_a_xi_grad_v(ost.b, dphi, pair.oimp, pair.sigma, xi)
        hmg.rhs_a_xi_grad_v(st.b, pair.g, xi)
        assert relerr(st.b.get(), ost.b) <= 1e-13
        # integrals over an element prefix
        v = pair.rand(L)
        v1 = pair.rand(L)
        st.x.set(v)
        st.v.set(v1)
        ne = pair.mesh.nelements
        for nsub in (ne, ne // 2 + 1, 1):
            a = od.integrate_area(pair.oops[-1], pair.oimp, nsub)
            assert abs(hmg.integrate_area(pair.g, nsub) - a) <= 1e-13 * abs(a)
            f = od.integrate_first_term(v, dphi, pair.oimp, nsub, pair.oops[-1], pair.sigma, xi)
            assert abs(hmg.integrate_first_term(st.x, pair.g, nsub, xi) - f) <= 1e-12 * max(abs(f), 1.0)
            t = od.integrate_terms(v, v1, pair.oimp, nsub, pair.oops[-1])
            assert abs(hmg.integrate_terms(st.x, st.v, pair.g, nsub) - t) <= 1e-12 * max(abs(t), 1.0)
        # next right-hand side
        od.next_rhs(ost.b, v, pair.oimp, pair.oops[-1])
        hmg.next_rhs(st.b, st.x, pair.g)
        assert relerr(st.b.get(), ost.b) <= 1e-13
    finally:
        pair.close()


@pytest.mark.parametrize("n,dim,refinements,tol", [(1, 2, 3, 1e-5), (0, 3, 2, 1e-4)], ids=["tri-n1-r3", "tet-n0-r2"])
def test_homogenized_coefficient_matches_oracle(n, dim, refinements, tol):
    """checkerboard_homogenization(n, Tri64|Tet64, refinements=..., tolerance=...) on identical seeded inputs."""
    rng = np.random.default_rng(42)
    radius = od.compute_box_radius(0, n) + od.compute_boundary_layer(1.0, n)
    cells = np.where(rng.random((2 * radius,) * dim + (dim,)) < 0.5, 1.0, 9.0)
    base, _ = od.make_base(dim, n)
    nf = hmg.inputs.nf_of_level(dim, refinements + 1)
    x0 = np.asfortranarray(rng.random((nf, base.nelements)))
    so, ho = od.checkerboard_homogenization(n, dim, refinements=refinements, tolerance=tol, sigma_cells=cells, x0=x0.copy(order="F"))
    sg, hg = hmg.driver.checkerboard_homogenization(n, dim, refinements=refinements, tolerance=tol, sigma_cells=cells, x0=x0)
    assert len(ho) == len(hg)
    for a, b in zip(ho, hg):
        assert len(a) == len(b)                       # the same number of V-cycles
        for (ro, s_o, _), (rg, s_g, _) in zip(a, b):
            assert abs(rg - ro) <= 1e-10 * ro
            assert abs(s_g - s_o) <= 1e-8 * abs(s_o)
    assert abs(sg - so) <= 1e-8 * abs(so)


@pytest.mark.parametrize("name", ["homogenization_tri_n1_r3", "homogenization_tet_n0_r2", "homogenization_C1_tri_n3_r4"])
def test_homogenized_coefficient_matches_golden(name):
    """The device driver against the oracle's committed golden histories (tests/golden/, no oracle run here);
    `homogenization_C1_tri_n3_r4` is BASELINE.json configs[0], the README example
    checkerboard_homogenization(3, Tri64, refinements=4, tolerance=1e-3)."""
    import json
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    import make_golden as mg
    gold = json.load(open(os.path.join(here, "golden", name + ".json")))
    n, dim, refinements, tol, seed = mg.CASES[name]
    cells, x0, _ = mg.inputs(n, dim, refinements, seed)
    sg, hg = hmg.driver.checkerboard_homogenization(n, dim, refinements=refinements, tolerance=tol, sigma_cells=cells, x0=x0)
    assert [len(s) for s in hg] == [len(s) for s in gold["history"]]
    for a, b in zip(hg, gold["history"]):
        for (rg, s_g, _), (ro, s_o, _) in zip(a, b):
            assert abs(rg - ro) <= 1e-10 * ro
            assert abs(s_g - s_o) <= 1e-8 * abs(s_o)
    assert abs(sg - gold["sigma"]) <= 1e-8 * abs(gold["sigma"])
