"""Parity of the CUDA path (through the C ABI) with the CPU oracle on identical seeded inputs.

Tolerances: A*x 1e-12 relative, residual history 1e-10 relative per cycle.
Pure data movement (interface sums in owner order, constraint, gather/scatter) is bit-exact.
"""
import numpy as np
import pytest

import hmgb200 as hmg
from parity_common import Pair, relerr
from oracle import implicit as oi
from oracle import operators as oo
from oracle import multigrid as om

pytestmark = pytest.mark.gpu

CASES = [(2, 5, 5), (2, 3, 7), (3, 3, 4), (3, 2, 5), (2, 4, 1), (3, 2, 2)]
IDS = ["tri-c5-L5", "tri-c3-L7", "tet-c3-L4", "tet-c2-L5", "tri-c4-L1", "tet-c2-L2"]


# BASELINE.json configs[4]: random piecewise coefficient field; and the magnitude-ordered mesh of the driver
# (elements of a cell are no longer consecutive)
CASES += [(2, 4, 5, "random"), (3, 2, 4, "random"), (3, 3, 3, "ordered")]
IDS += ["tri-c4-L5-random", "tet-c2-L4-random", "tet-c3-L3-ordered"]
# the depths the bench quotes (C4: 3D with 6 grids, C2: 2D with 8 grids): other launch shapes of the apply kernel
# (lines per task, chunk size, two staging slots per converter warp in the fused direction update)
CASES += [(3, 2, 6), (2, 2, 8)]
IDS += ["tet-c2-L6", "tri-c2-L8"]


@pytest.fixture(params=CASES, ids=IDS)
def pair(request):
    dim, c, levels = request.param[:3]
    kind = request.param[3] if len(request.param) > 3 else "checkerboard"
    p = Pair(dim, c, levels, lam=0.7, field="random" if kind == "random" else "checkerboard", seed=2 if kind == "random" else 1,
             ordered=kind == "ordered")
    yield p
    p.close()


def test_upload_download_roundtrip(pair):
    for level in range(1, pair.levels + 1):
        x = pair.rand(level)
        assert np.array_equal(pair.g.state(level).x.set(x).get(), x)


def test_mul_matches_oracle(pair):
    """mul!(alpha, base, A, x, y) on every level (src/apply_local_operators.jl:85-120)."""
    for level in range(1, pair.levels + 1):
        x, y = pair.rand(level), pair.rand(level)
        st = pair.g.state(level)
        st.x.set(x)
        st.r.set(y)
        hmg.mul(-1.3, pair.g, st.x, st.r)
        expect = oo.mul(-1.3, pair.obase, pair.oops[level - 1], x, y.copy(order="F"))
        assert relerr(st.r.get(), expect) <= 1e-12, level


def test_mass_term_is_skipped_for_zero_lambda(pair):
    level = pair.levels
    x = pair.rand(level)
    st = pair.g.state(level)
    pair.g.set_lambda(0.0)
    pair.oops[level - 1].lam = 0.0
    st.x.set(x)
    st.r.fill(0.0)
    hmg.mul(1.0, pair.g, st.x, st.r)
    expect = oo.mul(1.0, pair.obase, pair.oops[level - 1], x, np.zeros_like(x))
    assert relerr(st.r.get(), expect) <= 1e-12


def test_interface_primitives_are_bit_exact(pair):
    for level in range(1, pair.levels + 1):
        x = pair.rand(level)
        st = pair.g.state(level)
        st.x.set(x)
        hmg.broadcast_interfaces(st.x, pair.g, level)
        expect = oi.broadcast_interfaces(x.copy(order="F"), pair.oimp, level)
        assert np.array_equal(st.x.get(), expect), level
        hmg.apply_constraint(st.x, level, pair.g)
        oi.apply_constraint(expect, level, pair.constraint, pair.oimp)
        assert np.array_equal(st.x.get(), expect), level
        y = pair.rand(level)
        st.r.set(y)
        hmg.zero_out_all_but_one(st.r, pair.g, level)
        assert np.array_equal(st.r.get(), oi.zero_out_all_but_one(y.copy(order="F"), pair.oimp, level)), level


def test_global_product_matches_oracle(pair):
    """Ap = broadcast(constraint(A p)) (src/multigrid.jl:58-61), the benchmarked A*x."""
    level = pair.levels
    p = pair.rand(level)
    oi.broadcast_interfaces(p, pair.oimp, level)
    oi.apply_constraint(p, level, pair.constraint, pair.oimp)
    st = pair.g.state(level)
    st.p.set(p)
    hmg.apply_global(pair.g, st.p, st.Ap)
    Ap = oo.mul(1.0, pair.obase, pair.oops[level - 1], p, np.zeros_like(p))
    oi.apply_constraint(Ap, level, pair.constraint, pair.oimp)
    oi.broadcast_interfaces(Ap, pair.oimp, level)
    assert relerr(st.Ap.get(), Ap) <= 1e-12


def test_local_residual_matches_oracle(pair):
    level = pair.levels
    os_ = pair.ostates[level - 1]
    os_.x[:, :] = pair.rand(level)
    os_.b[:, :] = pair.rand(level)
    st = pair.g.state(level)
    st.x.set(os_.x)
    st.b.set(os_.b)
    hmg.local_residual(pair.g, level)
    oo.local_residual(pair.oimp, pair.oops[level - 1], os_, level)
    assert relerr(st.r.get(), os_.r) <= 1e-12


def test_transfer_operators_match_oracle(pair):
    for k in range(2, pair.levels + 1):
        P = pair.oimp.reference.interops[k - 2]
        r = pair.rand(k)
        pair.g.state(k).r.set(r)
        hmg.restrict_to(pair.g, k)
        assert relerr(pair.g.state(k - 1).b.get(), P.T @ r) <= 1e-14, k
        xc, xf = pair.rand(k - 1), pair.rand(k)
        pair.g.state(k - 1).x.set(xc)
        pair.g.state(k).x.set(xf)
        hmg.interpolate_and_sum_to(pair.g, k)
        assert relerr(pair.g.state(k).x.get(), xf + P @ xc) <= 1e-14, k


def test_level1_gather_scatter_bit_exact(pair):
    v = pair.rand(1)
    st = pair.g.state(1)
    st.b.set(v)
    u = hmg.copy_to_base(pair.g, st.b)
    expect = oi.copy_to_base(np.zeros(pair.mesh.nnodes), v, pair.oimp)
    assert np.array_equal(u, expect)
    ub = pair.rng.random(pair.mesh.nnodes)
    hmg.distribute(pair.g, st.x, ub)
    assert np.array_equal(st.x.get(), oi.distribute(np.zeros_like(v, order="F"), ub, pair.oimp))


def test_dot_counts_all_stored_entries(pair):
    level = pair.levels
    a, b = pair.rand(level), pair.rand(level)
    st = pair.g.state(level)
    st.p.set(a)
    st.Ap.set(b)
    got = hmg.dot(pair.g, st.p, st.Ap)
    assert abs(got - om.dot(a, b)) <= 1e-12 * abs(om.dot(a, b))


def test_smoothing_steps_match_oracle(pair):
    level = pair.levels
    os_ = pair.ostates[level - 1]
    x = pair.rand(level)
    oi.broadcast_interfaces(x, pair.oimp, level)
    oi.apply_constraint(x, level, pair.constraint, pair.oimp)
    os_.x[:, :] = x
    os_.b[:, :] = pair.rand(level)
    st = pair.g.state(level)
    st.x.set(os_.x)
    st.b.set(os_.b)
    hmg.smoothing_steps(3, pair.g, level)
    om.smoothing_steps(3, pair.oimp, pair.oops[level - 1], os_, level)
    assert relerr(st.x.get(), os_.x) <= 1e-11
    assert relerr(st.r.get(), os_.r) <= 1e-10


def _vcycle_history(pair, cycles, internal_coarse):
    L = pair.levels
    top_o = pair.ostates[-1]
    x = pair.rand(L)
    oi.broadcast_interfaces(x, pair.oimp, L)
    oi.apply_constraint(x, L, pair.constraint, pair.oimp)
    top_o.x[:, :] = x
    oi.local_rhs(top_o.b, pair.oimp)
    st = pair.g.state(L)
    st.x.set(top_o.x)
    st.b.set(top_o.b)
    obl, A, interior = pair.obase_level()
    bl = hmg.BaseLevel(pair.g) if internal_coarse else hmg.BaseLevel(pair.g, A, interior)
    hist_o, hist_g = [], []
    for _ in range(cycles):
        om.vcycle(pair.oimp, obl, pair.oops, pair.ostates, L, 3)
        oi.zero_out_all_but_one(top_o.r, pair.oimp, L)
        hist_o.append(float(np.linalg.norm(top_o.r.ravel(order="K"))))
        hist_g.append(hmg.vcycle(pair.g, bl, L, 3, resnorm=True))
    return np.array(hist_o), np.array(hist_g), top_o.x, st.x.get()


@pytest.mark.parametrize("internal_coarse", [False, True], ids=["given-coarse-matrix", "assembled-coarse-matrix"])
def test_vcycle_residual_history_matches_oracle(pair, internal_coarse):
    """vcycle! (src/multigrid.jl:73-119) + the logged residual
    (src/examples/homogenized_coefficients.jl:286-287): <= 1e-10 relative per cycle."""
    if pair.levels < 2 or len(pair.obase_level()[2]) == 0:
        pytest.skip("needs at least two grids and an interior base node")
    ho, hg, xo, xg = _vcycle_history(pair, 5, internal_coarse)
    assert np.all(np.abs(hg - ho) <= 1e-10 * ho), (ho, hg)
    assert ho[-1] < 0.2 * ho[0]           # it does converge
    assert relerr(xg, xo) <= 1e-10


def test_batched_vcycles_equal_single_calls(pair):
    if pair.levels < 2 or len(pair.obase_level()[2]) == 0:
        pytest.skip("needs at least two grids and an interior base node")
    L = pair.levels
    st = pair.g.state(L)
    x0, b0 = pair.rand(L), pair.rand(L)
    bl = hmg.BaseLevel(pair.g)
    st.x.set(x0); st.b.set(b0)
    single = [hmg.vcycle(pair.g, bl, L, 3, resnorm=True) for _ in range(3)]
    st.x.set(x0); st.b.set(b0)
    batched = hmg.vcycles(pair.g, bl, L, 3, 3)
    assert np.array_equal(np.array(single), batched)      # deterministic reductions


@pytest.mark.parametrize("half_min", ["1", "1000000"], ids=["half-traffic-tiles", "full-matvec"])
def test_coarse_solve_variants_match_oracle(monkeypatch, half_min):
    """The coarsest-grid solve x = A^-1 b on GPU 0 (src/multigrid.jl:75-93): both mat-vec kernels (the tiled
    half-traffic one is the default from 2048 interior nodes on) against the oracle's sparse direct solve, on a base
    mesh with 19^2 = 361 interior nodes (3 x 3 tiles)."""
    monkeypatch.setenv("HMG_SYMV_HALF_MIN", half_min)
    p = Pair(2, 20, 3, lam=0.7)
    try:
        ho, hg, xo, xg = _vcycle_history(p, 3, True)
        assert np.all(np.abs(hg - ho) <= 1e-10 * ho), (ho, hg)
        assert relerr(xg, xo) <= 1e-10
    finally:
        p.close()


def _irregular_mesh(dim):
    """A base mesh that is not a lattice of unit cells: vertices jittered (every element has its own Jacobian),
    elements in a shuffled order (units of 32 mix orientations, neighbours are far apart in the column order)."""
    from oracle.mesh import hypercube, sort_element_nodes
    rng = np.random.default_rng(11)
    m = hypercube(dim, 3 if dim == 3 else 5)
    nodes = m.nodes.copy()
    lo, hi = nodes.min(axis=0), nodes.max(axis=0)
    inner = np.all((nodes > lo) & (nodes < hi), axis=1)
    nodes[inner] += 0.15 * (rng.random((int(inner.sum()), dim)) - 0.5)
    elements = sort_element_nodes(m.elements)[rng.permutation(m.nelements)]
    return hmg.Mesh(nodes, elements), rng.uniform(1.0, 9.0, size=(m.nelements, dim))


@pytest.mark.parametrize("dim,levels", [(2, 5), (3, 4)], ids=["tri-jittered-L5", "tet-jittered-L4"])
def test_irregular_base_mesh_matches_oracle(dim, levels):
    """Nothing in the device path assumes the checkerboard lattice: A*x and the V-cycle on a jittered, shuffled
    base mesh with a random anisotropic coefficient per element."""
    mesh, sigma = _irregular_mesh(dim)
    p = Pair(dim, 0, levels, lam=0.9, mesh=mesh, sigma=sigma)
    try:
        L = levels
        v = p.rand(L)
        oi.broadcast_interfaces(v, p.oimp, L)
        oi.apply_constraint(v, L, p.constraint, p.oimp)
        st = p.g.state(L)
        st.p.set(v)
        hmg.apply_global(p.g, st.p, st.Ap)
        Ap = oo.mul(1.0, p.obase, p.oops[L - 1], v, np.zeros_like(v))
        oi.apply_constraint(Ap, L, p.constraint, p.oimp)
        oi.broadcast_interfaces(Ap, p.oimp, L)
        assert relerr(st.Ap.get(), Ap) <= 1e-12
        ho, hg, xo, xg = _vcycle_history(p, 4, True)
        assert np.all(np.abs(hg - ho) <= 1e-10 * ho), (ho, hg)
        assert relerr(xg, xo) <= 1e-10
    finally:
        p.close()
