"""The launch-shape knobs that stay in the library are read when a context is created; each one
gets the same parity check as the default: A*x <= 1e-12, smoothing and V-cycle history <= 1e-10 against the oracle.
Also: the coarse inverse follows lambda."""
import numpy as np
import pytest

import hmgb200 as hmg
from parity_common import Pair, relerr
from oracle import implicit as oi, operators as oo, multigrid as om

pytestmark = pytest.mark.gpu

KNOBS = [
    {"HMG_CG_PAIRS": "0"},
    {"HMG_GRAPH": "0"},                                    # every V-cycle launched eagerly (default: CUDA graph from the 2nd on)                                 # interface sums of Ap as their own pass before the CG update
    {"HMG_FUSE_P": "0"},                                   # direction update as its own kernel (p_update + product)
    {"HMG_APPLY_WARPS": "8"},
    {"HMG_APPLY_RUN": "1", "HMG_APPLY_CHUNK_SHIFT": "5"},
    {"HMG_APPLY_RUN": "2", "HMG_APPLY_CHUNK_SHIFT": "6"},
    {"HMG_APPLY_CONVERTERS": "1"},
    {"HMG_APPLY_SLOT_SHIFT": "1"},
    {"HMG_APPLY_SEG_SHIFT": "4", "HMG_APPLY_OVERSUB": "3"},
    {"HMG_APPLY_SEG3_SHIFT": "3"},                         # 3D lines cut into segments of <= 8 nodes (default: whole lines below 6 grids)
    {"HMG_APPLY_SEG3_SHIFT": "30"},
]


@pytest.mark.parametrize("knobs", KNOBS, ids=["+".join(f"{k[4:]}={v}" for k, v in kn.items()) for kn in KNOBS])
@pytest.mark.parametrize("shape", [(3, 2, 5), (2, 3, 7)], ids=["tet-c2-L5", "tri-c3-L7"])
def test_knob_keeps_parity(monkeypatch, knobs, shape):
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    pair = Pair(*shape, lam=0.7)
    try:
        L = pair.levels
        p = pair.rand(L)
        oi.broadcast_interfaces(p, pair.oimp, L)
        oi.apply_constraint(p, L, pair.constraint, pair.oimp)
        st = pair.g.state(L)
        st.p.set(p)
        hmg.apply_global(pair.g, st.p, st.Ap)
        Ap = oo.mul(1.0, pair.obase, pair.oops[L - 1], p, np.zeros_like(p))
        oi.apply_constraint(Ap, L, pair.constraint, pair.oimp)
        oi.broadcast_interfaces(Ap, pair.oimp, L)
        assert relerr(st.Ap.get(), Ap) <= 1e-12
        top = pair.ostates[-1]
        top.x[:, :] = p
        oi.local_rhs(top.b, pair.oimp)
        st.x.set(top.x)
        st.b.set(top.b)
        obl, _, _ = pair.obase_level()
        bl = hmg.BaseLevel(pair.g)
        for _ in range(3):
            om.vcycle(pair.oimp, obl, pair.oops, pair.ostates, L, 3)
            oi.zero_out_all_but_one(top.r, pair.oimp, L)
            ro = float(np.linalg.norm(top.r.ravel(order="K")))
            rg = hmg.vcycle(pair.g, bl, L, 3, resnorm=True)
            assert abs(rg - ro) <= 1e-10 * ro
        assert relerr(st.x.get(), top.x) <= 1e-10
    finally:
        pair.close()


def _history(pair, bl, cycles=3):
    L = pair.levels
    top = pair.ostates[-1]
    obl, A, interior = pair.obase_level()
    out = []
    for _ in range(cycles):
        om.vcycle(pair.oimp, obl, pair.oops, pair.ostates, L, 3)
        oi.zero_out_all_but_one(top.r, pair.oimp, L)
        ro = float(np.linalg.norm(top.r.ravel(order="K")))
        rg = hmg.vcycle(pair.g, bl, L, 3, resnorm=True)
        out.append((ro, rg))
    return out


def test_coarse_inverse_follows_lambda():
    """hmg_set_lambda after the coarse factorisation: a matrix the library assembled is re-assembled on the next V-cycle
    (the reference's driver re-factorises at every outer step, src/examples/homogenized_coefficients.jl:259-261); a matrix
    handed over by the caller makes the V-cycle fail instead of silently solving with the old one."""
    pair = Pair(2, 6, 4, lam=1.0)
    try:
        L = pair.levels
        top = pair.ostates[-1]
        x = pair.rand(L)
        oi.broadcast_interfaces(x, pair.oimp, L)
        oi.apply_constraint(x, L, pair.constraint, pair.oimp)
        top.x[:, :] = x
        oi.local_rhs(top.b, pair.oimp)
        st = pair.g.state(L)
        st.x.set(top.x)
        st.b.set(top.b)
        bl = hmg.BaseLevel(pair.g)                       # assembled inside the library for lambda = 1
        for ro, rg in _history(pair, bl, 2):
            assert abs(rg - ro) <= 1e-10 * ro
        pair.g.set_lambda(0.25)                          # next outer step of the driver: lambda /= 2, twice
        pair.lam = 0.25
        for op in pair.oops:
            op.lam = 0.25
        for ro, rg in _history(pair, bl, 3):
            assert abs(rg - ro) <= 1e-10 * ro, (ro, rg)
        # a caller-owned matrix cannot follow: the call must fail
        _, A, interior = pair.obase_level()
        bl2 = hmg.BaseLevel(pair.g, A, interior)
        hmg.vcycle(pair.g, bl2, L, 3)
        pair.g.set_lambda(0.5)
        with pytest.raises(hmg.HmgError, match="lambda or sigma changed"):
            hmg.vcycle(pair.g, bl2, L, 3)
    finally:
        pair.close()


def test_graph_replay_is_bit_identical_to_eager_launches(monkeypatch):
    """From the second V-cycle of a kind on the launches are replayed from a CUDA graph (HMG_GRAPH, default on): the
    residual history and the solution must equal the eager run bit for bit (deterministic reductions, the same kernels
    with the same arguments), also across a change of lambda, which drops the captured graphs."""
    def run(graph):
        monkeypatch.setenv("HMG_GRAPH", graph)
        pair = Pair(3, 2, 4, lam=0.7)
        try:
            assert pair.g.comm_mode() == "single"
            L = pair.levels
            x = pair.rand(L)
            oi.broadcast_interfaces(x, pair.oimp, L)
            oi.apply_constraint(x, L, pair.constraint, pair.oimp)
            b = pair.rand(L)
            st = pair.g.state(L)
            st.x.set(x)
            st.b.set(b)
            bl = hmg.BaseLevel(pair.g)
            hist = [hmg.vcycle(pair.g, bl, L, 3, resnorm=True) for _ in range(4)]
            pair.g.set_lambda(0.35)
            hist += [hmg.vcycle(pair.g, bl, L, 3, resnorm=True) for _ in range(3)]
            hist += list(hmg.vcycles(pair.g, bl, L, 3, 3))
            return np.array(hist), st.x.get()
        finally:
            pair.close()
    h1, x1 = run("1")
    h0, x0 = run("0")
    assert np.array_equal(h1, h0)
    assert np.array_equal(x1, x0)


@pytest.mark.parametrize("steps", [0, 1, 2, 4])
def test_vcycle_with_other_step_counts(steps):
    """The last CG step of a smoothing call is cut down to alpha and x += alpha p where nobody reads its residual
    (DESIGN.md section 4): with 1 step that is also the FIRST step (direction = r), with 0 steps there is none.  The
    top level's step count varies, the levels below always take 2 (src/multigrid.jl:109)."""
    pair = Pair(3, 2, 4, lam=0.7)
    try:
        L = pair.levels
        top = pair.ostates[-1]
        x = pair.rand(L)
        oi.broadcast_interfaces(x, pair.oimp, L)
        oi.apply_constraint(x, L, pair.constraint, pair.oimp)
        top.x[:, :] = x
        oi.local_rhs(top.b, pair.oimp)
        st = pair.g.state(L)
        st.x.set(top.x)
        st.b.set(top.b)
        obl, _, _ = pair.obase_level()
        bl = hmg.BaseLevel(pair.g)
        for _ in range(3):
            om.vcycle(pair.oimp, obl, pair.oops, pair.ostates, L, steps)
            oi.zero_out_all_but_one(top.r, pair.oimp, L)
            ro = float(np.linalg.norm(top.r.ravel(order="K")))
            rg = hmg.vcycle(pair.g, bl, L, steps, resnorm=True)
            assert abs(rg - ro) <= 1e-10 * ro, (steps, rg, ro)
        assert relerr(st.x.get(), top.x) <= 1e-10
    finally:
        pair.close()
