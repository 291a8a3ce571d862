"""Partitioned contexts (one process per GPU, NCCL): the coarse elements are split over two ranks, only
interface partial sums and scalars move, the coarsest-grid solve stays on rank 0.  Results are compared
with the CPU oracle on the WHOLE mesh: A*x <= 1e-12, residual history <= 1e-10 per cycle."""
import ctypes as C

import numpy as np
import pytest

import hmgb200 as hmg

pytestmark = pytest.mark.gpu

CASES = [(3, 3, 4), (2, 5, 5), (3, 2, 5)]


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _worker(rank, world, nccl_id, case, results, peer="1"):
    import os
    os.environ["HMG_PEER"] = peer          # "1": peer memory over NVLink (default), "0": the NCCL path
    from parity_common import relerr
    from oracle.mesh import Mesh as OMesh
    from oracle.fem import build_local_diffusion_operators, build_local_mass_matrices, assemble_matrix
    from oracle.interfaces import list_boundary_nodes_edges_faces, list_interior_nodes
    from oracle.implicit import ImplicitFineGrid as OImplicit, ZeroDirichletConstraint
    from oracle.operators import L2PlusDivAGrad
    from oracle.multigrid import LevelState as OLevelState, BaseLevel as OBaseLevel
    from oracle import implicit as oi, operators as oo, multigrid as om

    dim, c, levels = case[:3]
    lam = 0.7
    mesh, sigma = hmg.inputs.checkerboard_problem(dim, c)
    if len(case) > 3 and case[3] == "random":
        # every element on a random rank: cut cells everywhere, most shared cells have owners on both sides
        owner = np.random.default_rng(23).integers(0, world, mesh.nelements).astype(np.int32)
        owner[:world] = np.arange(world)
    else:
        owner = hmg.inputs.spatial_partition(mesh, world)
    g = hmg.ImplicitFineGrid(mesh, levels, sigma, lam=lam, device=rank, owner_rank=owner, rank=rank, nranks=world,
                             nccl_id=nccl_id)
    l2g = g.local_elements()
    assert np.array_equal(l2g, np.nonzero(owner == rank)[0])
    assert g.comm_mode() == ("peer" if peer == "1" else "nccl")      # no silent fall-back to the other path
    obase = OMesh(mesh.nodes, mesh.elements)
    oimp = OImplicit(obase, levels)
    z = ZeroDirichletConstraint(*list_boundary_nodes_edges_faces(obase))
    ops = [L2PlusDivAGrad(d, m, z, lam, sigma) for d, m in
           zip(build_local_diffusion_operators(oimp.reference), build_local_mass_matrices(oimp.reference))]
    rng = np.random.default_rng(3)
    L = levels
    st = g.state(L)

    # interface sums across the cut
    x = np.asfortranarray(rng.random((oimp.nf(L), mesh.nelements)))
    st.x.set(np.asfortranarray(x[:, l2g]))
    hmg.broadcast_interfaces(st.x, g, L)
    assert np.allclose(st.x.get(), oi.broadcast_interfaces(x.copy(order="F"), oimp, L)[:, l2g], rtol=1e-14, atol=0)
    st.r.set(np.asfortranarray(x[:, l2g]))
    hmg.zero_out_all_but_one(st.r, g, L)
    assert np.array_equal(st.r.get(), oi.zero_out_all_but_one(x.copy(order="F"), oimp, L)[:, l2g])

    # global product and dot over all stored entries
    p = oi.broadcast_interfaces(np.asfortranarray(rng.random((oimp.nf(L), mesh.nelements))), oimp, L)
    oi.apply_constraint(p, L, z, oimp)
    st.p.set(np.asfortranarray(p[:, l2g]))
    hmg.apply_global(g, st.p, st.Ap)
    Ap = oo.mul(1.0, obase, ops[L - 1], p, np.zeros_like(p))
    oi.apply_constraint(Ap, L, z, oimp)
    oi.broadcast_interfaces(Ap, oimp, L)
    scale = np.abs(Ap).max()
    assert np.abs(st.Ap.get() - Ap[:, l2g]).max() <= 1e-12 * scale
    d = hmg.dot(g, st.p, st.Ap)
    assert abs(d - om.dot(p, Ap)) <= 1e-12 * abs(om.dot(p, Ap))

    # V-cycles: residual history of the whole problem
    ostates = [OLevelState(oimp, l) for l in range(1, levels + 1)]
    top = ostates[-1]
    top.x[:, :] = p
    oi.local_rhs(top.b, oimp)
    st.x.set(np.asfortranarray(top.x[:, l2g]))
    st.b.set(np.asfortranarray(top.b[:, l2g]))
    interior = list_interior_nodes(obase)
    A = assemble_matrix(obase, sigma=sigma, lam=lam)[interior][:, interior]
    obl = OBaseLevel(A, obase.nnodes, interior)
    bl = hmg.BaseLevel(g)
    hist = []
    for _ in range(4):
        om.vcycle(oimp, obl, ops, ostates, L, 3)
        oi.zero_out_all_but_one(top.r, oimp, L)
        ro = float(np.linalg.norm(top.r.ravel(order="K")))
        rg = hmg.vcycle(g, bl, L, 3, resnorm=True)
        assert abs(rg - ro) <= 1e-10 * ro, (rank, rg, ro)
        hist.append(rg)
    assert relerr(st.x.get(), top.x[:, l2g]) <= 1e-10
    assert hist[-1] < 0.5 * hist[0]
    g.close()
    results.put((rank, hist))


@pytest.mark.parametrize("case", CASES, ids=["tet-c3-L4", "tri-c5-L5", "tet-c2-L5"])
def test_two_gpus_match_the_oracle_on_the_whole_mesh(case):
    if _ngpus() < 2:
        pytest.skip("needs two GPUs")
    import torch                      # PyTorch's NCCL is the one every process binds
    import torch.multiprocessing as mp
    assert torch.cuda.is_available()
    raw = (C.c_ubyte * 128)()
    hmg._lib.check(hmg.load().hmg_nccl_unique_id(raw))
    ctx = mp.get_context("spawn")
    results = ctx.SimpleQueue()
    mp.spawn(_worker, args=(2, bytes(raw), case, results), nprocs=2, join=True)
    got = dict(results.get() for _ in range(2))
    assert got[0] == got[1]          # every rank sees the same residual history


def test_two_gpus_random_ownership_matches_the_oracle():
    """No spatial structure: every coarse element on a random one of the two ranks, so nearly every shared face, edge
    and vertex is a cut cell with several local owners on both sides -- the peer-memory exchange carries almost the
    whole interface sum."""
    if _ngpus() < 2:
        pytest.skip("needs two GPUs")
    import torch
    import torch.multiprocessing as mp
    raw = (C.c_ubyte * 128)()
    hmg._lib.check(hmg.load().hmg_nccl_unique_id(raw))
    ctx = mp.get_context("spawn")
    results = ctx.SimpleQueue()
    mp.spawn(_worker, args=(2, bytes(raw), (3, 3, 4, "random"), results), nprocs=2, join=True)
    got = dict(results.get() for _ in range(2))
    assert got[0] == got[1]


def test_two_gpus_nccl_path_matches_the_oracle():
    """HMG_PEER=0: scalar sums by ncclAllReduce, cut cells by grouped ncclSend / ncclRecv (what a box without peer
    mapping falls back to)."""
    if _ngpus() < 2:
        pytest.skip("needs two GPUs")
    import torch
    import torch.multiprocessing as mp
    raw = (C.c_ubyte * 128)()
    hmg._lib.check(hmg.load().hmg_nccl_unique_id(raw))
    ctx = mp.get_context("spawn")
    results = ctx.SimpleQueue()
    mp.spawn(_worker, args=(2, bytes(raw), (3, 3, 4), results, "0"), nprocs=2, join=True)
    got = dict(results.get() for _ in range(2))
    assert got[0] == got[1]


@pytest.mark.parametrize("case", [(3, 4, 4), (2, 6, 5)], ids=["tet-c4-L4", "tri-c6-L5"])
def test_four_gpus_match_the_oracle_on_the_whole_mesh(case):
    """2 x 2 blocks: cut edges and vertices are shared by up to four ranks (several peers per cut cell)."""
    if _ngpus() < 4:
        pytest.skip("needs four GPUs")
    import torch
    import torch.multiprocessing as mp
    raw = (C.c_ubyte * 128)()
    hmg._lib.check(hmg.load().hmg_nccl_unique_id(raw))
    ctx = mp.get_context("spawn")
    results = ctx.SimpleQueue()
    mp.spawn(_worker, args=(4, bytes(raw), case, results), nprocs=4, join=True)
    got = dict(results.get() for _ in range(4))
    assert got[0] == got[1] == got[2] == got[3]


def test_eight_gpus_match_the_oracle_on_the_whole_mesh():
    """2 x 2 x 2 octants of a 4^3-cell mesh, 4 grids: every rank has seven neighbours, the centre vertex is shared by all
    eight ranks, cut edges by four.  The same comparison with the oracle on the whole mesh as on two ranks."""
    if _ngpus() < 8:
        pytest.skip("needs eight GPUs")
    import torch
    import torch.multiprocessing as mp
    raw = (C.c_ubyte * 128)()
    hmg._lib.check(hmg.load().hmg_nccl_unique_id(raw))
    ctx = mp.get_context("spawn")
    results = ctx.SimpleQueue()
    mp.spawn(_worker, args=(8, bytes(raw), (3, 4, 4), results), nprocs=8, join=True)
    got = dict(results.get() for _ in range(8))
    assert all(got[r] == got[0] for r in range(1, 8))
