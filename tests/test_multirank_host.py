"""The multi-GPU host logic on two CPU processes (gloo): the partition of the coarse elements, the
split of the interface cells into rank-local and cut cells, the packed cut-buffer layout every rank
derives on its own, the owner-count weights of the fused dot products and the one-rank-reports-a-node
rule of the coarse solve.  Each rank emulates with numpy exactly what the device path does around its
NCCL calls (pack partial sums -> all-reduce -> unpack) and compares with the oracle on the whole mesh.
"""
import ctypes as C
import socket

import numpy as np
import pytest

import hmgb200 as hmg
from hmgb200 import _lib as L

CASES = [(3, 2, 3), (2, 4, 4)]      # dim, cells per side, grids


def vp(a):
    return a.ctypes.data_as(C.c_void_p)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def partition_of(lib, mesh, owner, rank, nranks):
    dim, ne, nn = mesh.dim, mesh.nelements, mesh.nnodes
    el1 = np.ascontiguousarray(mesh.elements + 1, dtype=np.int64)
    own = np.ascontiguousarray(owner, dtype=np.int32)
    nel = C.c_int64()
    L.check_host(lib.hmg_host_partition_elements(dim, ne, nn, vp(el1), vp(own), rank, nranks, C.byref(nel), None, None, None,
                                                 None, None))
    nel = nel.value
    out = dict(l2g=np.zeros(nel, np.int64), cmask=np.zeros(nel, np.uint16), mult=np.zeros((nel, 16), np.uint8),
               node_first=np.zeros(nn, np.int32), node_contrib=np.zeros(nn, np.uint8))
    L.check_host(lib.hmg_host_partition_elements(dim, ne, nn, vp(el1), vp(own), rank, nranks, C.byref(C.c_int64()),
                                                 vp(out["l2g"]), vp(out["cmask"]), vp(out["mult"]), vp(out["node_first"]),
                                                 vp(out["node_contrib"])))
    for cut in (0, 1):
        for kind in range(3):
            sizes = np.zeros(3, np.int64)
            L.check_host(lib.hmg_host_partition_cells(dim, ne, nn, vp(el1), vp(own), rank, nranks, kind, cut, vp(sizes), None,
                                                      None, None, None, None))
            nc, nent = int(sizes[0]), int(sizes[1])
            off, el, lid = np.zeros(nc + 1, np.int64), np.zeros(nent, np.int64), np.zeros(nent, np.int64)
            slot, first = np.zeros(nc, np.int64), np.zeros(nc, np.uint8)
            L.check_host(lib.hmg_host_partition_cells(dim, ne, nn, vp(el1), vp(own), rank, nranks, kind, cut, vp(sizes), vp(off),
                                                      vp(el), vp(lid), vp(slot) if cut else None, vp(first) if cut else None))
            out[(cut, kind)] = dict(off=off, el=el, lid=lid, slot=slot, first=first, nglobal=int(sizes[2]))
    out["shared_with"] = np.zeros((nranks, 3), np.int64)
    for kind in range(3):
        sizes = np.zeros(2, np.int64)
        L.check_host(lib.hmg_host_partition_peers(dim, ne, nn, vp(el1), vp(own), rank, nranks, kind, vp(sizes), None, None, None,
                                                  None, None))
        nc, npe = int(sizes[0]), int(sizes[1])
        poff, prank, pidx, mypos = np.zeros(nc + 1, np.int64), np.zeros(npe, np.int32), np.zeros(npe, np.int32), np.zeros(nc, np.int32)
        L.check_host(lib.hmg_host_partition_peers(dim, ne, nn, vp(el1), vp(own), rank, nranks, kind, vp(sizes), vp(poff), vp(prank),
                                                  vp(pidx), vp(mypos), vp(out["shared_with"])))
        out[(1, kind)].update(poff=poff, prank=prank, pidx=pidx, mypos=mypos)
    return out


def paired_rows(lib, dim, levels, level, kind, lid):
    n = C.c_int64()
    L.check_host(lib.hmg_host_interface_rows(dim, levels, level, kind, lid, None, C.byref(n)))
    rows = np.zeros(max(n.value, 1), np.int32)
    L.check_host(lib.hmg_host_interface_rows(dim, levels, level, kind, lid, vp(rows), C.byref(n)))
    return rows[:n.value]


def _worker(rank, world, port, case):
    import torch.distributed as dist
    from oracle.mesh import Mesh as OMesh
    from oracle.implicit import ImplicitFineGrid as OImplicit, broadcast_interfaces, copy_to_base, zero_out_all_but_one

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        import torch
        dim, c, levels = case[:3]
        lib = hmg.load()
        mesh, _ = hmg.inputs.checkerboard_problem(dim, c)
        if len(case) > 3 and case[3] == "random":
            # every element on a random rank: cut cells everywhere, owners of a cell on any subset of the ranks
            owner = np.random.default_rng(17).integers(0, world, mesh.nelements).astype(np.int32)
            owner[:world] = np.arange(world)
        else:
            owner = hmg.inputs.spatial_partition(mesh, world)
        assert set(owner.tolist()) == set(range(world))
        P = partition_of(lib, mesh, owner, rank, world)
        l2g = P["l2g"]
        assert np.array_equal(l2g, np.nonzero(owner == rank)[0])          # global order preserved
        oimp = OImplicit(OMesh(mesh.nodes, mesh.elements), levels)
        rng = np.random.default_rng(5)                                    # the same stream on both ranks
        nkinds = (1, 2) if dim == 2 else (0, 1, 2)
        for level in range(1, levels + 1):
            nf = oimp.nf(level)
            x = np.asfortranarray(rng.random((nf, mesh.nelements)))
            expect = broadcast_interfaces(x.copy(order="F"), oimp, level)
            xl = np.asfortranarray(x[:, l2g])
            rows = {(k, lid): paired_rows(lib, dim, levels, level, k, lid)
                    for k in nkinds for lid in range({0: 4, 1: 6 if dim == 3 else 3, 2: dim + 1}[k])}
            npc = {k: len(rows[(k, 0)]) for k in nkinds}
            # cells whose owners are all local: summed in place, ascending owner order
            for k in nkinds:
                m = P[(0, k)]
                for cell in range(len(m["off"]) - 1):
                    own = range(m["off"][cell], m["off"][cell + 1])
                    s = np.zeros(npc[k])
                    for o in own:
                        s = s + xl[rows[(k, int(m["lid"][o]))], m["el"][o]]
                    for o in own:
                        xl[rows[(k, int(m["lid"][o]))], m["el"][o]] = s
            # cut cells: packed partial sums, one all-reduce, totals written back
            base, total = {}, 0
            for k in (0, 1, 2):
                base[k] = total
                total += P[(1, k)]["nglobal"] * npc.get(k, 0) if k in nkinds else 0
            send = np.zeros(max(total, 1))
            for k in nkinds:
                m = P[(1, k)]
                for cell in range(len(m["off"]) - 1):
                    s = np.zeros(npc[k])
                    for o in range(m["off"][cell], m["off"][cell + 1]):
                        s = s + xl[rows[(k, int(m["lid"][o]))], m["el"][o]]
                    b = base[k] + m["slot"][cell] * npc[k]
                    send[b:b + npc[k]] = s
            recv = torch.from_numpy(send.copy())
            dist.all_reduce(recv)
            recv = recv.numpy()
            for k in nkinds:
                m = P[(1, k)]
                for cell in range(len(m["off"]) - 1):
                    b = base[k] + m["slot"][cell] * npc[k]
                    for o in range(m["off"][cell], m["off"][cell + 1]):
                        xl[rows[(k, int(m["lid"][o]))], m["el"][o]] = recv[b:b + npc[k]]
            assert np.allclose(xl, expect[:, l2g], rtol=1e-14, atol=0), (rank, level)

            # the same through the neighbour exchange: one message per rank that shares a cut cell, laid out as
            # [shared faces x npf][shared edges x npe][shared vertices]; totals added in ascending rank order
            xp = np.asfortranarray(x[:, l2g])
            for k in nkinds:                                  # local cells first, as above
                m = P[(0, k)]
                for cell in range(len(m["off"]) - 1):
                    own = range(m["off"][cell], m["off"][cell + 1])
                    s = np.zeros(npc[k])
                    for o in own:
                        s = s + xp[rows[(k, int(m["lid"][o]))], m["el"][o]]
                    for o in own:
                        xp[rows[(k, int(m["lid"][o]))], m["el"][o]] = s
            shared = [torch.zeros((world, 3), dtype=torch.int64) for _ in range(world)]
            dist.all_gather(shared, torch.from_numpy(P["shared_with"].copy()))
            shared = [t.numpy() for t in shared]
            for q in range(world):                            # both sides count the same shared cells
                assert np.array_equal(shared[q][rank], P["shared_with"][q])

            def layout(table):                                # kbase[(q, kind)] of the rank that owns `table`
                kb, at = {}, 0
                for q in range(world):
                    for k in (0, 1, 2):
                        kb[(q, k)] = at
                        at += int(table[q, k]) * npc.get(k, 0)
                return kb, at
            kb, mine = layout(P["shared_with"])
            longest = max(layout(t)[1] for t in shared)
            msg = np.zeros(max(longest, 1))
            part = {}
            for k in nkinds:
                m = P[(1, k)]
                for cell in range(len(m["off"]) - 1):
                    s = np.zeros(npc[k])
                    for o in range(m["off"][cell], m["off"][cell + 1]):
                        s = s + xp[rows[(k, int(m["lid"][o]))], m["el"][o]]
                    part[(k, cell)] = s
                    for j in range(m["poff"][cell], m["poff"][cell + 1]):
                        b = kb[(int(m["prank"][j]), k)] + int(m["pidx"][j]) * npc[k]
                        msg[b:b + npc[k]] = s
            got = [torch.zeros(len(msg), dtype=torch.float64) for _ in range(world)]
            dist.all_gather(got, torch.from_numpy(msg))
            for k in nkinds:
                m = P[(1, k)]
                for cell in range(len(m["off"]) - 1):
                    tot = np.zeros(npc[k])
                    peers = list(range(m["poff"][cell], m["poff"][cell + 1]))
                    assert list(m["prank"][peers]) == sorted(m["prank"][peers])
                    for n_, j in enumerate(peers):
                        if n_ == m["mypos"][cell]:
                            tot = tot + part[(k, cell)]
                        q = int(m["prank"][j])
                        b = layout(shared[q])[0][(rank, k)] + int(m["pidx"][j]) * npc[k]     # q's message for me
                        tot = tot + got[q].numpy()[b:b + npc[k]]
                    if len(peers) == m["mypos"][cell]:
                        tot = tot + part[(k, cell)]
                    for o in range(m["off"][cell], m["off"][cell + 1]):
                        xp[rows[(k, int(m["lid"][o]))], m["el"][o]] = tot
            assert np.allclose(xp, expect[:, l2g], rtol=1e-14, atol=0), (rank, level)

            # owner-count weights: sum_ranks sum_entries mult * p * Ap_local == dot(p, broadcast(Ap)) on the whole mesh
            p_glob = broadcast_interfaces(np.asfortranarray(rng.random((nf, mesh.nelements))), oimp, level)
            Ap_loc = np.asfortranarray(rng.random((nf, mesh.nelements)))
            w = np.ones((nf, len(l2g)))
            for k in nkinds:
                for lid in range({0: 4, 1: 6 if dim == 3 else 3, 2: dim + 1}[k]):
                    cls = lib.hmg_host_class_of(dim, k, lid)
                    w[rows[(k, lid)], :] = P["mult"][:, cls][None, :]
            part = torch.tensor([float(np.sum(w * p_glob[:, l2g] * Ap_loc[:, l2g]))], dtype=torch.float64)
            dist.all_reduce(part)
            ref = float(np.sum(p_glob * broadcast_interfaces(Ap_loc.copy(order="F"), oimp, level)))
            assert abs(part.item() - ref) <= 1e-12 * abs(ref), (rank, level)

            # zero_out_all_but_one: the globally first owner keeps its value (wherever it lives)
            z = np.asfortranarray(x[:, l2g])
            for cut in (0, 1):
                for k in nkinds:
                    m = P[(cut, k)]
                    for cell in range(len(m["off"]) - 1):
                        for o in range(m["off"][cell], m["off"][cell + 1]):
                            keep = o == m["off"][cell] and (cut == 0 or m["first"][cell])
                            if not keep:
                                z[rows[(k, int(m["lid"][o]))], m["el"][o]] = 0.0
            assert np.array_equal(z, zero_out_all_but_one(x.copy(order="F"), oimp, level)[:, l2g]), (rank, level)

        # coarse solve input: every base node is reported by exactly one rank
        v = broadcast_interfaces(np.asfortranarray(rng.random((dim + 1, mesh.nelements))), oimp, 1)
        u = np.zeros(mesh.nnodes)
        vrow = [int(paired_rows(lib, dim, levels, 1, 2, lid)[0]) for lid in range(dim + 1)]
        for n in range(mesh.nnodes):
            f = int(P["node_first"][n])
            if f >= 0 and P["node_contrib"][n]:
                u[n] = v[vrow[f & 7], l2g[f >> 3]]
        ut = torch.from_numpy(u)
        dist.all_reduce(ut)
        assert np.array_equal(ut.numpy(), copy_to_base(np.zeros(mesh.nnodes), v, oimp))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", CASES, ids=["tet-c2-L3", "tri-c4-L4"])
def test_two_ranks_reproduce_the_global_interface_sums(case):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), case), nprocs=2, join=True)


def test_four_ranks_share_edges_and_vertices():
    """2 x 2 blocks: the cut cells along the middle are shared by up to four ranks (several peers per cell)."""
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(4, _free_port(), (3, 2, 3)), nprocs=4, join=True)


def test_three_ranks_random_ownership():
    """No spatial structure at all: every coarse element on a random one of three ranks (an odd count, so no message
    layout is symmetric by accident) -- the cut-cell enumeration, the pairwise message layouts and the rank-ordered
    sums must still reproduce the global interface sums."""
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(3, _free_port(), (3, 2, 3, "random")), nprocs=3, join=True)


def test_eight_ranks_octants():
    """2 x 2 x 2 blocks (the partition of the 8-GPU bench): the centre vertex is shared by all eight ranks, the axes'
    edges by four, the mid-planes' faces by two; every rank exchanges with its seven neighbours."""
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(8, _free_port(), (3, 2, 2)), nprocs=8, join=True)


def test_single_rank_partition_is_the_whole_mesh():
    lib = hmg.load()
    mesh, _ = hmg.inputs.checkerboard_problem(3, 2)
    P = partition_of(lib, mesh, np.zeros(mesh.nelements, np.int32), 0, 1)
    assert np.array_equal(P["l2g"], np.arange(mesh.nelements))
    assert all(P[(1, k)]["nglobal"] == 0 for k in range(3))
    assert P["node_contrib"].min() == 1


def test_eight_ranks_octants():
    """2 x 2 x 2 octants (the partition of the 8-GPU runs): seven neighbours per rank, the centre vertex shared by all
    eight ranks, the cut edges along the three axes by four."""
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(8, _free_port(), (3, 2, 2)), nprocs=8, join=True)
