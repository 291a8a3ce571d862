"""Runs the host-side setup code of the library (reference element, topology, Dirichlet classes, partition over ranks,
neighbour-exchange layout, the host emulation of the apply sweeps) under AddressSanitizer + UBSan, on regular meshes and
on meshes with random holes / shuffled elements / random element ownership.  Driven by tools/asan_host_check.sh, which
builds csrc/{reference,topology,introspect}.cpp with g++ -fsanitize=address,undefined into /tmp."""
import ctypes as C, sys, numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
lib = C.CDLL(sys.argv[1] if len(sys.argv) > 1 else '/tmp/hmg_asan/libhmg_host_asan.so')
lib.hmg_host_last_error.restype = C.c_char_p
def chk(rc):
    if rc != 0: raise RuntimeError(lib.hmg_host_last_error().decode())
vp = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
i64 = C.c_int64
# 1. reference element, every level, both dims (+ refined mesh, interface rows, transfer, local matrix, apply sweep)
for dim, nl in ((2, 8), (3, 6)):
    for level in range(1, nl + 1):
        sizes = (i64 * 8)()
        chk(lib.hmg_host_reference(dim, nl, level, sizes, None, None, None))
        nf, ndir, nc = int(sizes[1]), int(sizes[6]), int(sizes[7])
        h2l = np.zeros(nf, np.int32); G = np.zeros((16 if dim == 3 else 8) * ndir * nc); mt = C.c_double()
        chk(lib.hmg_host_reference(dim, nl, level, sizes, vp(h2l), vp(G), C.byref(mt)))
        nel = i64()
        chk(lib.hmg_host_refined_mesh(dim, nl, level, None, None, C.byref(nel)))
        nodes = np.zeros((nf, dim)); el = np.zeros((nel.value, dim + 1), np.int64)
        chk(lib.hmg_host_refined_mesh(dim, nl, level, vp(nodes), vp(el), C.byref(nel)))
        for kind, nlid in ((0, 4 if dim == 3 else 0), (1, 6 if dim == 3 else 3), (2, dim + 1)):
            for lid in range(nlid):
                n = i64(); chk(lib.hmg_host_interface_rows(dim, nl, level, kind, lid, None, C.byref(n)))
                rows = np.zeros(max(1, n.value), np.int32)
                chk(lib.hmg_host_interface_rows(dim, nl, level, kind, lid, vp(rows), C.byref(n)))
        if nf <= 1000:
            coef = np.random.default_rng(0).random(nc)
            dense = np.zeros((nf, nf)); chk(lib.hmg_host_local_matrix(dim, nl, level, vp(coef), vp(dense)))
            if level >= 2:
                s2 = (i64 * 8)(); chk(lib.hmg_host_reference(dim, nl, level - 1, s2, None, None, None))
                T = np.zeros((nf, int(s2[1]))); chk(lib.hmg_host_transfer_matrix(dim, nl, level, vp(T)))
        x = np.random.default_rng(1).random(nf); y = np.zeros(nf); info = (i64 * 4)()
        coef = np.random.default_rng(2).random(nc)
        for seg in ((3, 5) if dim == 2 else (5,)):
            chk(lib.hmg_host_apply_sweep(dim, nl, level, seg, vp(coef), vp(x), vp(y), info))
    print('reference ok', dim)
# 2. topology / boundary / partition on regular and irregular meshes
import hmgb200 as hmg
rng = np.random.default_rng(3)
for dim, c in ((3, 3), (2, 6), (3, 2), (2, 3)):
    mesh, sigma = hmg.inputs.checkerboard_problem(dim, c)
    for holes in (False, True):
        elems = mesh.elements
        if holes:
            elems = elems[rng.permutation(len(elems))[: 2 * len(elems) // 3]]
        ne, nn = len(elems), mesh.nnodes
        el1 = np.ascontiguousarray(elems + 1, np.int64)
        for kind in range(4):
            nc_, nent = i64(), i64()
            chk(lib.hmg_host_topology(dim, ne, nn, vp(el1), kind, C.byref(nc_), C.byref(nent), None, None, None))
            off = np.zeros(nc_.value + 1, np.int64); e_ = np.zeros(max(1, nent.value), np.int64); l_ = np.zeros(max(1, nent.value), np.int64)
            chk(lib.hmg_host_topology(dim, ne, nn, vp(el1), kind, C.byref(nc_), C.byref(nent), vp(off), vp(e_), vp(l_)))
        cm = np.zeros(ne, np.uint16); it = np.zeros(nn, np.uint8)
        chk(lib.hmg_host_boundary(dim, ne, nn, vp(el1), vp(cm), vp(it)))
        coef = np.zeros((ne, 8)); sg = np.ascontiguousarray(rng.random((ne, dim)) + 1)
        chk(lib.hmg_host_element_coefficients(dim, ne, nn, vp(np.ascontiguousarray(mesh.nodes)), vp(el1), vp(sg), vp(coef), 8))
        for nranks in (1, 2, 3, 8):
            owner = rng.integers(0, nranks, ne).astype(np.int32); owner[:nranks] = np.arange(nranks)
            for rank in range(nranks):
                nel = i64()
                chk(lib.hmg_host_partition_elements(dim, ne, nn, vp(el1), vp(owner), rank, nranks, C.byref(nel), None, None, None, None, None))
                n = nel.value
                l2g = np.zeros(n, np.int64); cmk = np.zeros(n, np.uint16); mult = np.zeros((n, 16), np.uint8)
                nfst = np.zeros(nn, np.int32); ncon = np.zeros(nn, np.uint8)
                chk(lib.hmg_host_partition_elements(dim, ne, nn, vp(el1), vp(owner), rank, nranks, C.byref(nel), vp(l2g), vp(cmk), vp(mult), vp(nfst), vp(ncon)))
                for cut in (0, 1):
                    for kind in range(3):
                        sz = np.zeros(3, np.int64)
                        chk(lib.hmg_host_partition_cells(dim, ne, nn, vp(el1), vp(owner), rank, nranks, kind, cut, vp(sz), None, None, None, None, None))
                        ncell, nent = int(sz[0]), int(sz[1])
                        off = np.zeros(ncell + 1, np.int64); e_ = np.zeros(max(1, nent), np.int64); l_ = np.zeros(max(1, nent), np.int64)
                        sl = np.zeros(max(1, ncell), np.int64); fl = np.zeros(max(1, ncell), np.uint8)
                        chk(lib.hmg_host_partition_cells(dim, ne, nn, vp(el1), vp(owner), rank, nranks, kind, cut, vp(sz), vp(off), vp(e_), vp(l_), vp(sl) if cut else None, vp(fl) if cut else None))
                for kind in range(3):
                    sz = np.zeros(2, np.int64)
                    chk(lib.hmg_host_partition_peers(dim, ne, nn, vp(el1), vp(owner), rank, nranks, kind, vp(sz), None, None, None, None, None))
                    ncell, npe = int(sz[0]), int(sz[1])
                    po = np.zeros(ncell + 1, np.int64); pr = np.zeros(max(1, npe), np.int32); pi = np.zeros(max(1, npe), np.int32); mp = np.zeros(max(1, ncell), np.int32)
                    sw = np.zeros((nranks, 3), np.int64)
                    chk(lib.hmg_host_partition_peers(dim, ne, nn, vp(el1), vp(owner), rank, nranks, kind, vp(sz), vp(po), vp(pr), vp(pi), vp(mp), vp(sw)))
    print('topology/partition ok', dim, c)
print('ALL OK')
