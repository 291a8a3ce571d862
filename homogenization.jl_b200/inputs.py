"""Synthetic inputs of the benchmark configurations.

Input generators only -- the base mesh of the checkerboard problem, the per-element conductivity
and the initial guess.  They reproduce the *input format* of the reference
(src/tri/generate_grid.jl:6-35, src/tet/generate_grid.jl:6-45,
src/examples/homogenized_coefficients.jl:21-28, 485-503); randomness comes from
``numpy.random.default_rng(seed)`` because the reference's is unseeded.
"""
import numpy as np

from .api import Mesh


def hypercube(dim, n, origin=None, scale=1.0):
    """hypercube(Tri|Tet, n; origin, scale): n^dim unit cells, 2 triangles / 6 tetrahedra each.

    Node q = (a, b[, c]) in x-outer order has coordinates scale*(a, b[, c]) + origin; the
    reference looks corner nodes up through a column-major index table, which swaps the roles
    of the axes in the connectivity -- kept, since element order and orientation depend on it."""
    n1 = n + 1
    origin = np.ones(dim) if origin is None else np.asarray(origin, dtype=np.float64)
    grid = np.indices((n1,) * dim).reshape(dim, -1).T
    nodes = scale * grid + origin
    cells = np.indices((n,) * dim).reshape(dim, -1).T           # last axis fastest
    stride = n1 ** np.arange(dim)                                # column-major lookup table

    def corner(*bits):
        return (cells + np.asarray(bits)) @ stride

    if dim == 2:
        c = {b: corner(b & 1, b >> 1) for b in range(4)}         # c[b]: +x = bit 0, +y = bit 1
        tris = [(c[0], c[1], c[2]), (c[1], c[2], c[3])]
        el = np.stack([np.stack(t, axis=1) for t in tris], axis=1).reshape(-1, 3)
    else:
        c = {b: corner(b & 1, (b >> 1) & 1, b >> 2) for b in range(8)}
        n1_, n2, n3, n4, n5, n6, n7, n8 = (c[b] for b in range(8))
        tets = [(n1_, n2, n3, n7), (n1_, n2, n5, n7), (n2, n4, n3, n7),
                (n2, n4, n7, n8), (n2, n6, n5, n7), (n2, n6, n7, n8)]
        el = np.stack([np.stack(t, axis=1) for t in tets], axis=1).reshape(-1, 4)
    return Mesh(nodes, np.sort(el, axis=1))


def element_centers(mesh):
    p = mesh.nodes[mesh.elements]
    s = p[:, 0, :].copy()
    for i in range(1, p.shape[1]):
        s = s + p[:, i, :]
    return s / p.shape[1]


def order_by_magnitude(mesh):
    """order_nodes_and_elements_by_magnitude (src/examples/homogenized_coefficients.jl:21-28)."""
    inf = lambda a: np.max(np.abs(a), axis=1)
    I = np.argsort(inf(mesh.nodes), kind="stable")
    J = np.empty_like(I)
    J[I] = np.arange(len(I))
    out = Mesh(mesh.nodes[I], np.sort(J[mesh.elements], axis=1))
    order = np.argsort(inf(element_centers(out)), kind="stable")
    out.elements = np.ascontiguousarray(out.elements[order])
    return out


def checkerboard_cells(dim, n, seed=1):
    """generate_conductivity: per unit cell a diagonal tensor with entries 1 or 9, p = 1/2
    (src/examples/homogenized_coefficients.jl:485-488)."""
    rng = np.random.default_rng(seed)
    return np.where(rng.random((n,) * dim + (dim,)) < 0.5, 1.0, 9.0)


def random_field_cells(dim, n, seed=2, alpha=1.0, p=1.5):
    """Isotropic log-normal-like field exp(alpha |G|), G = white noise filtered by (1+|k|)^-p in
    Fourier space and normalised to unit variance (recipe of tools/generate_st1_field.jl:41-110)."""
    rng = np.random.default_rng(seed)
    noise = rng.standard_normal((n,) * dim)
    k = np.meshgrid(*[np.fft.fftfreq(n) * n for _ in range(dim)], indexing="ij")
    kn = np.sqrt(sum(q * q for q in k))
    G = np.real(np.fft.ifftn(np.fft.fftn(noise) * (1.0 + kn) ** (-p)))
    G = G / G.std()
    g = np.exp(alpha * np.abs(G))
    return np.repeat(g[..., None], dim, axis=-1)


def philox_normal(n, seed):
    """The standard-normal stream of the device generator (csrc/field.cu: Philox4x32-10, key = seed, counter = cell
    index, Box-Muller on two 53-bit uniforms), restated with numpy integer arithmetic."""
    i = np.arange(n, dtype=np.uint64)
    c = [(i & np.uint64(0xFFFFFFFF)), (i >> np.uint64(32)), np.zeros(n, np.uint64), np.zeros(n, np.uint64)]
    k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
    M = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c[0]
        p1 = np.uint64(0xCD9E8D57) * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ k0, p1 & M, (p0 >> np.uint64(32)) ^ c[3] ^ k1, p0 & M]
        k0 = (k0 + np.uint64(0x9E3779B9)) & M
        k1 = (k1 + np.uint64(0xBB67AE85)) & M
    u1 = ((((c[0] << np.uint64(32)) | c[1]) >> np.uint64(11)).astype(np.float64) + 0.5) * 2.0 ** -53
    u2 = ((((c[2] << np.uint64(32)) | c[3]) >> np.uint64(11)).astype(np.float64) + 0.5) * 2.0 ** -53
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


def random_field_cells_device(dim, n, seed=2, alpha=1.0, p=1.5, normalize=True, noise=None, device=0):
    """The same field through the library's cuFFT generator (hmg_generate_field; tools/generate_st1_field.jl:86-120).
    ``noise`` (shape (n,)*dim): the caller's white noise; None: drawn on the device (``philox_normal`` restates it)."""
    import ctypes as C
    from ._lib import load, check
    lib = load()
    ns = (C.c_int * 3)(*([n] * dim + [1] * (3 - dim)))
    out = np.empty((n,) * dim, dtype=np.float64)
    nz = None
    if noise is not None:
        nz = np.ascontiguousarray(noise, dtype=np.float64)
        assert nz.shape == (n,) * dim
    check(lib.hmg_generate_field(dim, ns, int(seed), float(alpha), float(p), int(bool(normalize)),
                                 nz.ctypes.data_as(C.c_void_p) if nz is not None else None,
                                 out.ctypes.data_as(C.c_void_p), int(device)))
    return np.repeat(out[..., None], dim, axis=-1)


def conductivity_per_element(mesh, cells, offset):
    """conductivity_per_element (src/examples/homogenized_coefficients.jl:494-503):
    the cell of an element is trunc(centroid + offset), 1-based."""
    idx = np.trunc(element_centers(mesh) + np.asarray(offset, dtype=np.float64)).astype(np.int64) - 1
    return np.ascontiguousarray(cells[tuple(idx[:, d] for d in range(mesh.dim))])


def checkerboard_problem(dim, c, field="checkerboard", seed=1, ordered=False):
    """Base mesh of c^dim unit cells centred at the origin + per-element sigma.
    Returns (mesh, sigma (Ne, dim))."""
    half = c / 2.0
    mesh = hypercube(dim, c, origin=(-half,) * dim)
    if ordered:
        mesh = order_by_magnitude(mesh)
    cells = checkerboard_cells(dim, c, seed) if field == "checkerboard" else random_field_cells(dim, c, seed)
    sigma = conductivity_per_element(mesh, cells, (half + 1.0,) * dim)
    return mesh, sigma


def nf_of_level(dim, level):
    m = 1 << (level - 1)
    return (m + 1) * (m + 2) // 2 if dim == 2 else (m + 1) * (m + 2) * (m + 3) // 6


def spatial_partition(mesh, nranks):
    """Owner rank of every coarse element: whole cells by spatial blocks:
    2 -> halves, 4 -> 2x2(x1), 8 -> 2x2x2 (3D) / 4x2 (2D)."""
    ctr = element_centers(mesh)
    lo, hi = mesh.nodes.min(axis=0), mesh.nodes.max(axis=0)
    dim = mesh.dim
    splits = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2) if dim == 3 else (4, 2, 1)}[nranks]
    owner = np.zeros(mesh.nelements, dtype=np.int32)
    mult = 1
    for d in range(dim):
        s = splits[d]
        ncell = int(round(hi[d] - lo[d]))
        cell = np.clip(np.floor(ctr[:, d] - lo[d]).astype(np.int64), 0, ncell - 1)
        owner += ((cell * s) // ncell * mult).astype(np.int32)
        mult *= s
    return owner
