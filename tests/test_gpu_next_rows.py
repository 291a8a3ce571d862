"""The callers and data formats either side of the V-cycle on the device:
the column-prefix copy between contexts (domain shrink, src/examples/homogenized_coefficients.jl:54-60, 297-339),
the row-prefix download and the VTK export of a coarse-level slice (:81-87), and the coarse solve through the
64-bit factorisation path (large base meshes; forced on a small one here)."""
import os

import numpy as np
import pytest

import hmgb200 as hmg
from parity_common import Pair, relerr
from oracle import driver as od

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dim,c,levels", [(2, 6, 5), (3, 3, 4), (3, 2, 5)], ids=["tri-c6-L5", "tet-c3-L4", "tet-c2-L5"])
def test_download_rows_is_the_row_prefix(dim, c, levels):
    mesh, sigma = hmg.inputs.checkerboard_problem(dim, c)
    g = hmg.ImplicitFineGrid(mesh, levels, sigma)
    try:
        rng = np.random.default_rng(3)
        for level in (levels, max(1, levels - 1)):
            x = np.asfortranarray(rng.random((g.nf(level), mesh.nelements)))
            st = g.state(level)
            st.x.set(x)
            for k in range(1, level + 1):
                nrows = g.nf(k)
                assert np.array_equal(st.x.get_rows(nrows), x[:nrows])          # bit-exact data movement
            assert st.x.get_rows(0).shape == (0, mesh.nelements)
            assert np.array_equal(st.x.get_rows(7 if g.nf(level) >= 7 else 1), x[:7 if g.nf(level) >= 7 else 1])
            with pytest.raises(ValueError):
                st.x.get_rows(g.nf(level) + 1)
    finally:
        g.close()


@pytest.mark.parametrize("dim,c,levels", [(2, 7, 4), (3, 3, 3), (3, 2, 5)], ids=["tri-c7-L4", "tet-c3-L3", "tet-c2-L5"])
def test_copy_columns_from_is_the_column_prefix(dim, c, levels):
    """shrink_level_state: l.x[:, OneTo(n)] across contexts, for prefixes that end inside, at the end of and at the
    start of a unit of 32 columns; the padding columns of the destination's last unit stay zero (dots run over them)."""
    mesh, sigma = hmg.inputs.checkerboard_problem(dim, c, ordered=True)
    big = hmg.ImplicitFineGrid(mesh, levels, sigma)
    try:
        x = np.asfortranarray(np.random.default_rng(5).random((big.nf(levels), mesh.nelements)))
        big.state(levels).x.set(x)
        ne = mesh.nelements
        for n_keep in sorted({ne, ne - 1, max(1, ne // 2), min(ne, 64), min(ne, 33), 5}):
            used = np.unique(mesh.elements[:n_keep])            # monotone renumbering keeps the elements sorted
            sub = hmg.Mesh(mesh.nodes[used], np.searchsorted(used, mesh.elements[:n_keep]))
            small = hmg.ImplicitFineGrid(sub, levels, np.ascontiguousarray(sigma[:n_keep]))
            try:
                dst = small.state(levels)
                dst.v.fill(7.0)
                dst.v.copy_columns_from(big.state(levels).x)
                assert np.array_equal(dst.v.get(), x[:, :n_keep])
                # sum over ALL stored entries including the padding: must equal the sum over the kept columns
                got = hmg.dot(small, dst.v, dst.v)
                assert abs(got - float(np.sum(x[:, :n_keep] ** 2))) <= 1e-12 * got
            finally:
                small.close()
        with pytest.raises(hmg.HmgError):
            big.state(levels).x.copy_columns_from(big.state(levels).p)        # same context
    finally:
        big.close()


def test_driver_device_shrink_equals_host_shrink(tmp_path):
    """checkerboard_homogenization with a domain shrink (the outer loop only shrinks from n = 5 on: 56 -> 55 cells of
    radius after the first step): moving the column prefix on the device gives the very same history as the download /
    upload path, and `save` writes the files of export_domain / export_unknown."""
    n, dim, refinements = 5, 2, 1
    rng = np.random.default_rng(42)
    radius = od.compute_box_radius(0, n) + od.compute_boundary_layer(1.0, n)
    cells = np.where(rng.random((2 * radius,) * dim + (dim,)) < 0.5, 1.0, 9.0)
    ne0 = 2 * (2 * radius) ** 2
    nf = hmg.inputs.nf_of_level(dim, refinements + 1)
    x0 = np.asfortranarray(rng.random((nf, ne0)))
    kw = dict(refinements=refinements, tolerance=1e-3, sigma_cells=cells, x0=x0)
    s_host, h_host = hmg.driver.checkerboard_homogenization(n, dim, shrink="host", **kw)
    prefix = str(tmp_path) + os.sep
    s_dev, h_dev = hmg.driver.checkerboard_homogenization(n, dim, shrink="device", save=2, save_prefix=prefix, **kw)
    assert len(h_host) == 2                                   # the outer loop did shrink once
    assert s_dev == s_host and h_dev == h_host                # bit-identical
    mesh, _, cd = hmg.vtk.read_vtu(prefix + "checkerboard.vtu")
    assert mesh.nelements == ne0 and cd["a"].shape == (ne0, dim)
    sizes = []
    for k in range(2):
        full, pd, _ = hmg.vtk.read_vtu(prefix + f"ahom_{k}.vtu")
        assert full.nnodes == pd["v"].shape[0] and full.nnodes % nf == 0
        assert np.all(np.isfinite(pd["v"])) and np.any(pd["v"] != 0.0)
        sizes.append(full.nnodes // nf)
    assert sizes[0] == ne0 and sizes[1] < ne0                 # the second step ran on the shrunken mesh


def test_export_unknown_is_the_coarse_level_slice(tmp_path):
    pair = Pair(3, 2, 4, lam=0.7)
    try:
        L = pair.levels
        x = pair.rand(L)
        st = pair.g.state(L)
        st.x.set(x)
        for level in (1, 3):
            path = hmg.vtk.export_unknown(pair.g, st.x, 5, level, str(tmp_path / f"ahom_5_l{level}"))
            full, pd, _ = hmg.vtk.read_vtu(path)
            nfl = pair.g.nf(level)
            assert np.array_equal(pd["v"], x[:nfl].reshape(-1, order="F"))      # x[1:Nf(level), :][:]
            assert full.nnodes == nfl * pair.mesh.nelements
    finally:
        pair.close()


@pytest.mark.parametrize("dim,c,levels", [(2, 9, 3), (3, 4, 3)], ids=["tri-c9-L3", "tet-c4-L3"])
def test_coarse_solver_64bit_path_matches_default(dim, c, levels):
    """The factorisation path for base meshes beyond potri's 32-bit range (>= 46 340 interior nodes), forced here on a
    small mesh: same residual history as the default path to rounding, and both match the oracle's exact coarse solve."""
    def history(force):
        pair = Pair(dim, c, levels, lam=0.7)
        try:
            if force:
                os.environ["HMG_COARSE_64BIT"] = "1"
            try:
                bl = hmg.BaseLevel(pair.g)
            finally:
                os.environ.pop("HMG_COARSE_64BIT", None)
            L = pair.levels
            st = pair.g.state(L)
            st.x.set(pair.rand(L))
            hmg.broadcast_interfaces(st.x, pair.g, L)
            hmg.apply_constraint(st.x, L, pair.g)
            st.b.set(pair.rand(L))
            return hmg.vcycles(pair.g, bl, L, 3, 4), st.x.get()
        finally:
            pair.close()
    h32, x32 = history(False)
    h64, x64 = history(True)
    assert np.all(h32 > 0) and np.allclose(h64, h32, rtol=1e-10, atol=0)
    assert relerr(x64, x32) <= 1e-10
