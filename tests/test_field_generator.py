"""The cuFFT coefficient-field generator (tools/generate_st1_field.jl:86-120) against the numpy
restatement of the recipe that produces the inputs of BASELINE.json configs[4] (homogenization.jl_b200/inputs.py)."""
import numpy as np
import pytest

import hmgb200 as hmg


def test_philox_restatement_known_answer():
    """Philox4x32-10 known-answer vector of the Random123 distribution (counter = key = 0 -> 6627e8d5 e169c58d
    bc57ac4c 9b00dbd8), through the same integer pipeline philox_normal uses, and sane moments of the stream."""
    M = np.uint64(0xFFFFFFFF)
    c = [np.zeros(1, np.uint64) for _ in range(4)]
    k0 = k1 = np.uint64(0)
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c[0]
        p1 = np.uint64(0xCD9E8D57) * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ k0, p1 & M, (p0 >> np.uint64(32)) ^ c[3] ^ k1, p0 & M]
        k0 = (k0 + np.uint64(0x9E3779B9)) & M
        k1 = (k1 + np.uint64(0xBB67AE85)) & M
    assert [int(v[0]) for v in c] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    g = hmg.inputs.philox_normal(1 << 16, 7)
    assert abs(g.mean()) < 0.02 and abs(g.std() - 1.0) < 0.02
    assert not np.array_equal(g, hmg.inputs.philox_normal(1 << 16, 8))


@pytest.mark.gpu
@pytest.mark.parametrize("dim,n", [(2, 16), (2, 64), (3, 8), (3, 16)])
@pytest.mark.parametrize("normalize", [True, False])
def test_device_field_matches_numpy_recipe(dim, n, normalize):
    rng = np.random.default_rng(5)
    noise = rng.standard_normal((n,) * dim)
    got = hmg.inputs.random_field_cells_device(dim, n, alpha=0.8, p=1.5, normalize=normalize, noise=noise)
    k = np.meshgrid(*[np.fft.fftfreq(n) * n for _ in range(dim)], indexing="ij")
    kn = np.sqrt(sum(q * q for q in k))
    G = np.real(np.fft.ifftn(np.fft.fftn(noise) * (1.0 + kn) ** (-1.5)))
    if normalize:
        G = G / G.std()
    expect = np.exp(0.8 * np.abs(G))
    assert got.shape == (n,) * dim + (dim,)
    for d in range(dim):
        assert np.max(np.abs(got[..., d] - expect) / expect) <= 1e-12


@pytest.mark.gpu
def test_device_field_equals_inputs_random_field_cells():
    """Same noise -> the very array inputs.random_field_cells feeds the C5 sweep."""
    n, dim = 32, 2
    noise = np.random.default_rng(2).standard_normal((n,) * dim)
    expect = hmg.inputs.random_field_cells(dim, n, seed=2)
    got = hmg.inputs.random_field_cells_device(dim, n, noise=noise)
    assert np.max(np.abs(got - expect) / expect) <= 1e-12


@pytest.mark.gpu
def test_device_noise_is_the_restated_philox_stream():
    n, dim = 16, 3
    noise = hmg.inputs.philox_normal(n ** dim, 11).reshape((n,) * dim)
    a = hmg.inputs.random_field_cells_device(dim, n, seed=11)                 # noise drawn on the device
    b = hmg.inputs.random_field_cells_device(dim, n, seed=11, noise=noise)    # the numpy restatement fed in
    assert np.max(np.abs(a - b) / b) <= 1e-10
    assert a.min() >= 1.0                                                     # exp(alpha |G|) >= 1


def test_generator_rejects_odd_extents_without_touching_the_gpu_path():
    import ctypes as C
    lib = hmg.load()
    ns = (C.c_int * 3)(5, 4, 1)
    out = np.zeros(20)
    assert lib.hmg_generate_field(2, ns, 1, 1.0, 1.5, 1, None, out.ctypes.data_as(C.c_void_p), 0) != 0
    assert b"even" in lib.hmg_last_error() or b"CUDA" in lib.hmg_last_error()


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(8, 12), (6, 4, 10)])
def test_device_field_with_different_extents(shape):
    """The C entry point takes one extent per axis (tools/generate_st1_field.jl:86-90 does too): the wrapped frequency
    index of every axis uses that axis' own half length."""
    import ctypes as C
    dim = len(shape)
    noise = np.random.default_rng(9).standard_normal(shape)
    lib = hmg.load()
    ns = (C.c_int * 3)(*(list(shape) + [1] * (3 - dim)))
    out = np.empty(shape)
    hmg._lib.check(lib.hmg_generate_field(dim, ns, 0, 0.5, 1.5, 1, noise.ctypes.data_as(C.c_void_p),
                                          out.ctypes.data_as(C.c_void_p), 0))
    k = np.meshgrid(*[np.fft.fftfreq(n) * n for n in shape], indexing="ij")
    kn = np.sqrt(sum(q * q for q in k))
    G = np.real(np.fft.ifftn(np.fft.fftn(noise) * (1.0 + kn) ** (-1.5)))
    expect = np.exp(0.5 * np.abs(G / G.std()))
    assert np.max(np.abs(out - expect) / expect) <= 1e-12
