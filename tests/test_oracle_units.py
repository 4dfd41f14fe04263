"""CPU tests of the oracle's building blocks.  They restate the reference's own unit tests
(tests/test_kernel.cpp:17-89: sub2ind, xrot / yrot / zrot of unit vectors by 90 deg to 1e-5, relax to 1e-5) and add
known-answer tests for the third-party RNG arithmetic the oracle restates (C++11 [rand.predef]: the 10000th value
of minstd_rand seeded 1 is 399268537; of mt19937 seeded 5489 is 4123659995)."""
import ctypes as C
import math

import numpy as np
import pytest


@pytest.fixture(scope="module")
def lib(oracle):
    l = oracle._lib(oracle.LIB_ORACLE)
    l.swo_sub2ind.restype = C.c_int64
    l.swo_sub2ind.argtypes = [C.c_int64] * 6
    l.swo_erfcinv.restype = C.c_double
    l.swo_erfcinv.argtypes = [C.c_double]
    l.swo_step_sigma.restype = C.c_double
    l.swo_step_sigma.argtypes = [C.c_double, C.c_int32]
    l.swo_tesla_to_deg_per_step.restype = C.c_float
    l.swo_tesla_to_deg_per_step.argtypes = [C.c_float, C.c_int32]
    return l


def _rot(lib, name, m0, theta_deg):
    s, c = np.float32(math.sin(math.radians(theta_deg))), np.float32(math.cos(math.radians(theta_deg)))
    a = np.asarray(m0, np.float32)
    out = np.zeros(3, np.float32)
    getattr(lib, name)(C.c_float(s), C.c_float(c), a.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return out


def test_sub2ind_3d_row_major(lib):  # tests/test_kernel.cpp:17-23
    assert lib.swo_sub2ind(1, 2, 3, 10, 10, 10) == 1 * 10 * 10 + 2 * 10 + 3
    assert lib.swo_sub2ind(4, 0, 7, 5, 6, 9) == 4 * 9 * 6 + 7


def test_xrot(lib):  # tests/test_kernel.cpp:36-48
    assert np.allclose(_rot(lib, "swo_xrot", [0, 0, 1], 90.0), [0, -1, 0], atol=1e-5)


def test_yrot(lib):  # tests/test_kernel.cpp:50-62
    assert np.allclose(_rot(lib, "swo_yrot", [1, 0, 0], 90.0), [0, 0, -1], atol=1e-5)


def test_zrot(lib):  # tests/test_kernel.cpp:64-76
    assert np.allclose(_rot(lib, "swo_zrot", [1, 0, 0], 90.0), [0, 1, 0], atol=1e-5)


def test_relax(lib):  # tests/test_kernel.cpp:78-89
    m0 = np.array([1.0, 0.5, -0.5], np.float32)
    out = np.zeros(3, np.float32)
    lib.swo_relax(C.c_float(0.9), C.c_float(0.8), m0.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    assert np.allclose(out, [1.0 * 0.8, 0.5 * 0.8, 1.0 + 0.9 * (-0.5 - 1.0)], atol=1e-5)


@pytest.mark.parametrize("ph,expect", [(0.0, "x+"), (180.0, "x-"), (90.0, "y+"), (-90.0, "y-"), (270.0, "y-")])
def test_xrot_withphase_fast_paths(lib, ph, expect):  # kernels.cuh:160-181
    th = 37.0
    s, c = np.float32(math.sin(math.radians(th))), np.float32(math.cos(math.radians(th)))
    m0 = np.array([0.3, -0.4, 0.8], np.float32)
    out = np.zeros(3, np.float32)
    lib.swo_xrot_withphase(C.c_float(s), C.c_float(c), C.c_float(ph), m0.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    name, sign = {"x+": ("swo_xrot", 1), "x-": ("swo_xrot", -1), "y+": ("swo_yrot", 1), "y-": ("swo_yrot", -1)}[expect]
    assert np.array_equal(out, _rot(lib, name, m0, sign * th))


def test_xrot_withphase_general_is_axis_rotation(lib):  # kernels.cuh:183-188: Rz(ph) Rx(th) Rz(-ph)
    th, ph = 50.0, 33.5
    s, c = np.float32(math.sin(math.radians(th))), np.float32(math.cos(math.radians(th)))
    m0 = np.array([0.3, -0.4, 0.8], np.float32)
    out = np.zeros(3, np.float32)
    lib.swo_xrot_withphase(C.c_float(s), C.c_float(c), C.c_float(ph), m0.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    ax = np.array([math.cos(math.radians(ph)), math.sin(math.radians(ph)), 0.0])
    t = math.radians(th)
    v = m0.astype(np.float64)
    rod = v * math.cos(t) + np.cross(ax, v) * math.sin(t) + ax * ax.dot(v) * (1 - math.cos(t))
    assert np.allclose(out, rod, atol=1e-6)


def _stream(lib, fn, sps, n):
    out = np.zeros(n, np.float32)
    getattr(lib, fn)(C.c_uint64(sps), C.c_uint32(n), out.ctypes.data_as(C.c_void_p))
    return out


def test_minstd_known_answer(lib):
    """uniform = (x-1)/2^31 of the stream after seeding with `s` and discarding `s` values.  With s = 1 the stream is
    x_2, x_3, ...; the standard's check value is x_10000 = 399268537, i.e. element 9998 of that stream."""
    u = _stream(lib, "swo_minstd_uniforms", 1, 9999)
    assert u[9998] == np.float32(399268537 - 1) / np.float32(2147483648.0)
    assert u[0] == np.float32(48271 * 48271 % 2147483647 - 1) / np.float32(2147483648.0)


def test_mt19937_known_answer(lib):
    """seed 5489, discard 5489, then the (10000-5489)th output must be the standard's 10000th value 4123659995."""
    u = _stream(lib, "swo_mt_uniforms", 5489, 10000 - 5489)
    assert u[-1] == np.float32(np.float32(4123659995) / np.float32(4294967296.0))


def test_erfcinv_inverts_libm_erfc(lib):
    for x in [4.7e-10, 1e-6, 0.0034, 0.1, 0.5, 0.999, 1.0, 1.5, 1.9966]:
        y = lib.swo_erfcinv(x)
        assert math.erfc(y) == pytest.approx(x, rel=5e-15, abs=1e-300)


@pytest.mark.parametrize("fn", ["swo_minstd_normals", "swo_mt_normals"])
def test_normal_streams_are_standard_normal(lib, fn):
    z = _stream(lib, fn, 12345, 200000).astype(np.float64)
    assert abs(z.mean()) < 4 / math.sqrt(z.size)
    assert z.var() == pytest.approx(1.0, abs=0.02)
    assert (z**4).mean() == pytest.approx(3.0, abs=0.15)
    assert np.mean(np.abs(z) < 1.0) == pytest.approx(0.6827, abs=0.005)


def test_prepare_formulas(lib):  # simulation_parameters.cuh:227-245, monte_carlo.cu:241
    assert lib.swo_step_sigma(1e-9, 50) == pytest.approx(1e-3 * math.sqrt(2 * 1e-9 * 50), rel=1e-15)  # 0.316 um
    k = lib.swo_tesla_to_deg_per_step(9.4, 50)
    assert k == pytest.approx(9.4 * 50e-6 * 267515315.0 * 57.2957795130823, rel=1e-6)
