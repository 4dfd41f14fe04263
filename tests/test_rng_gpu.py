"""GPU: the random-number building blocks of the FAST walk kernel (run through the swk_debug_rng test hook of the C-ABI) equal their
numpy restatements, which tests/test_rng.py pins on Random123's known-answer vectors:
  Philox4x32-10 with the kernel's fixed key and Philox2x32-10 (permeability stream): word for word, including the published vectors;
  Box-Muller on MUFU lg2 / sqrt / sin / cos: within 2e-5 (1 + |n|) of the float64 evaluation, and N(0,1) in distribution."""
import ctypes as C

import numpy as np
import pytest

import philox_ref as pr
from test_rng import KAT2

pytestmark = pytest.mark.gpu


def _rng(sw, which, inp):
    inp = np.ascontiguousarray(inp, np.uint32)
    out = np.zeros((inp.shape[0], 8), np.uint32)
    with sw.Engine(0) as e:
        e._ck(e._lib.swk_debug_rng(e._h, which, inp.ctypes.data_as(C.c_void_p), inp.shape[0], out.ctypes.data_as(C.c_void_p)))
    return out


@pytest.fixture(scope="module")
def sw(engine_lib):
    import spinwalk_b200 as sw

    assert engine_lib.swk_device_count() > 0
    return sw


def _counters(n, seed):
    rng = np.random.default_rng(seed)
    c = rng.integers(0, 1 << 32, (n, 4), dtype=np.uint64).astype(np.uint32)
    c[:8] = np.array([[0, 0, 0, 0], [0xFFFFFFFF] * 4, [1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1], [0x80000000] * 4, [0xFFFFFFFF, 0, 0xFFFFFFFF, 0]],
                     np.uint32)
    return c


def test_displacement_stream_is_philox4x32_10(sw):
    c = _counters(100_000, 1)
    assert np.array_equal(_rng(sw, 0, c)[:, :4], pr.philox4x32(c, pr.FIXED_KEY))


def test_permeability_stream_is_philox2x32_10(sw):
    c = _counters(100_000, 2)
    for i, (ctr, key, _) in enumerate(KAT2):  # the published vectors themselves, on the device
        c[8 + i] = (ctr[0], ctr[1], key, 0)
    got = _rng(sw, 1, c)
    want = pr.philox2x32(c[:, :2], c[:, 2])[:, 0]
    assert np.array_equal(got[:, 0], want)
    for i, (_, _, out) in enumerate(KAT2):
        assert int(got[8 + i, 0]) == out[0]
    u = got[:, 1].view(np.float32)
    assert np.array_equal(u, ((want >> 9).astype(np.float64) / 2.0**23).astype(np.float32)) and u.min() >= 0 and u.max() < 1


def test_box_muller_normals(sw):
    from scipy import stats

    n = 1 << 18
    ctr = np.zeros((n, 4), np.uint32)
    ctr[:, 0] = np.arange(n)
    ctr[:, 1] = 10
    ctr[:, 2] = 777
    blocks = _rng(sw, 0, ctr)[:, :4]  # the walk's own chain: counter -> block -> six normals
    got = _rng(sw, 2, blocks)[:, :6].view(np.float32).astype(np.float64)
    want = pr.normals6(blocks)
    assert np.abs(got - want).max() <= 2e-5 * (1 + np.abs(want).max()), np.abs(got - want).max()
    edge = np.array([[0, 0, 0, 0], [0xFFFFFFFF] * 4, [0x1FF, 0x1FF, 0x1FF, 0], [0xFFFFFE00, 0xFFFFFE00, 0xFFFFFE00, 0xFFFFFFFF]], np.uint32)
    g_edge = _rng(sw, 2, edge)[:, :6].view(np.float32).astype(np.float64)
    assert np.isfinite(g_edge).all() and np.abs(g_edge).max() <= 5.66  # u = 1 (radius 0) and u = 2^-23 (radius 5.65) stay finite
    assert np.abs(g_edge - pr.normals6(edge)).max() <= 2e-4
    flat = got.ravel()
    assert abs(flat.mean()) < 4 / np.sqrt(flat.size) and abs(flat.var() - 1) < 4 * np.sqrt(2 / flat.size)
    assert stats.kstest(flat, "norm").statistic < 1.63 / np.sqrt(flat.size)
