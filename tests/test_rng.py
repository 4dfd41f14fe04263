"""CPU: the numpy restatements of the FAST mode's generators (tests/philox_ref.py) reproduce the known-answer vectors that the
Random123 library ships for philox4x32-10 and philox2x32-10 (kat_vectors), and the Box-Muller layout gives standard normals."""
import numpy as np

import philox_ref as pr

F = 0xFFFFFFFF
KAT4 = [((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
        ((F, F, F, F), (F, F), (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
        ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0), (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1))]
KAT2 = [((0, 0), 0, (0xFF1DAE59, 0x6CD10DF2)), ((F, F), F, (0x2C3F628B, 0xAB4FD7AD)), ((0x243F6A88, 0x85A308D3), 0x13198A2E, (0xDD7CE038, 0xF62A4C12))]


def test_philox4x32_10_known_answers():
    for ctr, key, want in KAT4:
        assert tuple(int(v) for v in pr.philox4x32(np.array([ctr], np.uint32), key)[0]) == want


def test_philox2x32_10_known_answers():
    for ctr, key, want in KAT2:
        assert tuple(int(v) for v in pr.philox2x32(np.array([ctr], np.uint32), key)[0]) == want


def test_box_muller_layout_gives_standard_normals():
    """sequential counters -> Philox blocks -> six normals each: moments and a Kolmogorov-Smirnov distance against N(0,1); the
    two attempts fed by one block are uncorrelated."""
    from scipy import stats

    n = 1 << 18
    ctr = np.zeros((n, 4), np.uint32)
    ctr[:, 0] = np.arange(n)
    ctr[:, 2] = 12345
    z = pr.normals6(pr.philox4x32(ctr, pr.FIXED_KEY))
    assert np.abs(z).max() <= np.sqrt(2 * np.log(2.0**23)) + 1e-9  # 5.65: what bounds the fixed-point step (walk_fast.cuh)
    flat = z.ravel()
    assert abs(flat.mean()) < 4 / np.sqrt(flat.size) and abs(flat.var() - 1) < 4 * np.sqrt(2 / flat.size)
    assert stats.kstest(flat, "norm").statistic < 1.63 / np.sqrt(flat.size)  # 1 % critical value
    c = np.corrcoef(z.T)
    assert np.abs(c - np.eye(6)).max() < 5 / np.sqrt(n)
