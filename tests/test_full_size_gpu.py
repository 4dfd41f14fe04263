"""BASELINE.json configs[1] at FULL size (600^3 phantom with field map, 1e7 spins x 50 FoV scales x 800 steps = 4e11 spin-steps)
through size-independent properties — the oracle cannot run this in reasonable time (CPU reference: ~1e8 spin-steps/s):

  * conservation: every (scale) counts exactly S spins at the echo, none lost (impermeable walls never trap a spin for 1e4 draws);
  * impermeability: P_XY = identity => a spin never changes substrate, so the per-substrate count at the echo equals the
    occupancy of the START voxels, for every one of the 50 scales, as exact integers;
  * shard invariance at scale: two engines on halves of the spins give the sums of the single run (only the FP32/FP64
    association order of the ensemble sums differs);
  * physics sanity: |S| <= 1 everywhere and the spin-echo signal at the largest scales is within 2 % of exp(-TE/T2)
    (vessels far apart, static dephasing refocused by the 180)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_c2_full_size_properties(engine_lib):
    import torch

    import bench
    import spinwalk_b200 as sw

    cfg_kw, ph, _ = bench.workload("c2", None, None)
    cfg = sw.SimConfig(**cfg_kw)
    S, K = cfg.n_spins, len(cfg.scales)
    mask2, fm2, fov = bench.make_phantom_2d(ph)
    n = ph["n"]
    dev = torch.device("cuda", 0)
    mask_d = torch.from_numpy(mask2).to(dev)[:, :, None].expand(n, n, n).contiguous()
    fm_d = torch.from_numpy(fm2).to(dev)[:, :, None].expand(n, n, n).contiguous()
    xyz0 = bench.make_positions(S, fov, cfg.seed)
    # occupancy of the start voxels (the kernel truncates position * n / fov like kernels.cuh:53-60)
    vox = np.minimum((xyz0.astype(np.float64) * (n / np.asarray(fov, np.float64))).astype(np.int64), n - 1)
    start_sub = mask2[vox[:, 0], vox[:, 1]]
    n_start = np.bincount(start_sub, minlength=2).astype(np.float64)

    with sw.Engine(0) as e:
        e.set_phantom(mask_d, fm_d, fov)
        del mask_d, fm_d
        e.set_sequence(cfg)
        e.set_spins(xyz0)
        st = e.run_device(mode=sw.MODE_FAST, flags=0)
        full = e.sums()
        halves = []
        for first, cnt in ((0, S // 2), (S // 2, S - S // 2)):
            e.set_spins(xyz0[first:first + cnt], None, first)
            e.run_device(mode=sw.MODE_FAST, flags=0)
            halves.append(e.sums())
    assert st["lost"] == 0 and st["steps"] == S * K * 800
    assert full.shape == (K, 1, 2, 4)
    assert np.array_equal(full[:, 0, :, 3], np.broadcast_to(n_start, (K, 2))), "a spin changed substrate through an impermeable wall"
    assert np.array_equal(halves[0][..., 3] + halves[1][..., 3], full[..., 3])
    assert np.allclose(halves[0][..., :3] + halves[1][..., :3], full[..., :3], rtol=0, atol=2.0)  # sums of 1e7 O(1) terms
    tot = full.sum(axis=2)[:, 0]
    sig = np.hypot(tot[:, 0], tot[:, 1]) / S
    assert (sig <= 1.0 + 1e-6).all() and (sig > 0.3).all()
    assert abs(sig[-1] / np.exp(-20.0 / 41.0) - 1.0) < 0.02, sig[-5:]
