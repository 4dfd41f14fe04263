"""GPU parity tests of the phantom generator, all through the C-ABI of include/spinwalk_phantom.h.

Bar: BIT-EXACT.  Mask, field map (compared as uint32 bit patterns), shape list and volume fraction must equal the oracle
(oracle/phantom_oracle.c, itself pinned on the unmodified reference generator, tests/test_phantom_oracle.py) and the committed
reference goldens (tests/golden/phantom/*.npz: SHA-256 of the reference's own byte strings)."""
import hashlib
import os

import numpy as np
import pytest

from phantom_cases import CASES, TWOPOOLS

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "phantom")


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def pp(oracle):
    from oracle import pyphantom

    return pyphantom


@pytest.fixture(scope="module")
def pg(engine_lib):
    from spinwalk_b200 import phantom_gen

    return phantom_gen


def spec_of(pg, kw):
    return pg.PhantomSpec(shape=kw["shape"], fov_um=kw["fov_um"], resolution=kw["resolution"], dchi=kw.get("dchi", 0.11e-6), oxy_level=kw.get("Y", 0.78),
                          radius_um=kw.get("radius_um", 8.0), volume_fraction=kw.get("volume_fraction", 4.0), orientation_deg=kw.get("orientation_deg", 90.0),
                          seed=kw.get("seed", 0))


def assert_same(mask, fm, st, ph):
    assert mask.dtype == np.uint8 and np.array_equal(mask, ph.mask)
    if ph.fieldmap is None:
        assert fm is None
    else:
        assert fm.dtype == np.float32
        bad = np.flatnonzero(fm.view(np.uint32).ravel() != ph.fieldmap.view(np.uint32).ravel())
        assert bad.size == 0, f"{bad.size} field-map voxels differ, first at {np.unravel_index(bad[0], fm.shape)}: {fm.ravel()[bad[0]]!r} vs {ph.fieldmap.ravel()[bad[0]]!r}"
    assert st["n_shapes"] == len(ph.shapes)
    assert np.float32(st["volume_fraction"]) == np.float32(ph.bvf)


@pytest.mark.parametrize("name", sorted(CASES))
def test_generator_bit_exact_vs_oracle_and_reference_golden(pg, pp, name):
    kw = CASES[name]
    mask, fm, fov, st = pg.generate(spec_of(pg, kw))
    assert_same(mask, fm, st, pp.oracle(**kw))
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    assert digest(mask) == str(gold["mask_sha256"])
    if fm is not None:
        assert digest(fm) == str(gold["fieldmap_sha256"])
    assert np.float32(st["volume_fraction"]) == gold["bvf"]
    assert np.array_equal(fov, np.full(3, np.float32(kw["fov_um"]) * np.float32(1e-6), np.float32))  # phantom_base.cpp:63
    assert st["n_launches"] >= (0 if kw["shape"] == TWOPOOLS else 1)


@pytest.mark.parametrize("name", ["cyl_bold", "cyl_random_oblique", "cyl_many"])
def test_cylinder_exact_kernel_gives_the_same_bits(pg, pp, name, monkeypatch):
    """Forces every column outside a cylinder through cyl_exact_kernel (voxel-by-voxel, z residual in place)."""
    monkeypatch.setenv("SWK_PHANTOM_FORCE_EXACT", "1")
    kw = CASES[name]
    mask, fm, _, st = pg.generate(spec_of(pg, kw))
    assert st["exact_columns"] > 0 and st["n_launches"] == 3
    assert_same(mask, fm, st, pp.oracle(**kw))


def test_generator_writes_into_device_tensors(pg, pp):
    import torch

    for name in ("cyl_random_oblique", "sph_fixed"):
        kw = CASES[name]
        n = kw["resolution"]
        mask_d = torch.full((n, n, n), 7, dtype=torch.uint8, device="cuda")
        fm_d = torch.full((n, n, n), 3.0, dtype=torch.float32, device="cuda")
        _, _, _, st = pg.generate(spec_of(pg, kw), out=(mask_d, fm_d))
        assert_same(mask_d.cpu().numpy(), fm_d.cpu().numpy(), st, pp.oracle(**kw))


def test_engine_resident_phantom_feeds_the_walk(pg, pp):
    """swk_generate_phantom: the generated phantom never visits the host, and the walk on it equals the walk on the same arrays
    uploaded through swk_set_phantom (bitwise: same voxels, same spins, same RNG)."""
    import spinwalk_b200 as sw

    kw = CASES["cyl_bold"]
    spec = spec_of(pg, kw)
    ph = pp.oracle(**kw)
    cfg = sw.SimConfig(n_spins=4096, seed=5, TR_us=4000, timestep_us=50, TE_us=[2000], RF_T_us=[0], RF_FA_deg=[90.0], RF_PH_deg=[0.0],
                       scales=[0.5, 1.0, 2.0])
    fov = np.full(3, np.float32(kw["fov_um"]) * np.float32(1e-6), np.float32)
    rng = np.random.default_rng(1)
    xyz0 = (rng.random((4096, 3), dtype=np.float32) * 0.98 + 0.01) * fov
    outs = []
    for generated in (True, False):
        with sw.Engine(0) as e:
            if generated:
                st = e.generate_phantom(spec)
                m, f = e.get_phantom()
                assert_same(m, f, st, ph)
            else:
                e.set_phantom(ph.mask, ph.fieldmap, fov)
            e.set_sequence(cfg)
            e.set_spins(xyz0)
            e.run_device(mode=sw.MODE_FAST, flags=sw.OUT_ALL)
            outs.append((e.download(), e.sums()))
    (a, sa), (b, sb) = outs
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert np.array_equal(sa[..., 3], sb[..., 3])
    assert np.allclose(sa, sb, rtol=1e-5, atol=1e-4)  # FP32 shared-memory + FP64 global atomics: only the association order of the ensemble sums differs


def test_generator_error_conventions(pg):
    with pytest.raises(pg.PhantomError, match="too large"):
        pg.generate(pg.PhantomSpec(shape=1, fov_um=10.0, resolution=8, radius_um=6.0, seed=1))
    with pytest.raises(pg.PhantomError, match="no usable CUDA device"):
        pg.generate(pg.PhantomSpec(shape=2, fov_um=10.0, resolution=8), device=99)


def test_generator_random_specs_bit_exact(pg, pp, monkeypatch):
    """Randomised `spinwalk phantom` options (seeded): resolutions that are not multiples of the 8x8x32 sphere tile or of the
    4096-voxel broadcast chunk, radii below and above the voxel size, every orientation, with and without field map."""
    import random

    monkeypatch.setenv("SWK_PHANTOM_MAX_REJECTIONS", "200000")
    rnd = random.Random(20241017)
    done = 0
    while done < 16:
        shape = rnd.choice([0, 1])
        fov = float(np.float32(rnd.uniform(20.0, 300.0)))
        res = rnd.choice([7, 9, 16, 17, 23, 31, 33, 40, 47])
        random_radius = rnd.random() < 0.6
        rfrac = rnd.uniform(0.03, 0.2)
        kw = dict(shape=shape, fov_um=fov, resolution=res, radius_um=float(np.float32(fov * rfrac)) * (-1.0 if random_radius else 1.0),
                  volume_fraction=float(np.float32(rnd.uniform(1.0, 12.0) if random_radius else rnd.uniform(8.0, 20.0))), Y=rnd.choice([-1.0, 0.0, 0.6, 0.78, 1.0]),
                  orientation_deg=float(np.float32(rnd.uniform(-10.0, 190.0))), dchi=rnd.choice([0.11e-6, 0.273e-6 * 0.4]), seed=rnd.randint(0, 2 ** 31 - 1))
        try:
            mask, fm, _, st = pg.generate(spec_of(pg, kw))
        except pg.PhantomError as e:
            assert "does not converge" in str(e)
            continue
        assert_same(mask, fm, st, pp.oracle(**kw))
        done += 1


# ---- triangle-mesh phantom (`spinwalk phantom -p`, swk_phantom_mesh) ----

def test_mesh_phantom_bit_exact_vs_oracle_and_reference_golden(pg, pp):
    import meshes

    for name, make in sorted(meshes.MESHES.items()):
        v, f = make()
        gold = np.load(os.path.join(GOLD, "mesh_" + name + ".npz"))
        for fov, res in ((60.0, 24), (100.0, 37), (45.0, 65)):
            mask, fov_m, st = pg.generate_mesh(fov, res, v, f)
            want = pp.oracle_mesh(fov, res, v, f)
            assert np.array_equal(mask, want), (name, fov, res, int(mask.sum()), int(want.sum()))
            assert st["n_shapes"] == len(f) and st["n_launches"] == 1
            assert np.float32(st["volume_fraction"]) == np.float32(want.sum() * 100.0 / want.size)
            if f"sha256_{res}" in gold:
                assert digest(mask) == str(gold[f"sha256_{res}"])
            assert np.array_equal(fov_m, np.full(3, np.float32(fov) * np.float32(1e-6), np.float32))


def test_mesh_phantom_degenerate_inputs(pg, pp):
    import meshes

    v, f = meshes.box()
    # no triangles: every voxel is outside the (empty) bounding box
    mask, _, st = pg.generate_mesh(50.0, 16, v, f[:0])
    assert mask.sum() == 0 and st["n_shapes"] == 0
    # a mesh larger than the FoV: only the part inside is sampled, like the reference (no clipping of the mesh)
    big = v * 0.2  # 200 um box in a 50 um FoV
    mask, _, _ = pg.generate_mesh(50.0, 16, big, f)
    assert np.array_equal(mask, pp.oracle_mesh(50.0, 16, big, f)) and mask.all()
    with pytest.raises(pg.PhantomError, match="out of range"):
        pg.generate_mesh(50.0, 16, v, f + np.uint64(5))


# ---- the reference's own phantom tests (tests/test_phantom.cpp:17-34), on the GPU generator ----

def test_reference_cylinder_creation(pg, pp):
    """BOOST_AUTO_TEST_CASE(cylinder_creation): cylinder(600, 300, 0.11e-6, -1, -20, vf = 10, orientation 0, seed 10).run(false); |actual - vf| < 2."""
    kw = dict(shape=0, fov_um=600.0, resolution=300, dchi=0.11e-6, Y=-1.0, radius_um=-20.0, volume_fraction=10.0, orientation_deg=0.0, seed=10)
    mask, fm, _, st = pg.generate(spec_of(pg, kw))
    assert fm is None and mask.shape == (300, 300, 300)
    assert abs(st["volume_fraction"] - 10.0) < 2.0
    want = pp.oracle(**kw)
    assert np.array_equal(mask, want.mask) and np.float32(st["volume_fraction"]) == np.float32(want.bvf)


def test_reference_sphere_creation(pg, pp):
    """BOOST_AUTO_TEST_CASE(sphere_creation): sphere(600, 300, 0.11e-6, -1, -20, vf = 12, seed 0).run(false); |actual - vf| < 2."""
    kw = dict(shape=1, fov_um=600.0, resolution=300, dchi=0.11e-6, Y=-1.0, radius_um=-20.0, volume_fraction=12.0, seed=0)
    mask, fm, _, st = pg.generate(spec_of(pg, kw))
    assert fm is None
    assert abs(st["volume_fraction"] - 12.0) < 2.0
    want = pp.oracle(**kw)
    assert np.array_equal(mask, want.mask) and np.float32(st["volume_fraction"]) == np.float32(want.bvf)
