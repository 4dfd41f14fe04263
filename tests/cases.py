"""Small named `sim` cases shared by the oracle pinning tests, the golden fixtures and the GPU parity
tests.  Each returns (Case, mask, fieldmap_T, fov_m, xyz0) with seeded, deterministic inputs; sizes are
chosen so that the CPU oracle finishes in seconds."""
from __future__ import annotations

import numpy as np

from oracle import pyoracle as po
from spinwalk_b200.phantoms import cylinder_phantom, sphere_phantom

_cache = {}


def _cyl(n=64, fov_um=64.0, r=6.0, bvf=8.0, seed=1, nz=None):
    key = ("cyl", n, fov_um, r, bvf, seed, nz)
    if key not in _cache:
        _cache[key] = cylinder_phantom(n, fov_um, radius_um=r, bvf_pct=bvf, seed=seed, nz=nz)
    return _cache[key]


def _sph(n=48, fov_um=48.0, r=-8.0, vf=30.0, seed=2, fieldmap=False):
    key = ("sph", n, fov_um, r, vf, seed, fieldmap)
    if key not in _cache:
        _cache[key] = sphere_phantom(n, fov_um, radius_um=r, vf_pct=vf, seed=seed, fieldmap=fieldmap)
    return _cache[key]


def _xyz0(case, fov):
    return po.init_positions(case.seed, fov, case.n_spins, "oracle")


def gre(n_spins=512, scales=(0.05, 0.5, 1.0, 4.0, 20.0)):
    """config/gre.ini on a cylinder phantom: 90 deg pulse, echo at 20 ms, impermeable vessels."""
    mask, fm, fov = _cyl()
    c = po.Case(fov=tuple(fov), phantom_size=mask.shape, n_spins=n_spins, TR_us=40000, TE_tp=[400], scales=list(scales))
    return c, mask, fm, fov, _xyz0(c, fov)


def se(n_spins=512, scales=(0.1, 1.0, 8.0)):
    """config/se.ini: 90 - 180(phase 90) at 10 ms - echo at 20 ms."""
    mask, fm, fov = _cyl()
    c = po.Case(fov=tuple(fov), phantom_size=mask.shape, n_spins=n_spins, TR_us=40000, TE_tp=[400], scales=list(scales),
                RF_FA_deg=[90.0, 180.0], RF_PH_deg=[0.0, 90.0], RF_tp=[0, 200])
    return c, mask, fm, fov, _xyz0(c, fov)


def pgse(n_spins=384, pxy=(1.0, 0.05, 0.9, 1.0), cross_fov=1):
    """PGSE-like: two rectangular gradient lobes around a 180, gradient scaling, permeable spheres,
    no fieldmap, long T1/T2, spins may cross the FoV."""
    mask, _, fov = _sph()
    lobe = list(range(40, 100)) + list(range(140, 200))
    g = [30.0] * 60 + [30.0] * 60
    c = po.Case(fov=tuple(fov), phantom_size=mask.shape, n_spins=n_spins, TR_us=12500, TE_tp=[240], timestep_us=50,
                RF_FA_deg=[90.0, 180.0], RF_PH_deg=[0.0, 90.0], RF_tp=[0, 120],
                gradient_tp=lobe, gradX_mTm=g, gradY_mTm=[0.5 * v for v in g], gradZ_mTm=[0.0] * len(g),
                scales=[0.0, 0.5, 1.0, 2.0], scale_type=po.SCALE_GRADIENT, cross_fov=cross_fov,
                T1_ms=[9999999.0, 9999999.0], T2_ms=[9999999.0, 9999999.0], pXY=list(pxy),
                diffusivity=[1e-9, 2e-9], B0=3.0)
    return c, mask, None, fov, _xyz0(c, fov)


def ssfp(n_spins=256):
    """bSSFP-like: many short TRs (dummy scans from the 5*T1/TR rule), 16 deg pulses, linear phase cycling
    scaled by WHAT_TO_SCALE=2, quadratic term non-zero, short T1/T2 so relaxation matters."""
    mask, fm, fov = _cyl()
    c = po.Case(fov=tuple(fov), phantom_size=mask.shape, n_spins=n_spins, TR_us=2000, TE_tp=[20], timestep_us=50,
                RF_FA_deg=[16.0], RF_PH_deg=[0.0], RF_tp=[0], n_dummy_scan=-1, T1_ms=[20.0, 30.0], T2_ms=[10.0, 8.0],
                linear_phase_cycling=180.0, quadratic_phase_cycling=7.0, scales=[0.0, 0.37, 1.0], scale_type=po.SCALE_PHASE,
                B0=3.0)
    return c, mask, fm, fov, _xyz0(c, fov)


def multi_echo(n_spins=320):
    """three echoes, ideal dephasing events, arbitrary RF phases, three substrates, anisotropic grid and FoV,
    partially permeable, echo and RF on the same timepoint."""
    n = (40, 48, 56)
    rng = np.random.default_rng(5)
    mask = np.zeros(n, np.uint8)
    mask[10:30, 12:36, :] = 1
    mask[14:22, 18:30, 10:40] = 2
    fm = (rng.standard_normal(n) * 2e-8).astype(np.float32)
    fov = np.array([40e-6, 60e-6, 84e-6], np.float32)
    c = po.Case(fov=tuple(fov), phantom_size=n, n_spins=n_spins, TR_us=20000, timestep_us=40,
                TE_tp=[100, 250, 499], RF_FA_deg=[70.0, 35.0, 120.0], RF_PH_deg=[15.0, -90.0, 33.5], RF_tp=[0, 250, 300],
                dephasing_deg=[90.0, 270.0], dephasing_tp=[50, 260],
                diffusivity=[1e-9, 0.5e-9, 2.5e-9], T1_ms=[1500.0, 900.0, -1.0], T2_ms=[60.0, 45.0, 30.0],
                pXY=[1.0, 0.3, 0.0, 0.6, 1.0, 0.25, 1.0, 0.8, 1.0], scales=[0.5, 1.0, 3.0], B0=7.0, seed=77)
    return c, mask, fm, fov, _xyz0(c, fov)


def trajectory(n_spins=96):
    """RECORD_TRAJECTORY=1 with one dummy scan."""
    mask, fm, fov = _cyl()
    c = po.Case(fov=tuple(fov), phantom_size=mask.shape, n_spins=n_spins, TR_us=5000, TE_tp=[50], timestep_us=50,
                n_dummy_scan=1, record_trajectory=1, scales=[0.2, 1.0], seed=3)
    return c, mask, fm, fov, _xyz0(c, fov)


def gradient_rng_free(n_side=9):
    """config/gradient.ini: D = 0 (no RNG at all), one gradient sample => closed-form phase 2 pi per 900 um."""
    n = 16
    mask = np.zeros((n, n, n), np.uint8)
    fov = np.full(3, 900e-6, np.float32)
    g = np.linspace(0.0, 900e-6, n_side, endpoint=False, dtype=np.float32) + np.float32(10e-6)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    xyz0 = np.stack([X.ravel(), Y.ravel(), Z.ravel()], 1).astype(np.float32)
    c = po.Case(fov=tuple(fov), phantom_size=mask.shape, n_spins=xyz0.shape[0], TR_us=40000, TE_tp=[400],
                RF_FA_deg=[90.0], RF_PH_deg=[90.0], RF_tp=[0], gradient_tp=[200], gradX_mTm=[521.9377], gradY_mTm=[0.0],
                gradZ_mTm=[0.0], diffusivity=[0.0], T1_ms=[1000e3], T2_ms=[1000e3], pXY=[1.0], scales=[1.0], seed=10)
    return c, mask, None, fov, xyz0


def stuck(n_spins=256):
    """tiny MAX_ITERATIONS and a fine impermeable checkerboard: many spins are declared lost (kernels.cu:155-159);
    small FoV scale with CROSS_FOV=0 also produces out-of-range exits (kernels.cu:141-147)."""
    n = 32
    i = np.indices((n, n, n)).sum(0)
    mask = (i % 2).astype(np.uint8)
    fov = np.full(3, 48e-6, np.float32)
    c = po.Case(fov=tuple(fov), phantom_size=mask.shape, n_spins=n_spins, TR_us=10000, TE_tp=[50, 150], timestep_us=50,
                max_iterations=3, scales=[0.02, 1.0], seed=21)
    return c, mask, None, fov, _xyz0(c, fov)


def ragged(n_spins=257):
    """ragged launch (one spin more than a block) on a tiny anisotropic 3 x 5 x 7 phantom whose FoV is a few step lengths wide:
    every spin wraps around the periodic FoV many times (CROSS_FOV=1), half-permeable interface, two echoes."""
    mask = np.zeros((3, 5, 7), np.uint8)
    mask[1, 1:4, 2:6] = 1
    rng = np.random.default_rng(11)
    fm = (rng.standard_normal(mask.shape) * 5e-8).astype(np.float32)
    fov = np.array([3e-6, 5e-6, 7e-6], np.float32)
    c = po.Case(fov=tuple(fov), phantom_size=mask.shape, n_spins=n_spins, TR_us=10000, TE_tp=[60, 199], timestep_us=50, cross_fov=1,
                pXY=[1.0, 0.5, 0.5, 1.0], T2_ms=[41.0, 25.0], scales=[0.3, 1.0, 2.5], seed=5)
    return c, mask, fm, fov, _xyz0(c, fov)


def single():
    """one spin: the smallest input the reference accepts (NUMBER_OF_SPINS = 1)."""
    return ragged(n_spins=1)


def events_edge(n_spins=200):
    """event tables at their edges: echoes at timepoint 0, at the last timepoint and beyond the TR (never fires: its slot stays 0);
    RF, dephasing, gradient and echo on one timepoint; gradient samples at timepoint 0, in runs of two and next to the end of the TR;
    RF phases on the exact fast paths of xrot_withphase (180, 270, -90; kernels.cuh:160-195), flip angles > 180 and < 0;
    two dummy scans; gradient scaling with a zero and a negative scale."""
    mask, fm, fov = _cyl()
    c = po.Case(fov=tuple(fov), phantom_size=mask.shape, n_spins=n_spins, TR_us=10000, timestep_us=50, n_dummy_scan=2,
                TE_tp=[0, 57, 199, 230], RF_FA_deg=[45.0, 200.0, -30.0], RF_PH_deg=[180.0, 270.0, -90.0], RF_tp=[0, 57, 100],
                dephasing_deg=[33.0, 720.0], dephasing_tp=[57, 199],
                gradient_tp=[0, 1, 57, 58, 59, 198, 199], gradX_mTm=[5.0, -3.0, 2.0, 4.0, 4.0, 1.0, -6.0],
                gradY_mTm=[0.0, 1.0, 0.0, -2.0, 2.0, 0.0, 0.5], gradZ_mTm=[1.0, 1.0, 1.0, 0.0, 0.0, -1.0, 3.0],
                scales=[0.0, 1.0, -2.0], scale_type=po.SCALE_GRADIENT, T1_ms=[300.0, 500.0], T2_ms=[41.0, 60.0], seed=31, B0=3.0)
    return c, mask, fm, fov, _xyz0(c, fov)


def frozen(n_spins=300):
    """DIFFUSIVITY = 0 in one substrate behind permeable walls: a spin that enters a vessel stops drawing random numbers for good
    (kernels.cu:130 skips the draws when sigma == 0) and keeps accruing the phase of the voxel it froze in."""
    mask, fm, fov = _cyl()
    c = po.Case(fov=tuple(fov), phantom_size=mask.shape, n_spins=n_spins, TR_us=20000, TE_tp=[150, 399], diffusivity=[1e-9, 0.0],
                pXY=[1.0, 1.0, 1.0, 1.0], T2_ms=[41.0, 20.0], scales=[0.2, 1.0], seed=13)
    return c, mask, fm, fov, _xyz0(c, fov)


ALL = dict(gre=gre, se=se, pgse=pgse, ssfp=ssfp, multi_echo=multi_echo, trajectory=trajectory,
           gradient_rng_free=gradient_rng_free, stuck=stuck, ragged=ragged, single=single, events_edge=events_edge,
           frozen=frozen)


def to_simconfig(c: po.Case):
    """oracle Case (timepoints) -> product SimConfig (INI units, microseconds)."""
    from spinwalk_b200 import SimConfig

    us = lambda tp: [int(t) * int(c.timestep_us) for t in tp]  # noqa: E731
    return SimConfig(TR_us=c.TR_us, timestep_us=c.timestep_us, TE_us=us(c.TE_tp), RF_FA_deg=list(c.RF_FA_deg),
                     RF_PH_deg=list(c.RF_PH_deg), RF_T_us=us(c.RF_tp), dephasing_deg=list(c.dephasing_deg),
                     dephasing_T_us=us(c.dephasing_tp), gradient_X_mTm=list(c.gradX_mTm), gradient_Y_mTm=list(c.gradY_mTm),
                     gradient_Z_mTm=list(c.gradZ_mTm), gradient_T_us=us(c.gradient_tp), n_dummy_scan=c.n_dummy_scan,
                     linear_phase_cycling=c.linear_phase_cycling, quadratic_phase_cycling=c.quadratic_phase_cycling,
                     B0=c.B0, seed=c.seed, n_spins=c.n_spins, cross_fov=c.cross_fov, record_trajectory=c.record_trajectory,
                     max_iterations=c.max_iterations, scales=list(c.scales), scale_type=c.scale_type,
                     diffusivity=list(c.diffusivity), T1_ms=list(c.T1_ms), T2_ms=list(c.T2_ms), pXY=list(c.pXY))
