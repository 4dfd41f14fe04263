"""Seeded random `sim` cases over the whole parameter space of the hot path: phantom shape and substrates, field map or
none, anisotropic FoV, every scale type, boundary rule, dummy scans, event tables (RF / echo / dephasing / gradient, with coinciding
timepoints, runs of consecutive gradient samples, entries at 0 and beyond the TR), zero diffusivity, P_XY in {0, 1, in between}, negative
T1 / T2, small MAX_ITERATIONS, trajectory recording.  Shared by the oracle pinning (CPU) and, later, engine parity tests."""
import numpy as np

from oracle import pyoracle as po


def _times(rng, n_tp, k, lo=0, allow_beyond=True):
    """k strictly ascending timepoints; sometimes one at `lo`, sometimes consecutive ones, sometimes one beyond the TR."""
    if k == 0:
        return []
    hi = n_tp + (6 if allow_beyond and rng.random() < 0.3 else 0)
    t = set(int(x) for x in rng.integers(lo, max(lo + 1, hi), k))
    if rng.random() < 0.3:
        t.add(lo)
    if rng.random() < 0.5 and k >= 2:  # a run of consecutive timepoints
        s = int(rng.integers(lo, max(lo + 1, n_tp - 4)))
        t.update(range(s, s + int(rng.integers(2, 5))))
    return sorted(t)


def make(seed):
    rng = np.random.default_rng([seed, 20261017])
    dims = tuple(int(x) for x in rng.integers(3, 20, 3))
    ns = int(rng.integers(1, 5))
    mask = np.zeros(dims, np.uint8)
    for s in range(1, ns):  # a few boxes per substrate; every substrate is present
        for _ in range(int(rng.integers(1, 4))):
            a = [int(rng.integers(0, d)) for d in dims]
            b = [min(d, x + int(rng.integers(1, max(2, d // 2 + 1)))) for x, d in zip(a, dims)]
            mask[a[0]:b[0], a[1]:b[1], a[2]:b[2]] = s
        mask[tuple(int(rng.integers(0, d)) for d in dims)] = s
    if ns > 1 and rng.random() < 0.3:
        mask = (np.indices(dims).sum(0) % ns).astype(np.uint8)  # fine checkerboard: many crossings
    fm = None if rng.random() < 0.3 else (rng.standard_normal(dims) * 10.0 ** rng.uniform(-9, -7)).astype(np.float32)
    fov = (np.array(dims) * rng.uniform(0.4e-6, 3e-6, 3)).astype(np.float32)
    dt = int(rng.choice([10, 25, 50, 100]))
    n_tp = int(rng.integers(20, 160))
    n_rf = int(rng.integers(1, 5))
    special = [0.0, 90.0, 180.0, 270.0, -90.0]
    rf_tp = [0] + [t for t in _times(rng, n_tp, n_rf - 1, lo=1, allow_beyond=False) if t > 0][: n_rf - 1]
    n_rf = len(rf_tp)
    te_tp = _times(rng, n_tp, int(rng.integers(1, 5)))
    de_tp = _times(rng, n_tp, int(rng.integers(0, 4)))
    gr_tp = _times(rng, n_tp, int(rng.integers(0, 8)))
    scale_type = int(rng.integers(0, 3))
    if scale_type == po.SCALE_FOV:
        scales = [float(x) for x in 10.0 ** rng.uniform(-1.3, 1.2, int(rng.integers(1, 4)))]
    else:
        scales = [float(x) for x in rng.choice([0.0, 1.0, -1.5, 0.37, 2.0, 5.0], int(rng.integers(1, 4)))]
    pxy = rng.choice([0.0, 1.0, 0.05, 0.5, 0.9], (ns, ns)).astype(float)
    case = po.Case(
        fov=tuple(float(x) for x in fov), phantom_size=dims, n_spins=int(rng.integers(1, 80)), TR_us=n_tp * dt + int(rng.integers(0, dt)), timestep_us=dt,
        seed=int(rng.integers(1, 5000)),  # small: the CPU reference discard()s seed + spin values one by one
         max_iterations=int(rng.choice([2, 50, 10000])), B0=float(rng.choice([1.5, 3.0, 7.0, 9.4])),
        linear_phase_cycling=float(rng.choice([0.0, 180.0, 33.0])), quadratic_phase_cycling=float(rng.choice([0.0, 0.0, 117.0])),
        n_dummy_scan=int(rng.integers(0, 4)), cross_fov=int(rng.integers(0, 2)), record_trajectory=int(rng.random() < 0.2),
        diffusivity=[float(x) for x in rng.choice([0.0, 0.3e-9, 1e-9, 3e-9], ns, p=[0.1, 0.3, 0.4, 0.2])],
        T1_ms=[float(x) for x in rng.choice([-1.0, 30.0, 900.0, 2200.0], ns)], T2_ms=[float(x) for x in rng.choice([-1.0, 8.0, 41.0, 75.0], ns)],
        pXY=[float(x) for x in pxy.ravel()],
        RF_FA_deg=[float(x) for x in rng.choice([16.0, 45.0, 90.0, 180.0, 200.0, -30.0], n_rf)],
        RF_PH_deg=[float(rng.choice(special)) if rng.random() < 0.6 else float(rng.uniform(-200, 400)) for _ in range(n_rf)], RF_tp=rf_tp,
        TE_tp=te_tp, dephasing_deg=[float(x) for x in rng.uniform(-400, 400, len(de_tp))], dephasing_tp=de_tp,
        gradX_mTm=[float(x) for x in rng.uniform(-40, 40, len(gr_tp))], gradY_mTm=[float(x) for x in rng.uniform(-40, 40, len(gr_tp))],
        gradZ_mTm=[float(x) for x in rng.uniform(-40, 40, len(gr_tp))], gradient_tp=gr_tp, scales=scales, scale_type=scale_type)
    xyz0 = po.init_positions(case.seed, fov, case.n_spins, "oracle")
    return case, mask, fm, fov, xyz0
