"""GPU parity tests: the CUDA engine, called through the C-ABI (ctypes -> libspinwalk_b200.so), against
  (1) the reference's own cu_sim kernel compiled for sm_100a (oracle/_ref/libswref_cuda.so) on the same GPU,
  (2) the CPU oracle (oracle/sim_oracle.c),
on the same seeded inputs.  Bars:
  T (tissue index at echo)        bit-exact   (COMPAT mode vs reference cu_sim)
  XYZ1 (positions / trajectories) bit-exact   (COMPAT mode vs reference cu_sim)
  M1 (magnetisation)              |diff| <= 2e-6 absolute (FP32 round-off of the event math; the walk itself is exact)
  FAST mode (Philox)              ensemble means within 4.5 combined standard errors of the oracle's
"""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu

M1_TOL = 2e-6


@pytest.fixture(scope="module")
def sw(engine_lib):
    import spinwalk_b200 as sw

    assert engine_lib.swk_device_count() > 0, "no CUDA device visible: GPU tests must not fall back to anything"
    return sw


def _run_engine(sw, case, mask, fm, fov, xyz0, mode, **kw):
    cfg = cases.to_simconfig(case)
    with sw.Engine(0) as e:
        e.set_phantom(mask, fm, fov)
        e.set_sequence(cfg)
        return e.run(xyz0, mode=mode, **kw)


@pytest.mark.parametrize("name", list(cases.ALL))
def test_compat_bit_exact_vs_reference_cu_sim(sw, oracle, name):
    """Engine (COMPAT arithmetic) == the reference's cu_sim launched on the same device."""
    if not oracle.have_ref_cuda():
        pytest.skip("oracle/_ref/libswref_cuda.so not present")
    case, mask, fm, fov, xyz0 = cases.ALL[name]()
    ref = oracle.run_ref_cuda(case, fm, mask, xyz0)
    got = _run_engine(sw, case, mask, fm, fov, xyz0, sw.MODE_COMPAT)
    assert np.array_equal(got["T"], ref["T"]), "tissue index at echo differs"
    assert np.array_equal(got["XYZ1"].view(np.uint32), ref["XYZ1"].view(np.uint32)), "positions differ bitwise"
    d = np.abs(got["M1"] - ref["M1"]).max()
    assert d <= M1_TOL, f"max |dM1| = {d}"


@pytest.mark.parametrize("name", list(cases.ALL))
def test_compat_vs_cpu_oracle(sw, oracle, name):
    """Engine (COMPAT) vs the CPU oracle, minstd flavour.  Host erfcinvf (double based) and libdevice erfcinvf
    differ by ulps, so positions drift apart by ~1e-12 m and once in ~1e6 steps a spin sees the neighbouring
    voxel for a step (another field sample, or another permeability outcome).  Bar: >= 99% of (scale, spin)
    pairs have the same tissue at every echo, positions within 1e-9 m and M1 within 2e-4."""
    case, mask, fm, fov, xyz0 = cases.ALL[name]()
    ora = oracle.run_oracle(case, fm, mask, xyz0, flavour=oracle.RNG_MINSTD)
    got = _run_engine(sw, case, mask, fm, fov, xyz0, sw.MODE_COMPAT)
    same = ((got["T"] == ora["T"]).all(axis=2) & (np.abs(got["XYZ1"] - ora["XYZ1"]).max(axis=(2, 3)) < 1e-9)
            & (np.abs(got["M1"] - ora["M1"]).max(axis=(2, 3)) <= 2e-4))
    assert same.mean() >= 0.99, f"only {same.mean():.4f} of (scale, spin) pairs follow the oracle"
    # work counters agree with the oracle's to the same extent
    st, so = got["stats"], ora["stats"]
    assert st["steps"] == pytest.approx(so["steps"], rel=2e-3)
    assert st["mask_gathers"] == pytest.approx(so["mask_gathers"], rel=5e-3, abs=5)
    assert st["rejects"] == pytest.approx(so["rejects"], rel=2e-2, abs=5)
    assert st["lost"] == pytest.approx(so["lost"], abs=max(2, 0.01 * so["lost"]))


def _ensemble(M1, T, sub=None):
    """mean and standard error of Mx, My, Mz per (scale, echo) [optionally restricted to a substrate]."""
    w = np.ones(T.shape, bool) if sub is None else (T == sub)
    n = np.maximum(w.sum(axis=1), 1)[..., None]  # [K,E,1]
    m = (M1 * w[..., None]).sum(axis=1) / n
    v = ((M1 - m[:, None]) ** 2 * w[..., None]).sum(axis=1) / n
    return m, np.sqrt(v / n), w.mean(axis=1)


@pytest.mark.parametrize("name,n_spins", [("gre", 6000), ("se", 6000), ("pgse", 6000), ("ssfp", 2000), ("multi_echo", 6000), ("events_edge", 5000),
                                          ("frozen", 5000)])
def test_fast_mode_ensemble_vs_oracle(sw, oracle, name, n_spins):
    """FAST mode (Philox + Box-Muller + FP32 grid coordinates) is a different random stream, so parity is
    statistical: per (scale, echo, component) |mean_fast - mean_oracle| <= 4.5 sqrt(SE_fast^2 + SE_oracle^2),
    plus an absolute floor of 2e-3 for nearly deterministic components, and the same tissue occupancy."""
    case, mask, fm, fov, xyz0 = cases.ALL[name](n_spins=n_spins)
    ora = oracle.run_oracle(case, fm, mask, xyz0, flavour=oracle.RNG_MT19937)
    got = _run_engine(sw, case, mask, fm, fov, xyz0, sw.MODE_FAST)
    assert got["stats"]["lost"] == 0
    mo, so, oo = _ensemble(ora["M1"], ora["T"])
    mg, sg, og = _ensemble(got["M1"], got["T"])
    tol = 4.5 * np.sqrt(so**2 + sg**2) + 2e-3
    assert (np.abs(mo - mg) <= tol).all(), f"ensemble mismatch: max excess {(np.abs(mo - mg) - tol).max()}"
    for sub in range(case.n_substrate):
        _, _, fo = _ensemble(ora["M1"], ora["T"], sub)
        _, _, fg = _ensemble(got["M1"], got["T"], sub)
        assert np.abs(fo - fg).max() <= 4.5 * np.sqrt(0.25 / n_spins * 2) + 1e-3
    # step-size statistics: rms displacement per axis matches (free-ish diffusion at the largest scale)
    k = int(np.argmax(case.scales)) if case.scale_type == oracle.SCALE_FOV else 0
    s = case.scales[k] if case.scale_type == oracle.SCALE_FOV else 1.0
    do = ora["XYZ1"][k, :, -1, :] - xyz0 * np.float32(s)
    dg = got["XYZ1"][k, :, -1, :] - xyz0 * np.float32(s)
    if not case.cross_fov:
        assert np.allclose(do.std(axis=0), dg.std(axis=0), rtol=0.08)


def test_fast_rng_free_gradient_closed_form(sw, oracle):
    """config/gradient.ini: D = 0, one 50 us gradient sample of 521.9377 mT/m at 10 ms => the transverse phase
    at the echo is gamma*G*x*dt (2 pi per 900 um).  No RNG is involved, so FAST and COMPAT must agree with the
    oracle to FP32 round-off and with the closed form to 1e-3 rad."""
    case, mask, fm, fov, xyz0 = cases.gradient_rng_free()
    ora = oracle.run_oracle(case, fm, mask, xyz0, flavour=oracle.RNG_MINSTD)
    for mode in (sw.MODE_COMPAT, sw.MODE_FAST):
        got = _run_engine(sw, case, mask, fm, fov, xyz0, mode)
        assert np.abs(got["M1"] - ora["M1"]).max() <= (2e-6 if mode == sw.MODE_COMPAT else 2e-4)
        ph = np.arctan2(got["M1"][0, :, 0, 1], got["M1"][0, :, 0, 0])
        # RF 90 about y puts M along +x; phase advances by gamma*G*x*dt
        expect = 267515315.0 * 521.9377e-3 * xyz0[:, 0].astype(np.float64) * 50e-6
        dphi = np.angle(np.exp(1j * (ph - expect)))
        assert np.abs(dphi).max() < 1e-3


@pytest.mark.parametrize("mode_name", ["COMPAT", "FAST"])
def test_sums_match_per_spin_outputs(sw, mode_name):
    """the in-kernel ensemble sums [scale][echo][substrate][Mx,My,Mz,N] equal a host reduction of M1 / T."""
    mode = getattr(sw, "MODE_" + mode_name)
    case, mask, fm, fov, xyz0 = cases.multi_echo(n_spins=3000)
    got = _run_engine(sw, case, mask, fm, fov, xyz0, mode)
    K, E, ns = case.n_scales, case.n_TE, case.n_substrate
    valid = (got["M1"] != 0).any(axis=3)  # lost / unwritten echoes stay zero
    for sub in range(ns):
        w = (got["T"] == sub) & valid
        assert np.array_equal(got["sums"][:, :, sub, 3], w.sum(axis=1).astype(np.float64))
        ref = (got["M1"].astype(np.float64) * w[..., None]).sum(axis=1)
        assert np.allclose(got["sums"][:, :, sub, :3], ref, rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("mode_name", ["COMPAT", "FAST"])
def test_shards_are_invariant(sw, mode_name):
    """RNG and the dephasing term are keyed by the GLOBAL spin id: simulating [0,n/2) and [n/2,n) on separate
    engines gives exactly the arrays of the single run (what makes multi-GPU results independent of G)."""
    mode = getattr(sw, "MODE_" + mode_name)
    case, mask, fm, fov, xyz0 = cases.multi_echo(n_spins=1000)
    cfg = cases.to_simconfig(case)
    full = _run_engine(sw, case, mask, fm, fov, xyz0, mode)
    parts = []
    for first, n in ((0, 437), (437, 563)):
        with sw.Engine(0) as e:
            e.set_phantom(mask, fm, fov)
            e.set_sequence(cfg)
            parts.append(e.run(xyz0[first:first + n], spin_first=first, mode=mode))
    for key in ("M1", "XYZ1", "T"):
        cat = np.concatenate([p[key] for p in parts], axis=1)
        assert np.array_equal(cat, full[key]), key
    assert np.array_equal(parts[0]["sums"] + parts[1]["sums"], full["sums"])  # fixed-point sums: exact, whatever the split


def test_device_resident_run_and_device_positions(sw):
    """set_spins(NULL) draws positions on the device inside [1%,99%] of the FoV; run_device + download equals run()."""
    case, mask, fm, fov, xyz0 = cases.gre(n_spins=2048, scales=(1.0,))
    cfg = cases.to_simconfig(case)
    with sw.Engine(0) as e:
        e.set_phantom(mask, fm, fov)
        e.set_sequence(cfg)
        e.set_spins(xyz0)
        st = e.run_device(mode=sw.MODE_FAST, flags=sw.OUT_ALL | sw.RUN_STATS)
        m1, x1, t = e.download()
        host = e.run(xyz0, mode=sw.MODE_FAST)
        assert np.array_equal(m1, host["M1"]) and np.array_equal(t, host["T"]) and np.array_equal(x1, host["XYZ1"])
        assert st["steps"] == case.total_steps() and st["n_launches"] >= 1 and st["kernel_ms"] > 0
        e.set_spins(None, n_local=2048)
        e.run_device(mode=sw.MODE_FAST, flags=sw.OUT_XYZ1)
        _, x1, _ = e.download(M1=False, T=False)
        assert np.isfinite(x1).all() and (x1 >= 0).all() and (x1 < np.asarray(fov)).all()


def test_pipelined_host_run_equals_single_launch(sw, monkeypatch):
    """swk_run cuts large runs into slices (two compute streams, download of slice i overlapping the walk of slice i+1)
    and sorts spins slice-major; results are keyed by spin id, so they must equal the single-launch
    run_device + download bit for bit, and the ensemble sums must agree.  SWK_SLICES forces 3 slices at a test-sized run."""
    monkeypatch.setenv("SWK_SLICES", "3")
    case, mask, fm, fov, _ = cases.gre(n_spins=3 * 2**18 + 777, scales=(0.3, 2.0))
    case.TR_us, case.TE_tp = 2500, [20, 45]
    rng = np.random.default_rng(4)
    xyz0 = (rng.random((case.n_spins, 3), dtype=np.float32) * np.float32(0.98) + np.float32(0.01)) * np.asarray(fov, np.float32)
    cfg = cases.to_simconfig(case)
    for mode in (sw.MODE_FAST, sw.MODE_COMPAT):
        with sw.Engine(0) as e:
            e.set_phantom(mask, fm, fov)
            e.set_sequence(cfg)
            piped = e.run(xyz0, mode=mode, stats=False)
            assert piped["stats"]["n_launches"] >= 3  # three slices
            e.set_spins(xyz0)
            e.run_device(mode=mode, flags=sw.OUT_ALL)
            m1, x1, t = e.download()
            sums = e.sums()
        assert np.array_equal(piped["M1"], m1) and np.array_equal(piped["XYZ1"], x1) and np.array_equal(piped["T"], t)
        assert np.allclose(piped["sums"], sums, rtol=1e-6, atol=1e-3)
        assert piped["sums"][..., 3].sum() == case.n_spins * case.n_scales * case.n_TE


def test_error_conventions(sw):
    """bad inputs fail with a status + message (≙ the reference's `return false` + log line), never silently."""
    case, mask, fm, fov, xyz0 = cases.gre(n_spins=64, scales=(1.0,))
    cfg = cases.to_simconfig(case)
    with sw.Engine(0) as e:
        with pytest.raises(sw.EngineError, match="no phantom"):
            e.set_sequence(cfg)
            e.set_spins(xyz0)
            e.run_device()
        e.set_phantom(mask, fm, fov)
        bad = cases.to_simconfig(case)
        bad.RF_T_us = [50]
        with pytest.raises(sw.EngineError, match="first RF start time"):
            e.set_sequence(bad)
        bad = cases.to_simconfig(case)
        bad.TE_us = [20000, 10000]
        with pytest.raises(sw.EngineError, match="ascending"):
            e.set_sequence(bad)
        one = cases.to_simconfig(case)
        one.diffusivity, one.T1_ms, one.T2_ms, one.pXY = [1e-9], [1000.0], [50.0], [1.0]
        e.set_sequence(one)
        e.set_spins(xyz0)
        with pytest.raises(sw.EngineError, match="substrate"):  # mask has 2 substrates, config 1 (monte_carlo.cu:113-118)
            e.run_device()
    with pytest.raises(sw.EngineError, match="not available"):
        sw.Engine(999)


def test_fast_pgse_free_diffusion_known_answer(sw):
    """Known answer of the reference's own demo (demo/spinwalk_dwi.ipynb cells 7-20): PGSE (dwi -b ... -v 1 0 0 -d 15 10 20) on a
    phantom whose walls are fully permeable (all P_XY = 1), no relaxation => S(b) = exp(-b D).  The notebook's recorded fit of the
    reference's output is D = 9.79e-10 (simulated 1e-9), a = 0.9945; a 4 % window is required of the FAST path at 2^18 spins
    (Monte-Carlo error of |S| ~ 2e-3)."""
    from spinwalk_b200.phantoms import sphere_lattice_phantom
    from spinwalk_b200.sequences import pgse

    b = [100.0, 500.0, 1000.0, 2000.0, 3000.0, 0.0]
    # FoV 1 mm as in the notebook's scale (600 um there): with CROSS_FOV = 1 a spin that wraps around the FoV jumps by one FoV in the
    # gradient's frame and is lost to the signal; at 11 um rms displacement that is ~1 % of the spins here (a = 0.9945 there)
    mask, _, fov = sphere_lattice_phantom(128, 1000.0, 250.0, 35.0, seed=3)
    S = 1 << 18
    cfg = sw.SimConfig(TR_us=60050, TE_us=[60000], timestep_us=50, seed=21, n_spins=S, cross_fov=1, B0=9.4,
                       diffusivity=[1e-9, 1e-9], T1_ms=[9999999.0] * 2, T2_ms=[9999999.0] * 2, pXY=[1.0] * 4, **pgse(b))
    with sw.Engine(0) as e:
        e.set_phantom(mask, None, fov)
        e.set_sequence(cfg)
        e.set_spins(None, n_local=S)
        e.run_device(mode=sw.MODE_FAST, flags=0)
        sums = e.sums()  # [K][E][sub][Mx,My,Mz,N]
    tot = sums.sum(axis=2)[:, 0, :]
    assert np.all(tot[:, 3] == S)
    sig = np.hypot(tot[:, 0], tot[:, 1]) / S
    assert abs(sig[-1] - 1.0) < 1e-3  # b = 0
    bb = np.asarray(b[:-1]) * 1e6  # s/mm^2 -> s/m^2
    slope, intercept = np.polyfit(bb, np.log(sig[:-1]), 1)
    assert abs(-slope / 1e-9 - 1.0) < 0.04, f"fitted D = {-slope:.3e}, a = {np.exp(intercept):.4f}, S = {sig}"
    assert abs(np.exp(intercept) - 1.0) < 0.015, f"fitted D = {-slope:.3e}, a = {np.exp(intercept):.4f}, S = {sig}"


@pytest.mark.parametrize("scale_type_name", ["FOV", "PHASE_CYCLING"])
def test_rebinned_long_run_equals_uninterrupted_run(sw, oracle, monkeypatch, scale_type_name):
    """Long runs (many TRs) are paused at TR boundaries to re-sort the spins by their current voxel (engine.cu run_impl).  The
    pause stores position, magnetisation, substrate and the RNG block counter and the walk resumes from exactly that state, so
    the outputs must equal the uninterrupted run bit for bit — with one order per scale (FoV scaling) and with a shared order
    (phase-cycling scaling, one walker per (spin, scale): SWK_RUN_NO_ONEWALK — the default for such scales, one walk for all of them, is
    never paused).  SWK_REBIN_SCANS=3 forces a pause every 3 TRs on a test-sized run."""
    case, mask, fm, fov, xyz0 = cases.ssfp(n_spins=1500)
    if scale_type_name == "FOV":
        case.scale_type, case.scales = oracle.SCALE_FOV, [0.7, 1.0, 1.9]
    cfg = cases.to_simconfig(case)
    per_scale = sw.RUN_NO_ONEWALK
    with sw.Engine(0) as e:
        e.set_phantom(mask, fm, fov)
        e.set_sequence(cfg)
        e.set_spins(xyz0)
        assert e.n_dummy_scan >= 20
        st0 = e.run_device(mode=sw.MODE_FAST, flags=sw.OUT_ALL | sw.RUN_NO_REBIN | sw.RUN_STATS | per_scale)
        ref = e.download() + (e.sums(),)
        monkeypatch.setenv("SWK_REBIN_SCANS", "3")
        st1 = e.run_device(mode=sw.MODE_FAST, flags=sw.OUT_ALL | sw.RUN_STATS | per_scale)
        got = e.download() + (e.sums(),)
        if scale_type_name == "PHASE_CYCLING":  # one walk for all scales: one launch whatever the re-binning knob says, same walk
            st3 = e.run_device(mode=sw.MODE_FAST, flags=sw.OUT_ALL | sw.RUN_STATS)
            one = e.download()
            assert st3["n_launches"] <= 3 and np.array_equal(one[1], ref[1]) and np.array_equal(one[2], ref[2])
            np.testing.assert_allclose(one[0], ref[0], rtol=0, atol=3e-5)
        monkeypatch.delenv("SWK_REBIN_SCANS")
        st2 = e.run_device(mode=sw.MODE_FAST, flags=sw.OUT_ALL | per_scale)  # back to one launch: the start order is still valid
        again = e.download()
    assert st1["n_launches"] > st0["n_launches"] + 5
    for a, b, c in zip(ref[:3], got[:3], again):
        assert np.array_equal(a, b) and np.array_equal(a, c)
    assert np.allclose(ref[3], got[3], rtol=1e-6, atol=1e-3)
    for key in ("steps", "mask_gathers", "field_gathers", "rejects", "lost"):
        assert st0[key] == st1[key], key
    assert st2["n_launches"] <= 3  # one walk launch (+ the unpack pass and the conversion of the sums)


@pytest.mark.parametrize("name", ["se", "multi_echo", "events_edge"])
def test_fast_kernel_variants_agree(sw, name):
    """The FAST walk has three voxel fetches (packed word / mask byte + FP32 field / mask only) and runs with or without the
    locality order.  A lane's random stream depends on its own history only, so: unsorted == sorted bit for bit; the split
    fetch (RUN_NO_PACK: the exact FP32 field instead of the packed word's 20 mantissa bits) walks the very same path (T and XYZ1
    bitwise) and its magnetisation differs by the field rounding only (relative 2^-21 of the accrued phase)."""
    case, mask, fm, fov, xyz0 = cases.ALL[name](n_spins=700)
    cfg = cases.to_simconfig(case)
    with sw.Engine(0) as e:
        e.set_phantom(mask, fm, fov)
        e.set_sequence(cfg)
        e.set_spins(xyz0)
        e.run_device(mode=sw.MODE_FAST, flags=sw.OUT_ALL)
        base = e.download() + (e.sums(),)
        e.run_device(mode=sw.MODE_FAST, flags=sw.OUT_ALL | sw.RUN_NO_SORT)
        unsorted = e.download() + (e.sums(),)
        e.run_device(mode=sw.MODE_FAST, flags=sw.OUT_ALL | sw.RUN_NO_PACK)
        split = e.download() + (e.sums(),)
        e.run_device(mode=sw.MODE_FAST, flags=sw.OUT_ALL)  # and back: the order and the packed table are rebuilt
        again = e.download()
    for a, b, c in zip(base[:3], unsorted[:3], again):
        assert np.array_equal(a, b) and np.array_equal(a, c)
    assert np.allclose(base[3], unsorted[3], rtol=1e-9, atol=1e-3)  # sums: another summation order of the same FP32 values
    assert np.array_equal(base[1], split[1]) and np.array_equal(base[2], split[2]), "the split fetch must walk the same path"
    assert np.abs(base[0] - split[0]).max() <= 2e-4


def test_single_spin_and_ragged_sizes(sw, oracle):
    """sizes around the launch granularity (1, 255, 256, 257 spins; 256-thread blocks): the first spins of a larger population
    give the same results whatever the population size, in both modes (the dephasing-free case does not depend on n_spins)."""
    case, mask, fm, fov, xyz0 = cases.ragged(n_spins=257)
    for mode in (sw.MODE_COMPAT, sw.MODE_FAST):
        full = None
        for n in (257, 256, 255, 1):
            case.n_spins = n
            got = _run_engine(sw, case, mask, fm, fov, xyz0[:n], mode)
            assert got["M1"].shape == (3, n, 2, 3) and got["T"].shape == (3, n, 2)
            if full is None:
                full = got
            else:
                for k in ("M1", "XYZ1", "T"):
                    assert np.array_equal(got[k], full[k][:, :n]), (mode, n, k)


def test_zslab_walks_the_same_path(sw):
    """A phantom whose mask and field map do not depend on z — every cylinder phantom — is walked on the packed words of ONE z plane (the
    default; SWK_RUN_NO_ZSLAB keeps the full [nx][ny][nz] table).  They are the words of the full table, so the results are the full-table
    results bit for bit; a phantom that does depend on z takes the full table either way."""
    case, mask, fm, fov, xyz0 = cases.se(n_spins=900)
    assert (mask == mask[:, :, :1]).all() and (fm.view(np.uint32) == fm[:, :, :1].view(np.uint32)).all(), "the test phantom must be z-invariant"
    cfg = cases.to_simconfig(case)
    with sw.Engine(0) as e:
        e.set_phantom(mask, fm, fov)
        e.set_sequence(cfg)
        e.set_spins(xyz0)
        st0 = e.run_device(mode=sw.MODE_FAST, flags=sw.OUT_ALL | sw.RUN_STATS | sw.RUN_NO_ZSLAB)
        base = e.download() + (e.sums(),)
        st1 = e.run_device(mode=sw.MODE_FAST, flags=sw.OUT_ALL | sw.RUN_STATS)
        slab = e.download() + (e.sums(),)
        st2 = e.run_device(mode=sw.MODE_FAST, flags=sw.OUT_ALL)
    assert st1["n_launches"] == st2["n_launches"] + 2  # z-invariance check + slab packing happen once per phantom
    for a, b in zip(base, slab):
        assert np.array_equal(a, b)  # (the sums too: they are accumulated in integer fixed point)
    for key in ("steps", "mask_gathers", "field_gathers", "rejects", "lost"):
        assert st0[key] == st1[key], key
    # COMPAT mode walks the raw (substrate id, FP32 field) pairs of the same plane: the values of the full arrays, bit-identical results
    for name in ("se", "ssfp", "events_edge"):
        case, mask, fm, fov, xyz0 = cases.ALL[name]()
        cfg = cases.to_simconfig(case)
        with sw.Engine(0) as e:
            e.set_phantom(mask, fm, fov)
            e.set_sequence(cfg)
            e.set_spins(xyz0)
            st0 = e.run_device(mode=sw.MODE_COMPAT, flags=sw.OUT_ALL | sw.RUN_STATS | sw.RUN_NO_ZSLAB)
            base = e.download() + (e.sums(),)
            st1 = e.run_device(mode=sw.MODE_COMPAT, flags=sw.OUT_ALL | sw.RUN_STATS)
            slab = e.download() + (e.sums(),)
        for a, b in zip(base, slab):
            assert np.array_equal(a, b), name
        for key in ("steps", "mask_gathers", "field_gathers", "rejects", "lost"):
            assert st0[key] == st1[key], (name, key)
    # not invariant along z: the full table, whatever the flag
    case, mask, fm, fov, xyz0 = cases.multi_echo(n_spins=400)
    cfg = cases.to_simconfig(case)
    with sw.Engine(0) as e:
        e.set_phantom(mask, fm, fov)
        e.set_sequence(cfg)
        e.set_spins(xyz0)
        st_a = e.run_device(mode=sw.MODE_FAST, flags=sw.OUT_ALL)
        base = e.download()
        st_b = e.run_device(mode=sw.MODE_FAST, flags=sw.OUT_ALL | sw.RUN_NO_ZSLAB)
        again = e.download()
    assert st_a["n_launches"] >= st_b["n_launches"]  # the check itself ran in the first run only
    for a, b in zip(base, again):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("name", ["multi_echo", "ragged", "se"])
def test_brick_layout_walks_the_same_path(sw, monkeypatch, name):
    """SWK_BRICK=1 stores the packed voxel words in bricks of 2 x 2 x 4 voxels (one 64-byte fetch unit each) instead of row-major: the same
    words at other addresses, so the same results bit for bit — even, odd (3 x 5 x 7) and z-invariant (full table forced) phantoms."""
    case, mask, fm, fov, xyz0 = cases.ALL[name]()
    cfg = cases.to_simconfig(case)
    with sw.Engine(0) as e:
        e.set_phantom(mask, fm, fov)
        e.set_sequence(cfg)
        e.set_spins(xyz0)
        st0 = e.run_device(mode=sw.MODE_FAST, flags=sw.OUT_ALL | sw.RUN_STATS | sw.RUN_NO_ZSLAB)
        base = e.download() + (e.sums(),)
        monkeypatch.setenv("SWK_BRICK", "1")
        st1 = e.run_device(mode=sw.MODE_FAST, flags=sw.OUT_ALL | sw.RUN_STATS | sw.RUN_NO_ZSLAB)
        brick = e.download() + (e.sums(),)
        monkeypatch.delenv("SWK_BRICK")
        e.run_device(mode=sw.MODE_FAST, flags=sw.OUT_ALL | sw.RUN_NO_ZSLAB)  # and back: the table is rebuilt row-major
        again = e.download()
    for a, b, c in zip(base[:3], brick[:3], again):
        assert np.array_equal(a, b) and np.array_equal(a, c)
    assert np.array_equal(base[3], brick[3])
    for key in ("steps", "mask_gathers", "field_gathers", "rejects", "lost"):
        assert st0[key] == st1[key], key


def test_shared_and_private_streams_agree(sw):
    """The SHARED kernel variant (a block walks 32 spins x G scales and generates each spin's normals once) and the PRIVATE one (every thread
    generates its own) draw the same numbers for the same (spin, round): identical outputs, sums and counters — FoV, gradient and phase scaling."""
    for name in ("gre", "pgse", "ssfp", "multi_echo", "ragged"):
        case, mask, fm, fov, xyz0 = cases.ALL[name]()
        cfg = cases.to_simconfig(case)
        with sw.Engine(0) as e:
            e.set_phantom(mask, fm, fov)
            e.set_sequence(cfg)
            e.set_spins(xyz0)
            st0 = e.run_device(mode=sw.MODE_FAST, flags=sw.OUT_ALL | sw.RUN_STATS | sw.RUN_NO_ONEWALK)
            shared = e.download() + (e.sums(),)
            st1 = e.run_device(mode=sw.MODE_FAST, flags=sw.OUT_ALL | sw.RUN_STATS | sw.RUN_NO_SHARE | sw.RUN_NO_ONEWALK)
            private = e.download() + (e.sums(),)
        for a, b in zip(shared, private):
            assert np.array_equal(a, b), name
        for key in ("steps", "mask_gathers", "field_gathers", "rejects", "lost"):
            assert st0[key] == st1[key], (name, key)


@pytest.mark.parametrize("name", ["pgse", "ssfp", "events_edge", "stuck_gradient", "long_lobes"])
def test_one_walk_for_all_scales_equals_one_walk_per_scale(sw, name, monkeypatch):
    """WHAT_TO_SCALE = 1 (gradients) or 2 (phase cycling): every scale of a spin walks the same path — the reference re-seeds seed+spin per
    scale (kernels.cu:77-88) and the FoV is not scaled — so by default ONE walker per spin carries the magnetisation of every scale
    (walk_fast.cuh MULTI).  Against one walker per (spin, scale) (SWK_RUN_NO_ONEWALK): final positions, tissues at echo and the lost count are
    bit-identical, the counters are the per-scale counters times the number of scales, and the magnetisations agree to FP32 round-off (the
    gradient phase is scaled after the sum instead of term by term).  Gradient runs (PGSE lobes), single gradient samples, dummy scans with phase
    cycling, several echoes, abandoned spins; one launch and the sliced host run."""
    if name == "stuck_gradient":  # abandoned spins (kernels.cu:155-159) under gradient scaling: echoes written before the loss are kept, later ones read 0
        case, mask, fm, fov, xyz0 = cases.stuck()
        case.scales, case.scale_type = [0.0, 1.0, 3.0], cases.po.SCALE_GRADIENT
        case.gradient_tp, case.gradX_mTm, case.gradY_mTm, case.gradZ_mTm = [10, 11, 12, 100], [20.0, 20.0, 20.0, -5.0], [0.0, 1.0, 0.0, 0.0], [3.0, 0.0, 0.0, 0.0]
    elif name == "long_lobes":  # 6400 gradient samples: neither the sequence tables nor the pre-multiplied gradient table fit shared memory (the global-memory paths)
        case, mask, fm, fov, xyz0 = cases.pgse(n_spins=160)
        lobe = list(range(50, 3250)) + list(range(3400, 6600))
        case.TR_us, case.TE_tp, case.RF_tp = 6700 * 50, [6650], [0, 3300]
        case.gradient_tp, case.gradX_mTm, case.gradY_mTm, case.gradZ_mTm = lobe, [0.6] * len(lobe), [0.3] * len(lobe), [0.0] * len(lobe)
    else:
        case, mask, fm, fov, xyz0 = cases.ALL[name]()
    cfg = cases.to_simconfig(case)
    K = case.n_scales
    with sw.Engine(0) as e:
        e.set_phantom(mask, fm, fov)
        e.set_sequence(cfg)
        e.set_spins(xyz0)
        st1 = e.run_device(mode=sw.MODE_FAST, flags=sw.OUT_ALL | sw.RUN_STATS)
        one = e.download() + (e.sums(),)
        stk = e.run_device(mode=sw.MODE_FAST, flags=sw.OUT_ALL | sw.RUN_STATS | sw.RUN_NO_ONEWALK)
        per = e.download() + (e.sums(),)
        # ensemble sums only (no per-spin output): the walk goes chunk after chunk through a bounded magnetisation buffer (SWK_MULTI_CHUNK: 256 slots)
        monkeypatch.setenv("SWK_MULTI_CHUNK", "1")
        e.run_device(mode=sw.MODE_FAST, flags=0)
        assert np.array_equal(e.sums(), one[3]), name  # (fixed-point sums: exact whatever the chunking)
        monkeypatch.delenv("SWK_MULTI_CHUNK")
        monkeypatch.setenv("SWK_SLICES", "3")
        host = e.run(xyz0, mode=sw.MODE_FAST)
    (m1, x1, t, s1), (m1k, x1k, tk, sk) = one, per
    assert np.array_equal(x1, x1k) and np.array_equal(t, tk), name
    for k in range(1, K):
        assert np.array_equal(x1[k], x1[0]) and np.array_equal(t[k], t[0])
    np.testing.assert_allclose(m1, m1k, rtol=0, atol=3e-5, err_msg=name)
    assert np.abs(m1k).max() > 0.1 and (K < 2 or np.abs(m1k[0] - m1k[-1]).max() > 1e-3)  # the scales do differ
    np.testing.assert_allclose(s1[..., :3], sk[..., :3], rtol=0, atol=3e-5 * case.n_spins)
    assert np.array_equal(s1[..., 3], sk[..., 3])
    for key in ("steps", "mask_gathers", "field_gathers", "rejects", "lost"):
        assert st1[key] == stk[key], (name, key)
    if name == "stuck_gradient":
        assert st1["lost"] > 0
    assert np.array_equal(host["M1"], m1) and np.array_equal(host["XYZ1"], x1) and np.array_equal(host["T"], t) and np.array_equal(host["sums"], s1)


def test_scales_do_not_depend_on_their_neighbours(sw):
    """Like the reference, which launches every scale on its own (monte_carlo.cu:273-337), the result of a scale does not depend on which
    other scales are simulated with it (block composition, group size of the SHARED variant): 13 scales at once == each scale alone."""
    case, mask, fm, fov, xyz0 = cases.gre(n_spins=300, scales=tuple(0.05 * 1.5 ** i for i in range(13)))
    cfg = cases.to_simconfig(case)
    with sw.Engine(0) as e:
        e.set_phantom(mask, fm, fov)
        e.set_sequence(cfg)
        e.set_spins(xyz0)
        e.run_device(mode=sw.MODE_FAST, flags=sw.OUT_ALL)
        m1, x1, t = e.download()
        sums = e.sums()
        for k in (0, 5, 12):
            e.run_device(scales=[case.scales[k]], mode=sw.MODE_FAST, flags=sw.OUT_ALL)
            a, b, c = e.download()
            assert np.array_equal(a[0], m1[k]) and np.array_equal(b[0], x1[k]) and np.array_equal(c[0], t[k])
            assert np.array_equal(e.sums()[0], sums[k])


def test_row_copies_beyond_the_pitch_limit(sw, monkeypatch):
    """cudaMemcpy2D rejects pitches above cudaDeviceProp::memPitch (2^31 - 1): per-scale blocks of 2 GiB or more (1e8 spins x 2 echoes,
    trajectories) are copied scale by scale instead.  SWK_MEMPITCH fakes a tiny limit so that a test-sized run takes that path: several engines
    filling one set of host arrays (swk_set_host_rows), plain and sliced."""
    monkeypatch.setenv("SWK_MEMPITCH", "64")
    case, mask, fm, fov, xyz0 = cases.multi_echo(n_spins=1100)
    cfg = cases.to_simconfig(case)
    full = _run_engine(sw, case, mask, fm, fov, xyz0, sw.MODE_FAST)
    K, S, E = case.n_scales, case.n_spins, case.n_TE
    for slices in (None, "3"):
        if slices:
            monkeypatch.setenv("SWK_SLICES", slices)
        out = (np.zeros((K, S, E, 3), np.float32), np.zeros((K, S, 1, 3), np.float32), np.zeros((K, S, E), np.uint8))
        for first, n in ((0, 300), (300, 800)):
            with sw.Engine(0) as e:
                e.set_phantom(mask, fm, fov)
                e.set_sequence(cfg)
                e._ck(e._lib.swk_set_host_rows(e._h, S, first))
                e.run(xyz0[first:first + n], spin_first=first, mode=sw.MODE_FAST, out=out)
        for a, key in zip(out, ("M1", "XYZ1", "T")):
            assert np.array_equal(a, full[key]), (slices, key)
