#!/usr/bin/env python
"""bench.py — spin-steps/s of the `sim` hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c1|c3|c3r|c4|c5]
  python bench.py --workload ph-c2|ph-c5|ph-c3|ph-mesh ...     the phantom generator on the same recipes: voxels/s

A "step" is one pass of the hot path over the whole workload: all spins x all scales x all
timepoints of one phantom (what one iteration of the reference's phantom loop does,
src/sim/monte_carlo.cu:227-349).  Default workload = BASELINE.json configs[1]: SE BOLD (config/se.ini),
600^3 cylinder vessel phantom with susceptibility field map, 1e7 spins, the 50 FoV scales of
config/config_default.ini => 4.0e11 spin-steps per pass.

  value  whole-job throughput, inputs resident in HBM, device time (CUDA events on the engine's stream,
         max over ranks), output zero-fill included.
  e2e    same metric through swk_run(): HOST buffers in (pinned XYZ0), HOST buffers out (pinned M1, XYZ1, T
         + sums), H2D and D2H inside the timed region.
Under torchrun (N > 1) every rank simulates its own shard of N x spins (weak scaling), the phantom is
replicated, and the per-echo ensemble sums are all-reduced with NCCL.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DEFAULT_SCALES = [0.0125, 0.0147, 0.0173, 0.0204, 0.0240, 0.0283, 0.0333, 0.0392, 0.0462, 0.0544, 0.0641, 0.0754, 0.0888,
                  0.1046, 0.1231, 0.1450, 0.1707, 0.2010, 0.2367, 0.2787, 0.3282, 0.3865, 0.4551, 0.5358, 0.6309, 0.7429,
                  0.8748, 1.0301, 1.2129, 1.4282, 1.6817, 1.9803, 2.3318, 2.7456, 3.2330, 3.8069, 4.4826, 5.2783, 6.2152,
                  7.3184, 8.6174, 10.1470, 11.9481, 14.0689, 16.5662, 19.5067, 22.9692, 27.0463, 31.8471, 37.5000]  # config_default.ini:78-127


def workload(name: str, n_spins: int | None, n_scales: int | None):
    """Returns (SimConfig kwargs, phantom spec, description).  INI values of config/*.ini + config_default.ini."""
    base = dict(timestep_us=50, B0=9.4, seed=10, cross_fov=0, max_iterations=10000, diffusivity=[1e-9, 1e-9],
                T1_ms=[2200.0, 2200.0], T2_ms=[41.0, 41.0], pXY=[1.0, 0.0, 0.0, 1.0], scales=list(DEFAULT_SCALES), scale_type=0)
    if name == "c2":
        cfg = dict(base, TR_us=40000, TE_us=[20000], RF_FA_deg=[90.0, 180.0], RF_PH_deg=[0.0, 90.0], RF_T_us=[0, 10000])
        ph = dict(kind="cylinder", n=600, fov_um=600.0, radius_um=8.0, bvf=4.0, Y=0.78, seed=0)
        S, desc = 10_000_000, "C2 SE BOLD (config/se.ini), 600^3 cylinder phantom r=8um BVF 4% + fieldmap, 1e7 spins, 50 FoV scales"
    elif name == "c1":
        cfg = dict(base, TR_us=40000, TE_us=[20000], RF_FA_deg=[90.0], RF_PH_deg=[0.0], RF_T_us=[0])
        ph = dict(kind="cylinder", n=100, fov_um=100.0, radius_um=8.0, bvf=4.0, Y=0.78, seed=0)
        S, desc = 100_000, "C1 GRE BOLD (config/gre.ini), 100^3 cylinder phantom, 1e5 spins, 50 FoV scales"
    elif name == "c4":
        cfg = dict(base, TR_us=10000, TE_us=[5000], RF_FA_deg=[16.0], RF_PH_deg=[0.0], RF_T_us=[0], n_dummy_scan=-1,
                   linear_phase_cycling=180.0, scales=[1.0])
        ph = dict(kind="cylinder", n=600, fov_um=600.0, radius_um=8.0, bvf=4.0, Y=0.78, seed=0)
        S, desc = 10_000_000, "C4 bSSFP (config/ssfp.ini), 1101 TRs x 200 steps, 600^3 cylinder phantom, 1e7 spins, 1 scale"
    elif name == "c5":
        cfg = dict(base, TR_us=40000, TE_us=[20000], RF_FA_deg=[90.0], RF_PH_deg=[0.0], RF_T_us=[0])
        ph = dict(kind="cylinder", n=1000, fov_um=1000.0, radius_um=8.0, bvf=4.0, Y=0.78, seed=0)
        S, desc = 125_000_000, ("C5 GRE BOLD (config/gre.ini), 1000^3 cylinder phantom r=8um BVF 4% + fieldmap (9 GB per GPU), "
                                "1.25e8 spins per GPU (1e9 over 8), 50 FoV scales, ensemble sums only (no per-spin outputs)")
    elif name in ("c3", "c3r"):
        from spinwalk_b200.sequences import pgse

        seq = pgse([100.0 * i for i in range(1, 51)] + [0.0], (1.0, 0.0, 0.0), start_ms=15, delta_ms=10, DELTA_ms=20, timestep_us=50)
        restricted = name == "c3r"
        cfg = dict(base, TR_us=60050, TE_us=[60000], T1_ms=[9999999.0, 9999999.0], T2_ms=[9999999.0, 9999999.0],
                   pXY=[1.0, 0.05, 0.05, 1.0] if restricted else [1.0, 1.0, 1.0, 1.0], cross_fov=1, **seq)
        ph = dict(kind="spheres", n=400, fov_um=400.0, cell_um=40.0, vf=40.0, seed=0)
        S, desc = 10_000_000, ("C3 PGSE (dwi -b 100..5000,0 -v 1 0 0 -d 15 10 20), 400^3 sphere phantom (phantom -s -r -20 -v 40 -y -1 -e 0), no fieldmap, "
                               + ("P_XY = 0.05 (restricted)" if restricted else "P_XY = 1 (free diffusion)") + ", 1e7 spins, 51 gradient scales")
    else:
        raise SystemExit(f"unknown workload {name}")
    if n_spins:
        S = n_spins
    if n_scales:
        cfg["scales"] = cfg["scales"][:: max(1, len(cfg["scales"]) // n_scales)][:n_scales]
    cfg["n_spins"] = S
    return cfg, ph, desc


def make_phantom_2d(ph):
    """cylinders: one (x, y) plane (the phantom is invariant along z, phantom_cylinder.cpp:183-275); spheres: the 3-D mask."""
    from spinwalk_b200.phantoms import cylinder_phantom, sphere_lattice_phantom

    if ph["kind"] == "spheres":
        return sphere_lattice_phantom(ph["n"], ph["fov_um"], ph["cell_um"], ph["vf"], ph["seed"])
    return cylinder_phantom(ph["n"], ph["fov_um"], radius_um=ph["radius_um"], bvf_pct=ph["bvf"], Y=ph["Y"], seed=ph["seed"], planar=True)


def phantom_spec(ph):
    """The `spinwalk phantom` invocation of a workload's phantom as a spinwalk_b200.phantom_gen.PhantomSpec."""
    from spinwalk_b200 import phantom_gen as pg

    if ph["kind"] == "spheres":  # demo/spinwalk_dwi.ipynb: -s -r -20 -v 40 -y -1 -e 0
        return pg.PhantomSpec(shape=pg.SHAPE_SPHERE, fov_um=ph["fov_um"], resolution=ph["n"], oxy_level=-1.0, radius_um=-20.0, volume_fraction=ph["vf"], seed=ph["seed"])
    return pg.PhantomSpec(shape=pg.SHAPE_CYLINDER, fov_um=ph["fov_um"], resolution=ph["n"], oxy_level=ph["Y"], radius_um=ph["radius_um"], volume_fraction=ph["bvf"],
                          orientation_deg=90.0, seed=ph["seed"])


def full_phantom(ph, mask2, fm2):
    """host arrays [n, n, n] of the whole phantom (CPU baseline legs)."""
    n = ph["n"]
    if mask2.ndim == 3:
        return mask2, fm2
    mask = np.ascontiguousarray(np.broadcast_to(mask2[:, :, None], (n, n, n)))
    fm = None if fm2 is None else np.ascontiguousarray(np.broadcast_to(fm2[:, :, None], (n, n, n)))
    return mask, fm


def make_positions(S, fov, seed, first=0):
    """uniform in [1%,99%] of the FoV (distribution of monte_carlo.cu:142-151); numpy stream, chunked by global id."""
    rng = np.random.default_rng([seed, first])
    x = rng.random((S, 3), dtype=np.float32)
    f = np.asarray(fov, np.float32)
    return x * (np.float32(0.98) * f) + np.float32(0.01) * f


_JSON_OUT = None


def protect_stdout():
    """Only the JSON line may reach stdout: file descriptor 1 is pointed at stderr for everything else this process or its libraries
    print (NCCL's version banner goes straight to fd 1), and emit() writes to the saved original."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], threading.Event()
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows if len(r) > 3 + i)]
        pw = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "power_w_max": max(pw) if pw else None, "samples": len(self.rows)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def oracle_case(cfg_kw, n, fov, n_spins=None, scales=None):
    """SimConfig keyword arguments of a workload -> oracle.pyoracle.Case (times in timepoints, config_reader.cpp:39-46) for the
    reference legs (CPU reference, reference cu_sim) and the parity tests."""
    from oracle import pyoracle as po

    dt = cfg_kw["timestep_us"]
    tp = lambda us: [int(t) // dt for t in us]  # noqa: E731
    return po.Case(fov=tuple(fov), phantom_size=(n, n, n), n_spins=n_spins or cfg_kw["n_spins"], TR_us=cfg_kw["TR_us"], timestep_us=dt,
                   seed=cfg_kw["seed"], B0=cfg_kw["B0"], TE_tp=tp(cfg_kw["TE_us"]), RF_FA_deg=cfg_kw["RF_FA_deg"], RF_PH_deg=cfg_kw["RF_PH_deg"],
                   RF_tp=tp(cfg_kw["RF_T_us"]), n_dummy_scan=cfg_kw.get("n_dummy_scan", 0), linear_phase_cycling=cfg_kw.get("linear_phase_cycling", 0.0),
                   gradient_tp=tp(cfg_kw.get("gradient_T_us", [])), gradX_mTm=cfg_kw.get("gradient_X_mTm", []), gradY_mTm=cfg_kw.get("gradient_Y_mTm", []),
                   gradZ_mTm=cfg_kw.get("gradient_Z_mTm", []), diffusivity=cfg_kw["diffusivity"], T1_ms=cfg_kw["T1_ms"], T2_ms=cfg_kw["T2_ms"], pXY=cfg_kw["pXY"],
                   scales=list(cfg_kw["scales"] if scales is None else scales), scale_type=cfg_kw["scale_type"], cross_fov=cfg_kw["cross_fov"],
                   max_iterations=cfg_kw["max_iterations"])


def cpu_reference(cfg_kw, ph, mask2, fm2, fov, target_s=15.0, threads=None):
    """The reference's own CPU implementation of the path (oracle/_ref/libswref_cpu.so = unmodified kernels.cu built
    by g++, std::mt19937 arithmetic; falls back to the C port) on a bounded sample of the same workload."""
    from oracle import pyoracle as po

    n = ph["n"]
    mask, fm = full_phantom(ph, mask2, fm2)
    threads = threads or os.cpu_count() or 1
    kind = "reference" if po.have_ref_cpu() else "port"
    if kind == "port":
        po.build(ref=False)

    def run(n_spins, scales):
        c = oracle_case(cfg_kw, n, fov, n_spins, scales)
        x0 = make_positions(n_spins, fov, cfg_kw["seed"])
        f = po.run_ref if kind == "reference" else po.run_oracle
        r = f(c, fm, mask, x0, flavour=po.RNG_MT19937, threads=threads)
        return c.total_steps(), r["seconds"]

    scales = cfg_kw["scales"]
    steps, sec = run(max(threads * 8, 256), scales[:: max(1, len(scales) // 5)][:5])  # calibration
    rate = steps / max(sec, 1e-6)
    per_spin = len(scales) * (steps / (max(threads * 8, 256) * min(5, len(scales))))
    n_spins = int(max(threads * 8, min(cfg_kw["n_spins"], rate * target_s / per_spin)))
    steps, sec = run(n_spins, scales)
    if sec < 0.5 * target_s and n_spins < cfg_kw["n_spins"]:  # the short calibration under-estimates the rate (thread start-up): size once more
        n_spins = int(min(cfg_kw["n_spins"], n_spins * target_s / max(sec, 1e-3)))
        steps, sec = run(n_spins, scales)
    return {"value": steps / sec, "unit": "spin-steps/s", "cores": threads, "kind": kind,
            "sample": f"first {n_spins} spins x all {len(scales)} scales of the workload ({steps:.3g} spin-steps, {sec:.1f} s); "
                      f"low spin ids keep mt19937::discard(seed+spin) cheap, which flatters the CPU reference (SURVEY App. B-3)"}, steps, sec


def reference_cuda(cfg_kw, ph, mask2, fm2, fov, device, target_s=10.0):
    """The reference's EXISTING CUDA kernel (its untouched kernels.cu compiled for sm_100a into oracle/_ref/libswref_cuda.so, launched
    once per scale with a device sync like monte_carlo.cu:273-337) on a bounded sample of the same workload, same GPU.  Kernel time only
    (CUDA events around the launches); uploads and downloads of the harness are not counted."""
    from oracle import pyoracle as po

    if not po.have_ref_cuda():
        return None
    n = ph["n"]
    mask, fm = full_phantom(ph, mask2, fm2)

    def run(n_spins):
        c = oracle_case(cfg_kw, n, fov, n_spins)
        r = po.run_ref_cuda(c, fm, mask, make_positions(n_spins, fov, cfg_kw["seed"]), device=device)
        return c.total_steps(), r["kernel_ms"] * 1e-3

    steps, sec = run(200_000)  # calibration (also warms the context up)
    n_spins = int(min(cfg_kw["n_spins"], 4_000_000, max(200_000, 200_000 * target_s / max(sec, 1e-3))))
    steps, sec = run(n_spins)
    return {"value": steps / sec, "unit": "spin-steps/s", "kind": "reference cu_sim (src/sim/kernels.cu:56-63) built for sm_100a, one launch + sync per scale",
            "sample": f"first {n_spins} spins x all {len(cfg_kw['scales'])} scales of the workload ({steps:.3g} spin-steps, {sec:.2f} s of kernel time)"}


PHANTOM_RECIPES = {  # `spinwalk phantom` invocations behind the BASELINE configs
    "ph-c2": (dict(shape=0, fov_um=600.0, resolution=600, radius_um=8.0, volume_fraction=4.0, Y=0.78, orientation_deg=90.0, seed=0),
              "phantom -c -r 8 -v 4 -y 0.78 -n 90 -f 600 -z 600 -e 0 (C2's 600^3 vessel phantom: mask + field map)"),
    "ph-c5": (dict(shape=0, fov_um=1000.0, resolution=1000, radius_um=8.0, volume_fraction=4.0, Y=0.78, orientation_deg=90.0, seed=0),
              "phantom -c -r 8 -v 4 -y 0.78 -n 90 -f 1000 -z 1000 -e 0 (C5's 1000^3 vessel phantom: mask + field map, 5 GB)"),
    "ph-c3": (dict(shape=1, fov_um=400.0, resolution=400, radius_um=-20.0, volume_fraction=40.0, Y=-1.0, seed=0),
              "phantom -s -r -20 -v 40 -y -1 -f 400 -z 400 -e 0 (C3's 400^3 permeable-sphere phantom: mask only, 44k spheres)"),
    "ph-s256": (dict(shape=1, fov_um=256.0, resolution=256, radius_um=-20.0, volume_fraction=30.0, Y=0.78, seed=0),
                "phantom -s -r -20 -v 30 -y 0.78 -f 256 -z 256 -e 0 (spheres with dipole field map)"),
}


def mesh_bench(args):
    """`spinwalk phantom -p` on an 81 920-triangle icosphere (400 um across) in a 512 um FoV at 512^3 voxels: swk_phantom_mesh."""
    import tempfile

    from spinwalk_b200.phantoms import icosphere_mesh, write_ply

    fov, n = 512.0, 512
    v, f = icosphere_mesh(6, 0.2)
    desc = f"phantom -p -i icosphere.ply -f 512 -z 512 ({len(f)} triangles, sphere of 400 um)"
    V = n ** 3
    if args.impl == "reference":
        from oracle import pyphantom as pp

        if int(os.environ.get("RANK", 0)) != 0:
            return
        path = os.path.join(tempfile.mkdtemp(), "ico.ply")
        write_ply(path, v, f)
        nc = 256
        ts = []
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            pp.reference_mesh(fov, nc, path)
            if i >= args.warmup:
                ts.append(time.perf_counter() - t0)
        val = nc ** 3 * len(ts) / sum(ts)
        emit({"impl": "reference", "metric": "voxels/s", "value": val, "unit": "voxels/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": 1e3 * sum(ts) / len(ts), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64+f32", "data": "synthetic",
                          "config": {"workload": desc}, "gpu_launches": 0,
                          "cpu_baseline": {"value": val, "unit": "voxels/s", "cores": 1, "kind": "reference", "sample": f"the same mesh at {nc}^3 voxels"},
                          "e2e": {"value": val, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return
    import torch

    from spinwalk_b200 import phantom_gen as pg

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the phantom generator has no CPU path for the voxel fill")
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    for _ in range(args.warmup):
        pg.generate_mesh(fov, n, v, f, device=local_rank)
    ker_ms, prep_ms = 0.0, 0.0
    with ClockSampler(local_rank) as clk:
        for _ in range(args.steps):
            mask, _, st = pg.generate_mesh(fov, n, v, f, device=local_rank)
            ker_ms += st["kernel_ms"]
            prep_ms += st["place_ms"]
    t0 = time.perf_counter()  # e2e outside the sampler: its nvidia-smi forks stall the host thread for longer than the call takes
    for _ in range(args.steps):
        pg.generate_mesh(fov, n, v, f, device=local_rank)
    e2e_s = time.perf_counter() - t0
    peak, peak_src = measured_peaks()
    achieved = V / (ker_ms / args.steps * 1e-3) / 1e9
    line = {"metric": "voxels/s", "value": V * args.steps / (ker_ms * 1e-3), "unit": "voxels/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ker_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64+f32", "data": "synthetic",
            "config": {"workload": desc, "voxels": V, "triangles": len(f), "inside_pct": st["volume_fraction"], "host_prep_ms": prep_ms / args.steps,
                       "l2": "mask (134 MB) larger than L2, rewritten every pass"},
            "clocks": clk.summary(), "gpu_launches": args.steps,
            "e2e": {"value": V * args.steps / e2e_s, "unit": "voxels/s", "h2d_bytes_per_step": int(v.nbytes + f.nbytes), "d2h_bytes_per_step": V, "ms_per_step": 1e3 * e2e_s / args.steps,
                    "api": "swk_phantom_mesh (C-ABI) with host buffers: mesh in (leaf boxes + row binning on the host), mask out"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                         "kernel": "swk::phantom::mesh_fill_kernel", "algorithmic_bytes_per_launch": V, "kernel_ms_per_launch": ker_ms / args.steps,
                         "note": "1 B written per voxel; the kernel is bound by the FP64 hit test (about 30 double operations per voxel-candidate pair), not by HBM"}}
    if not args.no_cpu_baseline:
        try:
            from oracle import pyphantom as pp

            if pp.have_ref():
                path = os.path.join(tempfile.mkdtemp(), "ico.ply")
                write_ply(path, v, f)
                nc = 256
                t0 = time.perf_counter()
                pp.reference_mesh(fov, nc, path)
                sec = time.perf_counter() - t0
                line["cpu_baseline"] = {"value": nc ** 3 / sec, "unit": "voxels/s", "cores": 1, "kind": "reference",
                                        "sample": f"the same mesh at {nc}^3 voxels with the reference's phantom::ply (its std::execution::par_unseq loop runs serially without TBB): {sec:.1f} s"}
        except Exception as ex:
            line["cpu_baseline"] = {"value": None, "unit": "voxels/s", "cores": 1, "kind": "reference", "sample": f"failed: {ex}"}
    emit(line)


def phantom_bench(args):
    """The phantom generator on one GPU: `value` = voxels/s of the device voxel fill into device-resident buffers (CUDA events),
    `e2e` = swk_phantom_generate with HOST buffers (placement + fill + D2H), `cpu_baseline` = the reference's own generator classes
    (oracle/_ref/libswref_gen_omp.so: unmodified src/phantom/*.cpp with OpenMP, all host cores) on the same recipe."""
    kw, desc = PHANTOM_RECIPES[args.workload]
    n = kw["resolution"]
    V = n ** 3
    if args.impl == "reference":
        from oracle import pyphantom as pp

        if int(os.environ.get("RANK", 0)) != 0:
            return
        ts = []
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            pp.reference(omp=pp.have_ref(omp=True), **kw)
            if i >= args.warmup:
                ts.append(time.perf_counter() - t0)
        v = V * len(ts) / sum(ts)
        emit({"impl": "reference", "metric": "voxels/s", "value": v, "unit": "voxels/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": 1e3 * sum(ts) / len(ts), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64",
                          "data": "synthetic", "config": {"workload": desc}, "gpu_launches": 0,
                          "cpu_baseline": {"value": v, "unit": "voxels/s", "cores": os.cpu_count(), "kind": "reference", "sample": "the whole phantom"},
                          "e2e": {"value": v, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return
    import torch

    from spinwalk_b200 import phantom_gen as pg

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the phantom generator has no CPU path for the voxel fill")
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    spec = pg.PhantomSpec(shape=kw["shape"], fov_um=kw["fov_um"], resolution=n, oxy_level=kw["Y"], radius_um=kw["radius_um"], volume_fraction=kw["volume_fraction"],
                          orientation_deg=kw.get("orientation_deg", 90.0), seed=kw["seed"])
    mask_d = torch.empty((n, n, n), dtype=torch.uint8, device=dev)
    fm_d = torch.empty((n, n, n), dtype=torch.float32, device=dev) if spec.has_fieldmap else None
    out_bytes = V * (5 if spec.has_fieldmap else 1)
    for _ in range(args.warmup):
        pg.generate(spec, out=(mask_d, fm_d))
    ker_ms, place_ms, launches = 0.0, 0.0, 0
    with ClockSampler(local_rank) as clk:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            _, _, _, st = pg.generate(spec, out=(mask_d, fm_d))
            ker_ms += st["kernel_ms"]
            place_ms += st["place_ms"]
            launches += st["n_launches"]
        torch.cuda.synchronize(dev)
        wall_s = time.perf_counter() - t0
    clocks = clk.summary()
    del mask_d, fm_d
    torch.cuda.empty_cache()
    # e2e: the call `spinwalk phantom` makes — host buffers out (allocated and touched once, like the caller's std::vector)
    host = (np.zeros((n, n, n), np.uint8), np.zeros((n, n, n), np.float32) if spec.has_fieldmap else None)
    pg.generate(spec, device=local_rank, out_host=host)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pg.generate(spec, device=local_rank, out_host=host)
    e2e_s = time.perf_counter() - t0
    del host
    peak, peak_src = measured_peaks()
    achieved = out_bytes / (ker_ms / args.steps * 1e-3) / 1e9
    line = {"metric": "voxels/s", "value": V * args.steps / (ker_ms * 1e-3), "unit": "voxels/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ker_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
            "config": {"workload": desc, "voxels": V, "shapes": st["n_shapes"], "volume_fraction_pct": st["volume_fraction"], "placement_ms_host": place_ms / args.steps,
                       "exact_columns": st["exact_columns"], "l2": "outputs larger than L2 are rewritten every pass" if out_bytes > 126e6 else "output fits in L2"},
            "clocks": clocks, "wall_ms_per_step": 1e3 * wall_s / args.steps, "gpu_launches": launches,
            "e2e": {"value": V * args.steps / e2e_s, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": out_bytes, "ms_per_step": 1e3 * e2e_s / args.steps,
                    "api": "swk_phantom_generate (C-ABI) with pageable host buffers: placement + voxel fill + D2H of mask and field map"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                         "kernel": "swk::phantom::slab_broadcast_bulk_kernel" if kw["shape"] == 0 else "swk::phantom::sphere_fill_kernel",
                         "algorithmic_bytes_per_launch": out_bytes, "kernel_ms_per_launch": ker_ms / args.steps,
                         "note": "5 B written per voxel (1 B mask + 4 B field) or 1 B without field map; the sphere kernel is bound by two IEEE double "
                                 "divisions per (voxel, sphere) pair, not by HBM"}}
    if not args.no_cpu_baseline:
        try:
            from oracle import pyphantom as pp

            have_omp = pp.have_ref(omp=True)
            if have_omp or pp.have_ref():
                # bounded sample: at most 600^3 voxels at the recipe's voxel size (the reference needs 12 B of host grid per voxel and minutes beyond that)
                nc = min(n, 600)
                kw_cpu = dict(kw, resolution=nc, fov_um=kw["fov_um"] * nc / n)
                t0 = time.perf_counter()
                pp.reference(omp=have_omp, **kw_cpu)
                sec = time.perf_counter() - t0
                line["cpu_baseline"] = {"value": nc ** 3 / sec, "unit": "voxels/s", "cores": os.cpu_count() if have_omp else 1, "kind": "reference",
                                        "sample": f"{'the whole phantom' if nc == n else f'the same recipe at {nc}^3 voxels'} once with the reference's generator classes "
                                                  f"({'OpenMP' if have_omp else 'serial'} build): {sec:.1f} s"}
            else:
                import subprocess as sp

                sp.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True)
                zw = (0, max(1, n // 16))
                t0 = time.perf_counter()
                pp.oracle(zwin=zw, **kw)
                sec = time.perf_counter() - t0
                line["cpu_baseline"] = {"value": n * n * zw[1] / sec, "unit": "voxels/s", "cores": 1, "kind": "port", "sample": f"z slices [0, {zw[1]}) of the phantom: {sec:.1f} s"}
        except Exception as ex:
            line["cpu_baseline"] = {"value": None, "unit": "voxels/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {ex}"}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--spins", type=int, default=0, help="override spins per GPU (debug; makes the number non-headline)")
    ap.add_argument("--scales", type=int, default=0, help="override number of scales (debug)")
    ap.add_argument("--mode", default="fast", choices=["fast", "compat"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    protect_stdout()
    if args.workload == "ph-mesh":
        return mesh_bench(args)
    if args.workload in PHANTOM_RECIPES:
        return phantom_bench(args)

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    cfg_kw, ph, desc = workload(args.workload, args.spins or None, args.scales or None)
    S_per_gpu = cfg_kw["n_spins"]
    K = len(cfg_kw["scales"])

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        mask2, fm2, fov = make_phantom_2d(ph)
        vals, samples = [], None
        tgt = 12.0
        for i in range(args.warmup + args.steps):
            cb, steps, sec = cpu_reference(cfg_kw, ph, mask2, fm2, fov, target_s=tgt if i >= args.warmup else 3.0)
            if i >= args.warmup:
                vals.append((steps, sec))
                samples = cb
        tot_steps = sum(v[0] for v in vals)
        tot_sec = sum(v[1] for v in vals)
        v = tot_steps / tot_sec
        samples["value"] = v
        emit({"impl": "reference", "metric": "spin-steps/s", "value": v, "unit": "spin-steps/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_sec / max(1, args.steps),
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64",
                          "data": "synthetic", "config": {"workload": desc, "note": "each step = bounded sample of the workload on host cores"},
                          "cpu_baseline": samples, "gpu_launches": 0,
                          "e2e": {"value": v, "unit": "spin-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return

    # ------------------------------------------------------------------ our arm (GPU)
    import torch
    import torch.distributed as dist

    import spinwalk_b200 as sw
    from spinwalk_b200 import sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    mode = sw.MODE_FAST if args.mode == "fast" else sw.MODE_COMPAT
    cfg_kw_global = dict(cfg_kw, n_spins=S_per_gpu * world)  # weak scaling: global population grows with N
    cfg = sw.SimConfig(**cfg_kw_global)
    # the phantom: the reference's own recipe (`spinwalk phantom ...`), generated on every rank's device by the product generator
    # (include/spinwalk_phantom.h: bit-identical to the reference's generator, tests/test_phantom_gpu.py) — it never visits the host
    eng = sw.Engine(local_rank)
    gen = eng.generate_phantom(phantom_spec(ph))
    fov = eng.fov

    spin_first, n_local = sharding.shard_range(S_per_gpu * world, rank, world)  # weak scaling: S_per_gpu spins on every rank
    assert n_local == S_per_gpu
    xyz0_pin = torch.empty((S_per_gpu, 3), dtype=torch.float32, pin_memory=True)
    xyz0_pin.numpy()[:] = make_positions(S_per_gpu, fov, cfg.seed, spin_first)
    per_spin_out = args.workload != "c5"  # C5: 1e9 spins x 50 scales of per-spin output would be 650 GB: the reduce is the product
    out_flags = sw.OUT_ALL if per_spin_out else 0
    eng.set_sequence(cfg)
    eng.set_spins(xyz0_pin.numpy(), None, spin_first)
    E, ns = cfg.n_TE, cfg.n_substrate
    sums_d = torch.zeros((K, E, ns, 4), dtype=torch.float64, device=dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def one_pass():
        st = eng.run_device(mode=mode, flags=out_flags, d_sums_ptr=sums_d.data_ptr())
        sharding.allreduce_sums(sums_d)  # the one collective of the path: a few KB of per-echo ensemble sums (NCCL)
        return st

    # counters (voxel changes etc.) for the roofline's algorithmic bytes: same inputs, STATS kernel variant, untimed
    st_counts = eng.run_device(mode=mode, flags=out_flags | sw.RUN_STATS, d_sums_ptr=sums_d.data_ptr())

    for _ in range(args.warmup):
        one_pass()
    barrier()
    dev_ms, ker_ms = 0.0, 0.0
    with ClockSampler(local_rank) as clk:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            st = one_pass()
            dev_ms += st["device_ms"]
            ker_ms += st["kernel_ms"]
        barrier()
        wall_s = time.perf_counter() - t0
    clocks = clk.summary()
    t = torch.tensor([dev_ms, ker_ms, wall_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, ker_ms, wall_ms = (float(v) for v in t.tolist())
    steps_per_pass_rank = S_per_gpu * K * (eng.n_dummy_scan + 1) * cfg.n_timepoints
    total_steps = steps_per_pass_rank * world * args.steps
    value = total_steps / (dev_ms * 1e-3)

    # ---- end-to-end through swk_run with host buffers
    e2e = None
    if not args.no_e2e:
        if per_spin_out:
            out = (torch.empty((K, S_per_gpu, E, 3), dtype=torch.float32, pin_memory=True),
                   torch.empty((K, S_per_gpu, eng.trj, 3), dtype=torch.float32, pin_memory=True),
                   torch.empty((K, S_per_gpu, E), dtype=torch.uint8, pin_memory=True))
            out_np = tuple(o.numpy() for o in out)
        else:
            out, out_np = (), None
        h2d = xyz0_pin.numel() * 4 + K * 4
        d2h = sum(o.numel() * o.element_size() for o in out) + K * E * ns * 4 * 8
        eng.run(xyz0_pin.numpy(), None, spin_first, mode=mode, out=out_np, outputs=per_spin_out, stats=False)  # warm
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            r = eng.run(xyz0_pin.numpy(), None, spin_first, mode=mode, out=out_np, outputs=per_spin_out, stats=False)
            if world > 1:
                sums_d.copy_(torch.from_numpy(r["sums"]))
                sharding.allreduce_sums(sums_d)
        barrier()
        e2e_s = time.perf_counter() - t0
        te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": total_steps / float(te.item()), "unit": "spin-steps/s", "h2d_bytes_per_step": h2d * world,
               "d2h_bytes_per_step": d2h * world, "ms_per_step": 1e3 * float(te.item()) / args.steps,
               "api": "swk_run (C-ABI) with pinned host buffers: XYZ0 in; " + ("M1, XYZ1, T, sums out" if per_spin_out else "sums out")}
        del out, out_np

    # ---- the voxel fetch's own roofline, measured live on this device and this phantom (untimed diagnostic launch)
    try:
        probe = eng.probe_gather(threads_per_sm=2048, iters=2048)
    except Exception as ex:  # a diagnostic must never cost the bench line
        probe = {"gathers_per_s": None, "table_bytes": None, "error": str(ex)}

    # ---- roofline of the walk kernel: algorithmic bytes / mean launch duration
    peak, peak_src = measured_peaks()
    traffic, traffic_src = None, None
    try:  # DRAM bytes per launch of this very workload from the committed ncu capture (profiles/traffic.json)
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        ent = tj.get(f"{args.workload}:{args.mode}")
        if ent and not args.scales and (not args.spins or ent.get("spins_per_gpu")):
            # a capture taken with fewer spins scales linearly (every spin does the same work on average)
            traffic = ent["dram_bytes_per_launch"] * (S_per_gpu / ent["spins_per_gpu"] if ent.get("spins_per_gpu") else 1.0)
            traffic_src = ent["source"]
    except Exception:
        pass
    per_pass_bytes = (st_counts["mask_gathers"] * 1 + st_counts["field_gathers"] * 4
                      + S_per_gpu * K * (24 + ((13 * E + 12) if per_spin_out else 0)))
    ker_ms_per_launch = ker_ms / args.steps
    achieved = per_pass_bytes / (ker_ms_per_launch * 1e-3) / 1e9
    fetches_per_s = st_counts["mask_gathers"] / (ker_ms_per_launch * 1e-3)
    gather = {"voxel_fetches_per_s": fetches_per_s, "random_gather_peak_per_s": probe.get("gathers_per_s"),
              "frac_of_random_gather_peak": (fetches_per_s / probe["gathers_per_s"]) if probe.get("gathers_per_s") else None,
              "table_bytes": probe.get("table_bytes"),
              "hbm_64B_fetches_per_s": (traffic / 64 / (ker_ms_per_launch * 1e-3)) if traffic else None,
              # the probe's own HBM rate: its gathers minus the share a uniformly random access finds in L2 (L2 bytes / table bytes)
              "hbm_fetch_frac_of_probe": (traffic / 64 / (ker_ms_per_launch * 1e-3) / (probe["gathers_per_s"] * max(0.05, 1.0 - 126e6 / probe["table_bytes"])))
              if (traffic and probe.get("gathers_per_s") and probe.get("table_bytes")) else None,
              "note": "random_gather_peak = swk_probe_gather: dependent random 4-byte gathers over the same voxel table, nothing else "
                      "(HBM row-activation bound, DESIGN.md §5); voxel_fetches include L1/L2 hits of the larger FoV scales"}
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": traffic_src, "gather": gather,
                "peak_source": peak_src, "kernel": "swk::walk_fast_kernel" if mode == sw.MODE_FAST else "swk::walk_kernel<COMPAT>",
                "algorithmic_bytes_per_launch": per_pass_bytes, "kernel_ms_per_launch": ker_ms_per_launch,
                "bytes_per_spin_step": per_pass_bytes / steps_per_pass_rank,
                "p_voxel_change": st_counts["mask_gathers"] / max(1, st_counts["steps"]),
                "rejects_per_step": st_counts["rejects"] / max(1, st_counts["steps"]),
                "note": "gather-latency/issue-bound kernel: algorithmic bytes are ~1-5 B per spin-step, so the HBM fraction is "
                        "small by construction; see profiles/ for issue-slot and L2 sector counters"}

    line = {"metric": "spin-steps/s", "value": value, "unit": "spin-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if mode == sw.MODE_FAST else "f32+f64", "data": "synthetic",
            "config": {"workload": desc, "spins_per_gpu": S_per_gpu, "n_scales": K, "timepoints": cfg.n_timepoints,
                       "scans": eng.n_dummy_scan + 1, "spin_steps_per_pass": steps_per_pass_rank * world,
                       "rng": "philox4x32-10, one block per two steps + Box-Muller (SWK_MODE_FAST)" if mode == sw.MODE_FAST else "minstd_rand + erfcinvf (SWK_MODE_COMPAT, reference arithmetic)",
                       "l2": f"inputs larger than L2 (phantom {5 * ph['n'] ** 3 / 1e9:.2f} GB vs 126 MB)" if ph["n"] >= 600 else "phantom fits in L2; outputs (12.5 GB > L2) rewritten every pass",
                       "phantom": f"generated on the device by swk_generate_phantom: {gen['n_shapes']} shapes, volume fraction {gen['volume_fraction']:.3f} %, "
                                  f"voxel fill {gen['kernel_ms']:.2f} ms (bit-identical to the reference's `spinwalk phantom` for this recipe)",
                       "parallelism": f"spins sharded over {world} GPU(s), phantom replicated, NCCL all-reduce of per-echo sums",
                       **({"zslab": "SWK_ZSLAB=1: z-invariant phantom walked on its [nx][ny] slab (opt-in specialisation; the voxel table is then L1/L2-resident, NOT the default path and not the headline configuration)"} if os.environ.get("SWK_ZSLAB") else {})},
            "clocks": clocks, "wall_ms_per_step": wall_ms / args.steps, "gpu_launches": args.steps * st["n_launches"],
            "e2e": e2e, "roofline": roofline, "lost_spins": st_counts["lost"]}

    mask2 = fm2 = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.workload != "c5":  # (C5's 5 GB host phantom: use C2's baseline)
        mask2, fm2 = eng.get_phantom()  # the baselines below walk the very same voxels
        try:
            cb, _, _ = cpu_reference(cfg_kw, ph, mask2, fm2, fov, target_s=15.0)
            line["cpu_baseline"] = cb
        except Exception as ex:  # the baseline is a reported extra; never lose the GPU line over it
            line["cpu_baseline"] = {"value": None, "unit": "spin-steps/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}
    eng.close()
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.workload not in ("c5", "c4"):
        try:  # the reference's own CUDA kernel on this GPU, for context; never the thing measured above
            torch.cuda.empty_cache()
            line["reference_cuda"] = reference_cuda(cfg_kw, ph, mask2, fm2, fov, local_rank)
        except Exception as ex:
            line["reference_cuda"] = {"value": None, "sample": f"failed: {ex}"}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
