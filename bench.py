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


def workload_config(desc, ph, S_per_gpu, K, n_tp, scans, world, slab):
    """`config` of the JSON line: what the workload IS — the same dict, key for key, from the GPU arm and from `--impl reference`
    (everything that describes how an arm runs it goes to `details`)."""
    return {"workload": desc, "spins_per_gpu": int(S_per_gpu), "n_scales": int(K), "timepoints": int(n_tp), "scans": int(scans),
            "spin_steps_per_pass": int(S_per_gpu) * int(K) * int(n_tp) * int(scans) * int(world),
            "l2": ("per-spin outputs (12.5 GB > L2) rewritten every pass; the voxel table of this z-invariant phantom is its [nx][ny] slab (L1/L2 resident) — "
                   "`full_table` is the same workload on the full [nx][ny][nz] table (larger than L2)") if slab else
                  (f"inputs larger than L2 (voxel table {4 * ph['n'] ** 3 / 1e9:.2f} GB vs 126 MB)" if ph["n"] >= 400 else "phantom fits in L2; outputs rewritten every pass"),
            "parallelism": f"spins sharded over {world} GPU(s), phantom replicated, NCCL all-reduce of per-echo sums"}


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
    ap.add_argument("--no-extras", action="store_true", help="skip the sub-records (full_table, compat, scale_groups, non_invariant, gradient_scales, other_configs, north_star)")
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

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        mask2, fm2, fov = reference_phantom(ph)
        ref_case = oracle_case(cfg_kw, ph["n"], fov)
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
                          "data": "synthetic", "config": workload_config(desc, ph, cfg_kw["n_spins"], len(cfg_kw["scales"]), ref_case.n_timepoints, ref_case.n_dummy + 1,
                                                                         args.gpus, ph["kind"] != "spheres" and os.environ.get("SWK_NO_ZSLAB") is None),
                          "details": {"note": "each step = bounded sample of the workload on host cores (rank 0 only); the phantom is the recipe's own (oracle restatement of "
                                              "`spinwalk phantom`, bit-identical to what the GPU arm generates)"},
                          "cpu_baseline": samples, "gpu_launches": 0,
                          "e2e": {"value": v, "unit": "spin-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return

    # ------------------------------------------------------------------ our arm (GPU)
    import torch
    import torch.distributed as dist

    import spinwalk_b200 as sw

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = Walk(sw, torch, dist, dev, rank, world, local_rank)
    mode = sw.MODE_FAST if args.mode == "fast" else sw.MODE_COMPAT
    peak, peak_src = measured_peaks()

    # ---- the headline workload
    H = W.setup(args.workload, cfg_kw, ph)
    per_spin_out = args.workload != "c5"  # C5: 1e9 spins x 50 scales of per-spin output would be 650 GB: the reduce is the product
    out_flags = sw.OUT_ALL if per_spin_out else 0
    counts = W.counts(H, mode, out_flags)
    with ClockSampler(local_rank) as clk:
        T = W.timed(H, mode, out_flags, args.steps, args.warmup)
    clocks = clk.summary()
    K, E, ns, S_per_gpu = H["K"], H["E"], H["ns"], H["S"]
    value = H["steps_per_pass"] * world * args.steps / (T["dev_ms"] * 1e-3)

    e2e = None
    if not args.no_e2e:
        e2e = W.e2e(H, mode, per_spin_out, args.steps)

    # ---- roofline of the walk kernel: algorithmic bytes / mean launch duration; and the voxel fetch against its own ceiling
    roofline = W.roofline(H, counts, T, args.steps, per_spin_out, peak, peak_src, mode, args)
    if mode == sw.MODE_FAST and K >= 10 and cfg_kw["scale_type"] == 0 and not args.no_extras:
        try:
            roofline["gather"] = W.gather_roofline(H, mode)
        except Exception as ex:  # a diagnostic must never cost the bench line
            roofline["gather"] = {"error": str(ex)}

    line = {"metric": "spin-steps/s", "value": value, "unit": "spin-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": T["dev_ms"] / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if mode == sw.MODE_FAST else "f32+f64", "data": "synthetic",
            "config": workload_config(desc, ph, S_per_gpu, K, H["cfg"].n_timepoints, H["eng"].n_dummy_scan + 1, world, H["slab"]),
            "details": {"rng": "philox4x32-10 (one block per two rounds, shared by all FoV scales of a spin) + Box-Muller (SWK_MODE_FAST)" if mode == sw.MODE_FAST
                               else "minstd_rand + erfcinvf (SWK_MODE_COMPAT, reference arithmetic)",
                        "voxel_table": H["table"], "phantom": H["phantom_note"]},
            "clocks": clocks, "wall_ms_per_step": T["wall_ms"] / args.steps, "gpu_launches": args.steps * T["st"]["n_launches"],
            "e2e": e2e, "roofline": roofline, "lost_spins": counts["lost"]}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = W.cpu_baseline(H, cfg_kw, ph)
    W.close(H)
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.workload not in ("c5", "c4"):
        try:  # the reference's own CUDA kernel on this GPU, for context; never the thing measured above
            torch.cuda.empty_cache()
            mask2, fm2 = H["host_phantom"]
            line["reference_cuda"] = reference_cuda(cfg_kw, ph, mask2, fm2, H["fov"], local_rank)
        except Exception as ex:
            line["reference_cuda"] = {"value": None, "sample": f"failed: {ex}"}
    H.pop("host_phantom", None)

    # ---- sub-records (every N): the same engine on the configurations VERDICT r1 asked for, each a bounded, separately timed run
    if args.workload == "c2" and mode == sw.MODE_FAST and not args.no_extras and not args.spins and not args.scales:
        for name, fn in (("full_table", W.extra_full_table), ("compat", W.extra_compat), ("scale_groups", W.extra_scale_groups),
                         ("non_invariant", W.extra_non_invariant), ("gradient_scales", W.extra_gradient_scales), ("other_configs", W.extra_other_configs),
                         ("north_star", W.extra_north_star)):
            try:
                line[name] = fn(peak, peak_src, args)
            except Exception as ex:  # an extra must never cost the headline line
                line[name] = {"error": f"{type(ex).__name__}: {ex}"}
            torch.cuda.empty_cache()
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def reference_phantom(ph):
    """The workload's phantom for the CPU reference arm, WITHOUT a GPU: the recipe through the oracle's restatement of `spinwalk phantom`
    (bit-identical to the reference's generator and to the product's GPU generator, tests/test_phantom_oracle.py / test_phantom_gpu.py).
    Cylinder phantoms do not depend on z: one z slice is computed and the caller broadcasts it (full_phantom)."""
    from oracle import pyphantom as pp

    import subprocess as sp

    fov = (np.float32(ph["fov_um"]) * np.float32(1e-6),) * 3
    fov = tuple(float(f) for f in fov)
    try:
        sp.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True)
        if ph["kind"] == "spheres":
            kw = dict(shape=pp.SPHERE, fov_um=ph["fov_um"], resolution=ph["n"], Y=-1.0, radius_um=-20.0, volume_fraction=ph["vf"], seed=ph["seed"])
            r = pp.reference(omp=True, **kw) if pp.have_ref(omp=True) else pp.oracle(**kw)
            return r.mask, None, fov
        r = pp.oracle(zwin=(0, 1), shape=pp.CYLINDER, fov_um=ph["fov_um"], resolution=ph["n"], Y=ph["Y"], radius_um=ph["radius_um"],
                      volume_fraction=ph["bvf"], orientation_deg=90.0, seed=ph["seed"])
        return np.ascontiguousarray(r.mask[:, :, 0]), np.ascontiguousarray(r.fieldmap[:, :, 0]), fov
    except Exception:  # the numpy stand-in of the same recipe (statistically, not bitwise, the reference's shapes)
        return make_phantom_2d(ph)


class Walk:
    """The walk legs of the bench: one engine per workload, timed passes with the NCCL all-reduce of the sums inside."""

    def __init__(self, sw, torch, dist, dev, rank, world, local_rank):
        self.sw, self.torch, self.dist, self.dev, self.rank, self.world, self.local_rank = sw, torch, dist, dev, rank, world, local_rank

    def barrier(self):
        self.torch.cuda.synchronize(self.dev)
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def setup(self, name, cfg_kw, ph, spec=None):
        """engine with the workload's phantom (generated on this rank's device from the reference's own `spinwalk phantom` recipe — bit-identical
        to the reference's generator, tests/test_phantom_gpu.py — it never visits the host), sequence and this rank's shard of the spins."""
        from spinwalk_b200 import sharding

        sw, torch = self.sw, self.torch
        S = cfg_kw["n_spins"]
        cfg = sw.SimConfig(**dict(cfg_kw, n_spins=S * self.world))  # weak scaling: the global population grows with N
        eng = sw.Engine(self.local_rank)
        gen = eng.generate_phantom(spec or phantom_spec(ph))
        spin_first, n_local = sharding.shard_range(S * self.world, self.rank, self.world)
        assert n_local == S
        xyz0_pin = torch.empty((S, 3), dtype=torch.float32, pin_memory=True)
        xyz0_pin.numpy()[:] = make_positions(S, eng.fov, cfg.seed, spin_first)
        eng.set_sequence(cfg)
        eng.set_spins(xyz0_pin.numpy(), None, spin_first)
        K, E, ns = len(cfg.scales), cfg.n_TE, cfg.n_substrate
        slab = ph["kind"] != "spheres" and os.environ.get("SWK_NO_ZSLAB") is None
        return dict(name=name, eng=eng, cfg=cfg, S=S, K=K, E=E, ns=ns, fov=eng.fov, spin_first=spin_first, xyz0_pin=xyz0_pin, slab=slab, n=ph["n"],
                    sums_d=torch.zeros((K, E, ns, 4), dtype=torch.float64, device=self.dev),
                    steps_per_pass=S * K * (eng.n_dummy_scan + 1) * cfg.n_timepoints,
                    table=(f"z slab [nx][ny] of the z-invariant phantom: {4 * ph['n'] ** 2 / 1e6:.2f} MB (default, include/spinwalk_engine.h SWK_RUN_NO_ZSLAB)" if slab
                           else f"[nx][ny][nz] packed words: {4 * ph['n'] ** 3 / 1e9:.3f} GB" if gen.get("n_shapes") is not None and ph.get("Y", 0) is not None and ph["kind"] != "spheres" or (spec is not None)
                           else f"[nx][ny][nz] mask bytes: {ph['n'] ** 3 / 1e9:.3f} GB (no field map)"),
                    phantom_note=f"generated on the device by swk_generate_phantom: {gen['n_shapes']} shapes, volume fraction {gen['volume_fraction']:.3f} %, "
                                 f"voxel fill {gen['kernel_ms']:.2f} ms (bit-identical to the reference's `spinwalk phantom` for this recipe)")

    def one_pass(self, H, mode, flags, scales=None):
        from spinwalk_b200 import sharding

        sums = H["sums_d"] if scales is None else H["sums_d"][: len(scales)]
        st = H["eng"].run_device(scales=scales, mode=mode, flags=flags, d_sums_ptr=sums.data_ptr())
        sharding.allreduce_sums(sums)  # the one collective of the path: a few KB of per-echo ensemble sums (NCCL)
        return st

    def counts(self, H, mode, flags, scales=None):
        """work counters (voxel changes, rejections) of one untimed pass of the STATS kernel variant: the roofline's algorithmic bytes"""
        return self.one_pass(H, mode, flags | self.sw.RUN_STATS, scales)

    def timed(self, H, mode, flags, steps, warmup, scales=None):
        torch = self.torch
        for _ in range(warmup):
            self.one_pass(H, mode, flags, scales)
        self.barrier()
        dev_ms = ker_ms = 0.0
        t0 = time.perf_counter()
        for _ in range(steps):
            st = self.one_pass(H, mode, flags, scales)
            dev_ms += st["device_ms"]
            ker_ms += st["kernel_ms"]
        self.barrier()
        wall_s = time.perf_counter() - t0
        t = torch.tensor([dev_ms, ker_ms, wall_s * 1e3], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        dev_ms, ker_ms, wall_ms = (float(v) for v in t.tolist())
        return dict(dev_ms=dev_ms, ker_ms=ker_ms, wall_ms=wall_ms, st=st)

    def e2e(self, H, mode, per_spin_out, steps):
        """end to end through swk_run with host buffers: H2D of XYZ0 and D2H of M1 / XYZ1 / T / sums inside the timed region"""
        from spinwalk_b200 import sharding

        torch, eng = self.torch, H["eng"]
        K, S, E, ns = H["K"], H["S"], H["E"], H["ns"]
        if per_spin_out:
            out = (torch.empty((K, S, E, 3), dtype=torch.float32, pin_memory=True), torch.empty((K, S, eng.trj, 3), dtype=torch.float32, pin_memory=True),
                   torch.empty((K, S, E), dtype=torch.uint8, pin_memory=True))
            out_np = tuple(o.numpy() for o in out)
        else:
            out, out_np = (), None
        h2d = H["xyz0_pin"].numel() * 4 + K * 4
        d2h = sum(o.numel() * o.element_size() for o in out) + K * E * ns * 4 * 8
        x = H["xyz0_pin"].numpy()
        eng.run(x, None, H["spin_first"], mode=mode, out=out_np, outputs=per_spin_out, stats=False)  # warm
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            r = eng.run(x, None, H["spin_first"], mode=mode, out=out_np, outputs=per_spin_out, stats=False)
            if self.world > 1:
                H["sums_d"].copy_(torch.from_numpy(r["sums"]))
                sharding.allreduce_sums(H["sums_d"])
        self.barrier()
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(te, op=self.dist.ReduceOp.MAX)
        del out, out_np
        return {"value": H["steps_per_pass"] * self.world * steps / float(te.item()), "unit": "spin-steps/s", "h2d_bytes_per_step": h2d * self.world,
                "d2h_bytes_per_step": d2h * self.world, "ms_per_step": 1e3 * float(te.item()) / steps,
                "api": "swk_run (C-ABI) with pinned host buffers: XYZ0 in; " + ("M1, XYZ1, T, sums out" if per_spin_out else "sums out")}

    def roofline(self, H, counts, T, steps, per_spin_out, peak, peak_src, mode, args, steps_scale=1.0):
        sw = self.sw
        K, E, S = H["K"], H["E"], H["S"]
        per_pass_bytes = (counts["mask_gathers"] * (4 if H["eng"].has_fieldmap else 1)) * steps_scale + S * K * (24 + ((13 * E + 12) if per_spin_out else 0))
        ker_ms = T["ker_ms"] / steps
        achieved = per_pass_bytes / (ker_ms * 1e-3) / 1e9
        traffic, traffic_src, ncu_launches = stamped_traffic(f"{H['name']}:{args.mode}{'' if H['slab'] else ':full'}", S)
        return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                "ncu_launches": ncu_launches,
                "peak_source": peak_src, "kernel": "swk::walk_fast_kernel" if mode == sw.MODE_FAST else "swk::walk_compat_kernel",
                "algorithmic_bytes_per_launch": per_pass_bytes, "kernel_ms_per_launch": ker_ms, "bytes_per_spin_step": per_pass_bytes / H["steps_per_pass"],
                "attempts_with_voxel_change_per_step": counts["mask_gathers"] * steps_scale / max(1, H["steps_per_pass"]),
                "rejects_per_step": counts["rejects"] * steps_scale / max(1, H["steps_per_pass"]),
                "note": "algorithmic bytes (SURVEY §8d): 4 B packed voxel word per attempt whose voxel changed + 24 B in + (13 E + 12) B out per (spin, scale).  "
                        + ("With the z-slab table the words come from L1/L2: the launch is ISSUE bound (ncu: issue slots, profiles/README.md), so this HBM fraction is small by "
                           "construction; the gather roofline is measured by `full_table`." if H["slab"] else
                           "A 4-byte gather costs a 64-byte HBM access, so the byte fraction is small by construction; `gather` compares the walk with the random-gather probe.")}

    def cpu_baseline(self, H, cfg_kw, ph, target_s=15.0):
        try:
            mask, fm = H["eng"].get_phantom()  # the baseline walks the very same voxels
            H["host_phantom"] = (mask, fm)
            cb, _, _ = cpu_reference(cfg_kw, ph, mask, fm, H["fov"], target_s=target_s)
            return cb
        except Exception as ex:  # the baseline is a reported extra; never lose the GPU line over it
            return {"value": None, "unit": "spin-steps/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}

    def close(self, H):
        H["eng"].close()
        for k in ("xyz0_pin", "sums_d"):
            H.pop(k, None)

    # ------------------------------------------------------------------ sub-records
    def gather_roofline(self, H, mode, n_small=10):
        """the walk against the random-gather probe on the SAME table, on the smallest FoV scales only — there every attempt lands in a voxel far
        from the last one (sigma >= 4 voxels), so walk gathers and probe gathers are the same kind of access (L2 hits included on both sides)"""
        small = list(H["cfg"].scales[:n_small])
        c = self.counts(H, mode, 0, small)
        t = self.timed(H, mode, 0, 2, 1, small)
        probe = H["eng"].probe_gather(threads_per_sm=2048, iters=2048)
        walk = c["mask_gathers"] / (t["ker_ms"] / 2 * 1e-3)
        return {"scales": [small[0], small[-1]], "walk_gathers_per_s": walk, "probe_gathers_per_s": probe["gathers_per_s"], "table_bytes": probe["table_bytes"],
                "frac": walk / probe["gathers_per_s"], "kernel_ms": t["ker_ms"] / 2,
                "note": "the walk against the random-gather probe on the SAME voxel table, on the 10 smallest FoV scales only: there every attempt lands in a voxel far "
                        "from the last one (sigma >= 4 voxels), so both count the same kind of access.  probe = swk_probe_gather: dependent random 4-byte gathers with "
                        "the walk's load instruction and nothing else — a cache-resident table (z slab) is bound by the L1TEX tag rate (one 128-byte line per clock and SM "
                        "for a fully divergent warp load: 291 G/s), a table in HBM by row activations (DESIGN.md §5); walk = attempts whose voxel changed (STATS kernel "
                        "variant) / kernel time of the same scales"}

    def extra_full_table(self, peak, peak_src, args):
        """the headline workload on the FULL [nx][ny][nz] voxel table (SWK_RUN_NO_ZSLAB): what a phantom without an invariant axis costs, and the
        configuration the gather roofline applies to"""
        sw = self.sw
        cfg_kw, ph, desc = workload("c2", None, None)
        H = self.setup("c2", cfg_kw, ph)
        H["slab"] = False
        fl = sw.OUT_ALL | sw.RUN_NO_ZSLAB
        counts = self.counts(H, sw.MODE_FAST, fl)
        T = self.timed(H, sw.MODE_FAST, fl, 2, 1)
        rec = {"value": H["steps_per_pass"] * self.world * 2 / (T["dev_ms"] * 1e-3), "unit": "spin-steps/s", "ms_per_step": T["dev_ms"] / 2, "steps": 2, "warmup": 1,
               "config": {"workload": desc + "; full voxel table (864 MB, HBM resident)", "spins_per_gpu": H["S"]},
               "roofline": self.roofline(H, counts, T, 2, True, peak, peak_src, sw.MODE_FAST, args)}
        fl0 = sw.RUN_NO_ZSLAB
        H2 = dict(H)
        small = list(H["cfg"].scales[:10])
        c = self.one_pass(H2, sw.MODE_FAST, fl0 | sw.RUN_STATS, small)
        t = self.timed(H2, sw.MODE_FAST, fl0, 2, 1, small)
        probe = H["eng"].probe_gather(threads_per_sm=2048, iters=2048)
        walk = c["mask_gathers"] / (t["ker_ms"] / 2 * 1e-3)
        rec["roofline"]["gather"] = {"scales": [small[0], small[-1]], "walk_gathers_per_s": walk, "probe_gathers_per_s": probe["gathers_per_s"],
                                     "table_bytes": probe["table_bytes"], "frac": walk / probe["gathers_per_s"], "kernel_ms": t["ker_ms"] / 2,
                                     "note": "the walk against the random-gather probe on the SAME table, on the 10 smallest FoV scales only: there every attempt lands in a "
                                             "voxel far from the last one (sigma >= 4 voxels), so both count the same kind of access (L2 hits included on both sides).  "
                                             "probe = swk_probe_gather: dependent random 4-byte gathers with the walk's load instruction and nothing else (HBM row-activation "
                                             "bound, DESIGN.md §5); walk = attempts whose voxel changed (STATS kernel variant) / kernel time"}
        self.close(H)
        return rec

    def extra_compat(self, peak, peak_src, args):
        """SWK_MODE_COMPAT — the reference's own arithmetic (minstd + erfcinvf, FP64 positions), bit-exact against its cu_sim — on a 2e6-spin sample"""
        sw = self.sw
        cfg_kw, ph, desc = workload("c2", 2_000_000, None)
        H = self.setup("c2", cfg_kw, ph)
        T = self.timed(H, sw.MODE_COMPAT, sw.OUT_ALL, 2, 1)
        rec = {"value": H["steps_per_pass"] * self.world * 2 / (T["dev_ms"] * 1e-3), "unit": "spin-steps/s", "ms_per_step": T["dev_ms"] / 2, "steps": 2, "warmup": 1,
               "dtype": "f32+f64", "config": {"workload": desc.replace("1e7 spins", "2e6 of the 1e7 spins") + "; SWK_MODE_COMPAT (bit-exact T / XYZ1 vs the reference's cu_sim, "
                                              "tests/test_engine_gpu.py); z-invariant phantom: raw (substrate id, FP32 field) pairs of one z plane, the values of the full arrays", "spins_per_gpu": H["S"]}}
        self.close(H)
        return rec

    def extra_scale_groups(self, peak, peak_src, args):
        """kernel time of the headline workload per group of 10 consecutive FoV scales (1e7 spins each): where the pass spends its time"""
        sw = self.sw
        cfg_kw, ph, desc = workload("c2", None, None)
        H = self.setup("c2", cfg_kw, ph)
        out = []
        sc = list(H["cfg"].scales)
        for g in range(0, len(sc), 10):
            part = sc[g:g + 10]
            c = self.counts(H, sw.MODE_FAST, 0, part)
            t = self.timed(H, sw.MODE_FAST, 0, 1, 1, part)
            n = H["S"] * len(part) * H["cfg"].n_timepoints
            out.append({"scales": [part[0], part[-1]], "kernel_ms": t["ker_ms"], "spin_steps_per_s": n * self.world / (t["dev_ms"] * 1e-3),
                        "attempts_per_step": 1.0 + c["rejects"] / max(1, c["steps"]), "voxel_changes_per_step": c["mask_gathers"] / max(1, c["steps"])})
        self.close(H)
        return out

    def extra_non_invariant(self, peak, peak_src, args):
        """a phantom WITHOUT an invariant axis: SE BOLD on 400^3 random spheres with their dipole field map (`phantom -s -r -20 -v 30 -y 0.78`), full packed table (256 MB > L2)"""
        sw = self.sw
        from spinwalk_b200 import phantom_gen as pg

        cfg_kw, _, _ = workload("c2", 2_000_000, None)
        ph = dict(kind="spheres", n=400, fov_um=400.0, vf=30.0, seed=0, Y=0.78)
        spec = pg.PhantomSpec(shape=pg.SHAPE_SPHERE, fov_um=400.0, resolution=400, oxy_level=0.78, radius_um=-20.0, volume_fraction=30.0, seed=0)
        H = self.setup("c2s", cfg_kw, ph, spec=spec)
        H["slab"] = False
        counts = self.counts(H, sw.MODE_FAST, sw.OUT_ALL)
        T = self.timed(H, sw.MODE_FAST, sw.OUT_ALL, 2, 1)
        rec = {"value": H["steps_per_pass"] * self.world * 2 / (T["dev_ms"] * 1e-3), "unit": "spin-steps/s", "ms_per_step": T["dev_ms"] / 2, "steps": 2, "warmup": 1,
               "config": {"workload": "SE BOLD (config/se.ini) on `phantom -s -r -20 -v 30 -y 0.78 -f 400 -z 400 -e 0`: 400^3 random spheres + dipole field map (no invariant axis), "
                                      "2e6 spins x 50 FoV scales", "spins_per_gpu": H["S"], "phantom": H["phantom_note"]},
               "roofline": self.roofline(H, counts, T, 2, True, peak, peak_src, sw.MODE_FAST, args), "lost_spins": counts["lost"]}
        rec["roofline"]["gather"] = self.gather_roofline(H, sw.MODE_FAST)
        self.close(H)
        return rec

    def extra_gradient_scales(self, peak, peak_src, args):
        """BASELINE.json configs[2] (C3): PGSE on the 400^3 permeable-sphere phantom, 1e7 spins x 51 GRADIENT scales (b = 100..5000, 0).  Gradient scales do
        not change the walk and the reference replays one random stream per spin for every scale (kernels.cu:77-88): one walker per spin carries the
        magnetisation of all 51 scales (walk_fast.cuh MULTI; SWK_RUN_NO_ONEWALK walks every scale separately — the `per_scale_walks` leg, 1e6 spins)"""
        sw = self.sw
        out = {}
        for name in ("c3", "c3r"):
            cfg_kw, ph, desc = workload(name, None, None)
            H = self.setup(name, cfg_kw, ph)
            T = self.timed(H, sw.MODE_FAST, sw.OUT_ALL, 2, 1)
            rec = {"value": H["steps_per_pass"] * self.world * 2 / (T["dev_ms"] * 1e-3), "unit": "spin-steps/s", "ms_per_step": T["dev_ms"] / 2, "kernel_ms_per_step": T["ker_ms"] / 2,
                   "steps": 2, "warmup": 1, "config": {"workload": desc, "spins_per_gpu": H["S"], "n_scales": H["K"], "timepoints": H["cfg"].n_timepoints,
                                                       "note": "spin-steps = spins x scales x timepoints, the work the reference does; one walk per spin serves all scales"}}
            S1 = 1_000_000
            H["eng"].set_spins(H["xyz0_pin"].numpy()[:S1], None, H["spin_first"])
            H1 = dict(H, S=S1, steps_per_pass=S1 * H["K"] * H["cfg"].n_timepoints)
            T1 = self.timed(H1, sw.MODE_FAST, sw.OUT_ALL | sw.RUN_NO_ONEWALK, 1, 1)
            rec["per_scale_walks"] = {"value": H1["steps_per_pass"] * self.world / (T1["dev_ms"] * 1e-3), "unit": "spin-steps/s", "ms_per_step": T1["dev_ms"], "spins_per_gpu": S1,
                                      "note": "SWK_RUN_NO_ONEWALK: one walker per (spin, scale), as round 1 ran it"}
            self.close(H)
            self.torch.cuda.empty_cache()
            out[name] = rec
        return out

    def extra_other_configs(self, peak, peak_src, args):
        """the remaining BASELINE.json configurations, so that one line carries all five: C1 (GRE BOLD, 100^3, 1e5 spins x 50 FoV scales) at full size and
        C4 (bSSFP, 1101 TRs x 200 steps, one scale) on 2e6 of its 1e7 spins"""
        sw = self.sw
        out = {}
        for name, spins, steps in (("c1", None, 5), ("c4", 2_000_000, 1)):
            cfg_kw, ph, desc = workload(name, spins, None)
            H = self.setup(name, cfg_kw, ph)
            T = self.timed(H, sw.MODE_FAST, sw.OUT_ALL, steps, 1)
            out[name] = {"value": H["steps_per_pass"] * self.world * steps / (T["dev_ms"] * 1e-3), "unit": "spin-steps/s", "ms_per_step": T["dev_ms"] / steps, "steps": steps, "warmup": 1,
                         "config": {"workload": desc if spins is None else desc.replace("1e7 spins", f"{spins:.0e} of the 1e7 spins".replace("e+0", "e")), "spins_per_gpu": H["S"],
                                    "n_scales": H["K"], "timepoints": H["cfg"].n_timepoints, "scans": H["eng"].n_dummy_scan + 1}}
            self.close(H)
            self.torch.cuda.empty_cache()
        return out

    def extra_north_star(self, peak, peak_src, args):
        """BASELINE.json configs[4] / north_star: 1000^3 BOLD phantom, 1.25e8 spins per GPU (1e9 over 8), 50 FoV scales, ensemble sums only, NCCL all-reduce"""
        sw = self.sw
        cfg_kw, ph, desc = workload("c5", None, None)
        H = self.setup("c5", cfg_kw, ph)
        T = self.timed(H, sw.MODE_FAST, 0, 2, 1)
        rec = {"value": H["steps_per_pass"] * self.world * 2 / (T["dev_ms"] * 1e-3), "unit": "spin-steps/s", "n_gpus": self.world, "ms_per_step": T["dev_ms"] / 2, "steps": 2,
               "warmup": 1, "scaling": "weak", "config": {"workload": desc, "spins_per_gpu": H["S"], "spins_total": H["S"] * self.world, "voxel_table": H["table"],
                                                            "phantom": H["phantom_note"]},
               "gpu_launches": 2 * T["st"]["n_launches"]}
        # the same phantom on its full 4 GB voxel table, 2.5e7 of the spins: the gather-roofline configuration of the north star
        S_small = 25_000_000
        H["eng"].set_spins(H["xyz0_pin"].numpy()[:S_small], None, H["spin_first"])
        Hs = dict(H, S=S_small, slab=False, steps_per_pass=S_small * H["K"] * H["cfg"].n_timepoints)
        fl = sw.RUN_NO_ZSLAB
        counts = self.counts(Hs, sw.MODE_FAST, fl)
        Tf = self.timed(Hs, sw.MODE_FAST, fl, 1, 1)
        ft = {"value": Hs["steps_per_pass"] * self.world / (Tf["dev_ms"] * 1e-3), "unit": "spin-steps/s", "ms_per_step": Tf["dev_ms"], "steps": 1, "warmup": 1,
              "config": {"workload": "the same phantom on its full [nx][ny][nz] table (4 GB, SWK_RUN_NO_ZSLAB), 2.5e7 spins per GPU x 50 FoV scales, sums only", "spins_per_gpu": S_small},
              "roofline": self.roofline(Hs, counts, Tf, 1, False, peak, peak_src, sw.MODE_FAST, args)}
        small = list(H["cfg"].scales[:10])
        c = self.one_pass(Hs, sw.MODE_FAST, fl | sw.RUN_STATS, small)
        t = self.timed(Hs, sw.MODE_FAST, fl, 1, 1, small)
        probe = H["eng"].probe_gather(threads_per_sm=2048, iters=2048)
        walk = c["mask_gathers"] / (t["ker_ms"] * 1e-3)
        ft["roofline"]["gather"] = {"scales": [small[0], small[-1]], "walk_gathers_per_s": walk, "probe_gathers_per_s": probe["gathers_per_s"], "table_bytes": probe["table_bytes"],
                                    "frac": walk / probe["gathers_per_s"], "kernel_ms": t["ker_ms"],
                                    "note": "walk vs random-gather probe on the same 4 GB table, 10 smallest FoV scales (see full_table.roofline.gather.note)"}
        rec["full_table"] = ft
        if self.rank == 0 and self.world == 1 and not args.no_cpu_baseline:
            rec["cpu_baseline"] = self.cpu_baseline(H, cfg_kw, ph, target_s=10.0)
            H.pop("host_phantom", None)
        self.close(H)
        return rec


def source_stamp():
    """identity of the kernels a capture belongs to: SHA-256 over the CUDA sources (they travel with every snapshot)"""
    import hashlib

    h = hashlib.sha256()
    d = os.path.join(ROOT, "spinwalk_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.startswith(("walk_", "engine")):
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def stamped_traffic(key, spins_per_gpu):
    """DRAM bytes per launch from the committed ncu capture (profiles/traffic.json, written by scripts/make_traffic.py) — only when the capture was
    taken on THESE kernel sources; a stale capture is refused, not silently reused."""
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        return None, "no capture (profiles/traffic.json missing)", None
    ent = tj.get(key)
    if not ent:
        return None, f"no capture for {key}", None
    if ent.get("kernel_stamp") != source_stamp():
        return None, f"stale capture refused: taken on kernel sources {ent.get('kernel_stamp')}, this build is {source_stamp()}", None
    # per walk launch of the pass: issue-slot utilisation, active lanes per instruction, cache hit rates (what bounds a cache-resident table)
    return ent["dram_bytes_per_launch"] * (spins_per_gpu / ent["spins_per_gpu"]), ent["source"], ent.get("launches")


if __name__ == "__main__":
    main()
