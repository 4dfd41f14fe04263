"""Kernel time of the C2 workload (all 50 FoV scales, reduced spin count): full voxel table vs z slab (the default), SHARED vs PRIVATE kernel variant.
Diagnostic, not the bench.  python scripts/zslab_probe.py [spins] [c2|c5|c1]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import spinwalk_b200 as sw  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
cfg_kw, ph, _ = bench.workload(sys.argv[2] if len(sys.argv) > 2 else "c2", S, None)
cfg = sw.SimConfig(**cfg_kw)
eng = sw.Engine(0)
eng.generate_phantom(bench.phantom_spec(ph))
eng.set_sequence(cfg)
eng.set_spins(bench.make_positions(S, eng.fov, cfg.seed))
steps = S * len(cfg_kw["scales"]) * cfg.n_timepoints
for name, fl in (("full-table", sw.RUN_NO_ZSLAB), ("full-private", sw.RUN_NO_ZSLAB | sw.RUN_NO_SHARE), ("zslab", 0), ("zslab-private", sw.RUN_NO_SHARE)):
    eng.run_device(mode=sw.MODE_FAST, flags=sw.OUT_ALL | fl)
    ms = min(eng.run_device(mode=sw.MODE_FAST, flags=sw.OUT_ALL | fl)["kernel_ms"] for _ in range(2))
    print(f"{name:8s} kernel {ms:8.2f} ms  {steps / ms / 1e6:8.2f} Gsteps/s", flush=True)
