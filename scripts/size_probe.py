"""Kernel time of the 10 smallest FoV scales of C2 against the number of spins (PRIVATE variant), with and without the locality order.  Diagnostic."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import spinwalk_b200 as sw  # noqa: E402

cfg_kw, ph, _ = bench.workload("c2", 10_000_000, None)
cfg = sw.SimConfig(**cfg_kw)
eng = sw.Engine(0)
eng.generate_phantom(bench.phantom_spec(ph))
eng.set_sequence(cfg)
xyz = bench.make_positions(10_000_000, eng.fov, cfg.seed)
part = list(cfg.scales[:10])
for S in (1_000_000, 2_000_000, 5_000_000, 10_000_000):
    eng.set_spins(xyz[:S])
    for name, fl in (("sorted", 0), ("unsorted", sw.RUN_NO_SORT)):
        eng.run_device(scales=part, mode=sw.MODE_FAST, flags=fl)
        ms = min(eng.run_device(scales=part, mode=sw.MODE_FAST, flags=fl)["kernel_ms"] for _ in range(2))
        one = min(eng.run_device(scales=part[:1], mode=sw.MODE_FAST, flags=fl)["kernel_ms"] for _ in range(2))
        print(f"S {S:9d} {name:9s} 10 smallest scales {ms:8.2f} ms = {ms / S * 1e6:7.2f} ms per 1e6 spins; scale 0.0125 alone {one:8.2f} ms = {one / S * 1e6:7.2f} per 1e6", flush=True)
