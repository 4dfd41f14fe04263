"""Per-scale throughput of the walk kernel (diagnostic; not the bench).  python scripts/scale_sweep.py [--n 600] [--spins 1000000]"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import spinwalk_b200 as sw  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=600)
ap.add_argument("--spins", type=int, default=1_000_000)
ap.add_argument("--scales", type=str, default="0.0125,0.0641,0.2010,0.6309,1.0301,3.2330,10.147,37.5")
ap.add_argument("--modes", type=str, default="fast,compat")
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--flags", type=int, default=sw.OUT_ALL)
ap.add_argument("--nosort", action="store_true")
ap.add_argument("--dup", type=int, default=1, help="run every scale DUP times in one launch (DUP >= 2: the SHARED kernel variant, DUP walkers share a spin's normals)")
ap.add_argument("--nopack", action="store_true")
args = ap.parse_args()
if args.nosort:
    args.flags |= sw.RUN_NO_SORT
if args.nopack:
    args.flags |= sw.RUN_NO_PACK

cfg_kw, ph, _ = bench.workload("c2", args.spins, None)
ph["n"], ph["fov_um"] = args.n, float(args.n)
cfg = sw.SimConfig(**cfg_kw)
mask2, fm2, fov = bench.make_phantom_2d(ph)
dev = torch.device("cuda", 0)
n = args.n
mask_d = torch.from_numpy(mask2).to(dev)[:, :, None].expand(n, n, n).contiguous()
fm_d = torch.from_numpy(fm2).to(dev)[:, :, None].expand(n, n, n).contiguous()
eng = sw.Engine(0)
eng.set_phantom(mask_d, fm_d, fov)
eng.set_sequence(cfg)
eng.set_spins(bench.make_positions(args.spins, fov, 10))
for mode_name in args.modes.split(","):
    mode = sw.MODE_FAST if mode_name == "fast" else sw.MODE_COMPAT
    for s in [float(x) for x in args.scales.split(",")]:
        st = eng.run_device(scales=[s] * args.dup, mode=mode, flags=args.flags | sw.RUN_STATS)
        best = 1e30
        for _ in range(args.reps):
            best = min(best, eng.run_device(scales=[s] * args.dup, mode=mode, flags=args.flags)["kernel_ms"])
        steps = args.spins * 800 * args.dup
        print(f"{mode_name:6s} scale {s:8.4f}: {steps / best / 1e6:8.2f} Gsteps/s  kernel {best:8.2f} ms  p_chg {st['mask_gathers'] / st['steps']:.3f} "
              f"rej/step {st['rejects'] / st['steps']:.4f} lost {st['lost']}", flush=True)
