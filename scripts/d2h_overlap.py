"""Diagnostic: device-to-host copy bandwidth alone vs. while the walk kernel saturates HBM with random gathers.
python scripts/d2h_overlap.py"""
import os
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import spinwalk_b200 as sw  # noqa: E402

dev = torch.device("cuda", 0)
GB = 4
src = torch.empty(GB << 30, dtype=torch.uint8, device=dev)
dst = torch.empty(GB << 30, dtype=torch.uint8, pin_memory=True)
side = torch.cuda.Stream()


def d2h():
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.cuda.stream(side):
        dst.copy_(src, non_blocking=True)
    side.synchronize()
    return time.perf_counter() - t0


d2h()
print(f"D2H alone: {GB / d2h():.1f} GiB/s", flush=True)

cfg_kw, ph, _ = bench.workload("c2", 4_000_000, None)
cfg = sw.SimConfig(**cfg_kw)
mask2, fm2, fov = bench.make_phantom_2d(ph)
n = ph["n"]
mask_d = torch.from_numpy(mask2).to(dev)[:, :, None].expand(n, n, n).contiguous()
fm_d = torch.from_numpy(fm2).to(dev)[:, :, None].expand(n, n, n).contiguous()
eng = sw.Engine(0)
eng.set_phantom(mask_d, fm_d, fov)
eng.set_sequence(cfg)
eng.set_spins(bench.make_positions(4_000_000, fov, 10))
st = eng.run_device(flags=0)
print(f"kernel alone: {st['kernel_ms']:.0f} ms", flush=True)
res = {}
th = threading.Thread(target=lambda: res.update(st=eng.run_device(flags=0)))
th.start()
time.sleep(0.15)
t = d2h()
th.join()
print(f"D2H during kernel: {GB / t:.1f} GiB/s ({t * 1e3:.0f} ms); kernel with concurrent D2H: {res['st']['kernel_ms']:.0f} ms", flush=True)
for s in (0.0125, 1.03, 37.5):
    th = threading.Thread(target=lambda: [eng.run_device(scales=[s] * 10, flags=0) for _ in range(3)])
    th.start()
    time.sleep(0.1)
    t = d2h()
    th.join()
    print(f"D2H during scale-{s} kernel: {GB / t:.1f} GiB/s", flush=True)
