"""Engine (COMPAT) against the reference's own cu_sim on the seeded RANDOM cases of tests/random_cases.py (the oracle is pinned on them on
the CPU, tests/test_oracle_random.py).  Written when no GPU time was left to run it: run it first
(python scripts/random_parity_gpu.py [n_cases]) and, once it is green, promote it to tests/ as a -m gpu test.
Bars as in tests/test_engine_gpu.py: T and XYZ1 bitwise, M1 <= 2e-6; FAST mode must run every case to finite outputs (its lost-spin count is printed beside COMPAT's)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import random_cases  # noqa: E402
import spinwalk_b200 as sw  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
bad = 0
for seed in range(n):
    case, mask, fm, fov, xyz0 = random_cases.make(seed)
    ref = po.run_ref_cuda(case, fm, mask, xyz0)
    cfg = cases.to_simconfig(case)
    with sw.Engine(0) as e:
        e.set_phantom(mask, fm, fov)
        e.set_sequence(cfg)
        got = e.run(xyz0, mode=sw.MODE_COMPAT)
        fast = e.run(xyz0, mode=sw.MODE_FAST)
    okT = np.array_equal(got["T"], ref["T"])
    okX = np.array_equal(got["XYZ1"].view(np.uint32), ref["XYZ1"].view(np.uint32))
    dM = float(np.abs(got["M1"] - ref["M1"]).max()) if got["M1"].size else 0.0
    okF = bool(np.isfinite(fast["M1"]).all() and np.isfinite(fast["XYZ1"]).all())
    if not (okT and okX and dM <= 2e-6 and okF):
        bad += 1
        print(f"seed {seed}: T {okT} XYZ1 {okX} max|dM1| {dM:.3g} fast ok {okF} (lost compat {got['stats']['lost']} fast {fast['stats']['lost']})", flush=True)
print(f"{n - bad} / {n} random cases agree with the reference cu_sim")
sys.exit(1 if bad else 0)
