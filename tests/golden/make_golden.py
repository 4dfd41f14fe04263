"""Generates tests/golden/<case>_<flavour>.npz from the UNMODIFIED reference (oracle/_ref, built by
`make -C oracle ref` from /root/reference/src/sim/kernels.cu).  Run in the container that has /root/reference:
    python tests/golden/make_golden.py [case ...]      (no names: every case of tests/cases.py)
Each file records the toolchain, the seeded inputs (xyz0) and the reference outputs M1 / XYZ1 / T."""
import os
import platform
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import cases  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

po.build(ref=True)
tool = "g++ " + subprocess.run(["g++", "-dumpfullversion"], capture_output=True, text=True).stdout.strip() + \
       "; nvcc " + subprocess.run(["nvcc", "--version"], capture_output=True, text=True).stdout.strip().split("release ")[-1].split(",")[0] + \
       "; " + platform.platform()
for name, fn in cases.ALL.items():
    if len(sys.argv) > 1 and name not in sys.argv[1:]:
        continue
    case, mask, fm, fov, xyz0 = fn()
    for flavour, tag in ((po.RNG_MT19937, "mt19937"), (po.RNG_MINSTD, "minstd")):
        r = po.run_ref(case, fm, mask, xyz0, flavour=flavour)
        out = os.path.join(HERE, f"{name}_{tag}.npz")
        np.savez_compressed(out, case=name, flavour=flavour, toolchain=tool, reference="aghaeifar/SpinWalk v1.21.0 src/sim/kernels.cu",
                            xyz0=xyz0, M1=r["M1"], XYZ1=r["XYZ1"], T=r["T"])
        print(out, os.path.getsize(out))
