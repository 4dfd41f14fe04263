"""Golden vectors for the phantom generators, produced by the UNMODIFIED reference classes (oracle/_ref/libswref_gen.so, serial
build; g++ 13.3, baseline x86-64, -O2).  Run in the build container (needs /root/reference): python tests/golden/make_phantom_golden.py

A golden holds the placed shapes, the actual volume fraction, SHA-256 digests of the full mask / field-map byte strings (the work is
bit-exact, so a digest pins every voxel) and one z slice in the clear for diagnostics."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import pyphantom as pp  # noqa: E402
from phantom_cases import CASES  # noqa: E402


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    out = os.path.join(ROOT, "tests", "golden", "phantom")
    os.makedirs(out, exist_ok=True)
    for name, kw in CASES.items():
        ph = pp.reference(**kw)
        n = kw["resolution"]
        d = dict(shapes=ph.shapes, bvf=np.float32(ph.bvf), mask_sha256=digest(ph.mask), mask_slice=ph.mask[:, :, n // 2].copy())
        if ph.fieldmap is not None:
            d.update(fieldmap_sha256=digest(ph.fieldmap), fieldmap_slice=ph.fieldmap[:, :, n // 2].copy())
        np.savez_compressed(os.path.join(out, name + ".npz"), **d)
        print(name, len(ph.shapes), ph.bvf, d["mask_sha256"][:12])


if __name__ == "__main__":
    main()
