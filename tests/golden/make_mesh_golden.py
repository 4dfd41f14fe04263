"""Golden masks of the reference's mesh phantom (`spinwalk phantom -p`): phantom::ply::run(false) of the UNMODIFIED src/phantom/phantom_ply.cpp
(+ vendored happly) on the PLY files tests/meshes.py writes.  Run in the build container: python tests/golden/make_mesh_golden.py"""
import hashlib
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import meshes  # noqa: E402
from oracle import pyphantom as pp  # noqa: E402

SIZES = ((60.0, 24), (100.0, 37))


def main():
    out = os.path.join(ROOT, "tests", "golden", "phantom")
    d = tempfile.mkdtemp()
    for name, make in meshes.MESHES.items():
        v, f = make()
        path = os.path.join(d, name + ".ply")
        meshes.write_ply(path, v, f, fmt="binary_little_endian", vertex_type="double")
        rec = {}
        for fov, res in SIZES:
            mask, bvf = pp.reference_mesh(fov, res, path)
            rec[f"sha256_{res}"] = hashlib.sha256(mask.tobytes()).hexdigest()
            rec[f"slice_{res}"] = mask[:, :, res // 2].copy()
            rec[f"inside_{res}"] = np.int64(mask.sum())
            assert bvf == 0.0  # the reference never fills m_volume_fraction for mesh phantoms (phantom_ply.cpp:187)
        np.savez_compressed(os.path.join(out, "mesh_" + name + ".npz"), **rec)
        print(name, len(f), {k: (v if not isinstance(v, np.ndarray) else v.shape) for k, v in rec.items() if not k.startswith("slice")})


if __name__ == "__main__":
    main()
