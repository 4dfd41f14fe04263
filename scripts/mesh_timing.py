"""Where does the wall time of swk_phantom_mesh go?  python scripts/mesh_timing.py"""
import sys
import time

sys.path.insert(0, ".")
import numpy as np  # noqa: E402

from spinwalk_b200 import phantom_gen as pg  # noqa: E402
from spinwalk_b200.phantoms import icosphere_mesh  # noqa: E402

t0 = time.perf_counter()
v, f = icosphere_mesh(6, 0.2)
print("mesh build %.3f s" % (time.perf_counter() - t0), v.shape, f.shape)
for n in (512, 512, 512, 256):
    t0 = time.perf_counter()
    mask, _, st = pg.generate_mesh(512.0, n, v, f)
    print(n, "wall %.1f ms" % (1e3 * (time.perf_counter() - t0)), {k: round(x, 2) if isinstance(x, float) else x for k, x in st.items()}, flush=True)
