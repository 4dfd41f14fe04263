"""Small FAST + COMPAT runs for `compute-sanitizer --tool memcheck|racecheck|initcheck python scripts/sanitize_small.py`
(the reference has no sanitizer coverage of the kernel)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import cases  # noqa: E402
import spinwalk_b200 as sw  # noqa: E402

os.environ["SWK_REBIN_SCANS"] = "4"
for name in ("multi_echo", "pgse", "ssfp", "stuck"):
    case, mask, fm, fov, xyz0 = cases.ALL[name]()
    cfg = cases.to_simconfig(case)
    with sw.Engine(0) as e:
        e.set_phantom(mask, fm, fov)
        e.set_sequence(cfg)
        for mode in (sw.MODE_FAST, sw.MODE_COMPAT):
            r = e.run(xyz0[:200], mode=mode)
            assert np.isfinite(r["M1"]).all()
        # the kernel variants the default choice does not take for this case: one walker per (spin, scale) with and without shared normals
        # (pgse / ssfp default to ONE walk for all scales — the MULTI kernels — above), the full voxel table, trajectories of the SHARED variant
        e.set_spins(xyz0[:200])
        for fl in (sw.RUN_NO_ONEWALK, sw.RUN_NO_ONEWALK | sw.RUN_NO_SHARE, sw.RUN_NO_ZSLAB, sw.RUN_NO_PACK | sw.RUN_STATS):
            e.run_device(mode=sw.MODE_FAST, flags=sw.OUT_ALL | fl)
            assert np.isfinite(e.download()[0]).all()
        e.run_device(mode=sw.MODE_COMPAT, flags=sw.OUT_ALL | sw.RUN_NO_ZSLAB)
    print(name, "ok", flush=True)
case, mask, fm, fov, xyz0 = cases.gre(n_spins=300, scales=tuple(0.05 * 1.6 ** i for i in range(12)))  # SHARED + PRIVATE launches side by side, sliced host run
os.environ["SWK_SLICES"] = "3"
with sw.Engine(0) as e:
    e.set_phantom(mask, fm, fov)
    e.set_sequence(cases.to_simconfig(case))
    assert np.isfinite(e.run(xyz0, mode=sw.MODE_FAST)["M1"]).all()
print("gre x 12 scales ok", flush=True)
