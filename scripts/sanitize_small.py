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
    print(name, "ok", flush=True)
