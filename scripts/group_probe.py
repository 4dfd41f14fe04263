"""Kernel time of the C2 workload per group of 10 consecutive FoV scales and as a whole, for kernel-variant choices given on the command line.
Diagnostic, not the bench.  python scripts/group_probe.py SPINS WORKLOAD name[:ENV=VAL,...] ...   (flags via env: SWK_NO_SHARE, SWK_GROUP, SWK_SHARE_SIGMA; PROBE_MODE=compat, PROBE_FLAGS)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import spinwalk_b200 as sw  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
wl = sys.argv[2] if len(sys.argv) > 2 else "c2"
variants = sys.argv[3:] or ["default", "private:SWK_NO_SHARE=1", "shared-all:SWK_SHARE_SIGMA=1e9"]
cfg_kw, ph, _ = bench.workload(wl, S, None)
cfg = sw.SimConfig(**cfg_kw)
eng = sw.Engine(0)
eng.generate_phantom(bench.phantom_spec(ph))
eng.set_sequence(cfg)
eng.set_spins(bench.make_positions(S, eng.fov, cfg.seed))
sc = list(cfg.scales)
fl = int(os.environ.get("PROBE_FLAGS", "0"))
MODE = sw.MODE_COMPAT if os.environ.get("PROBE_MODE") == "compat" else sw.MODE_FAST
for v in variants:
    name, _, envs = v.partition(":")
    keys = []
    for kv in filter(None, envs.split(",")):
        k, _, val = kv.partition("=")
        os.environ[k] = val
        keys.append(k)
    row = []
    for g in range(0, len(sc), 10):
        part = sc[g:g + 10]
        eng.run_device(scales=part, mode=MODE, flags=fl)
        row.append(min(eng.run_device(scales=part, mode=MODE, flags=fl)["kernel_ms"] for _ in range(2)))
    eng.run_device(mode=MODE, flags=fl)
    ms = min(eng.run_device(mode=MODE, flags=fl)["kernel_ms"] for _ in range(2))
    print(f"{name:14s} groups of 10 scales: " + " ".join(f"{x:7.2f}" for x in row) + f" ms  sum {sum(row):7.2f}  all at once {ms:7.2f} ms = {S * len(sc) * cfg.n_timepoints / ms / 1e6:7.2f} Gsteps/s", flush=True)
    for k in keys:
        os.environ.pop(k, None)
