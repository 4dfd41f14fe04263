"""profiles/traffic.json from ncu few-metric captures of the walk launches, stamped with the identity of the kernel sources (bench.source_stamp):
bench.py reports `roofline.traffic` from it only while the stamp matches the build it runs.

  python scripts/make_traffic.py KEY=capture.csv[:SPINS] ...      e.g.  c2:fast=profiles/r02_traffic_c2.csv:10000000

A capture is the CSV of
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,smsp__inst_executed.sum,
      smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum
      --clock-control none -k regex:walk_fast --csv --log-file capture.csv python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-extras ...
Per pass bench.py launches the STATS kernel variants once (untimed counters) and the plain variants once; the plain ones (first template argument 0) are summed."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "second": 1e3, "s": 1e3}


def parse(path):
    rows = [r for r in csv.reader(open(path)) if r]
    hdr = next(r for r in rows if "Kernel Name" in r and "Metric Name" in r)
    i = {k: hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
    launches = {}
    for r in rows[rows.index(hdr) + 1:]:
        if len(r) != len(hdr):
            continue
        d = launches.setdefault(r[i["ID"]], {"name": r[i["Kernel Name"]]})
        v = float(r[i["Metric Value"]].replace(",", ""))
        d[r[i["Metric Name"]]] = v * UNIT.get(r[i["Metric Unit"]], 1.0)
    return [d for d in launches.values() if "walk_fast_kernel<0," in d["name"] or "walk_fast_kernel<false" in d["name"] or "walk_compat_kernel<0" in d["name"]]


out_path = os.path.join(ROOT, "profiles", "traffic.json")
out = {}
for arg in sys.argv[1:]:
    key, _, rest = arg.partition("=")
    path, _, spins = rest.partition(":")
    ks = parse(path)
    if not ks:
        raise SystemExit(f"{path}: no plain walk kernel launch found")
    ent = {"dram_bytes_read": sum(k.get("dram__bytes_read.sum", 0.0) for k in ks), "dram_bytes_write": sum(k.get("dram__bytes_write.sum", 0.0) for k in ks),
           "kernel_ms_under_ncu": sum(k.get("gpu__time_duration.sum", 0.0) for k in ks), "spins_per_gpu": int(spins) if spins else None,
           "kernel_stamp": bench.source_stamp(), "source": f"{os.path.relpath(path, ROOT)} (ncu few-metric capture of the walk launch(es) of one pass; kernel sources {bench.source_stamp()})",
           "launches": [{"kernel": k["name"][:90], "ms": k.get("gpu__time_duration.sum"), "dram_read": k.get("dram__bytes_read.sum"), "dram_write": k.get("dram__bytes_write.sum"),
                         "l1_hit_pct": k.get("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": k.get("lts__t_sector_hit_rate.pct"),
                         "issue_active_pct": k.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                         "threads_per_inst": k.get("smsp__thread_inst_executed_per_inst_executed.ratio"), "warp_inst": k.get("smsp__inst_executed.sum")} for k in ks]}
    ent["dram_bytes_per_launch"] = ent["dram_bytes_read"] + ent["dram_bytes_write"]
    out[key] = ent
json.dump(out, open(out_path, "w"), indent=1)
print(json.dumps(out, indent=1))
