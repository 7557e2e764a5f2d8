"""profiles/traffic.json from ncu captures: DRAM bytes per launch of the kernels bench.py reports a roofline for.
  python tools/ncu_traffic.py kernel_name=report.ncu-rep[:units_per_launch] ...
Each report is read with `ncu -i ... --page raw --csv`; the FIRST profiled launch whose name contains kernel_name counts.
The file is stamped with the hash of the CUDA sources (bench.source_stamp): bench.py ignores it once the kernels change."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def metrics_of(report: str, kernel: str):
    txt = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = rows[0]
    name_col = hdr.index("Kernel Name")
    for r in rows[2:]:
        if kernel in r[name_col]:
            def val(m):
                return float(r[hdr.index(m)].replace(",", "")) if m in hdr else None
            units = rows[1]
            def scaled(m):          # ncu prints bytes in a unit named in the second header row
                v = val(m)
                if v is None:
                    return None
                u = units[hdr.index(m)].lower()
                return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
            t = val("gpu__time_duration.sum")
            tu = units[hdr.index("gpu__time_duration.sum")].lower()
            t_us = t * {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(tu, 1)
            return {"dram_read_bytes": scaled("dram__bytes_read.sum"), "dram_write_bytes": scaled("dram__bytes_write.sum"),
                    "duration_us_under_ncu": t_us,
                    "dram_throughput_pct": val("dram__throughput.avg.pct_of_peak_sustained_elapsed"),
                    "issue_active_pct": val("sm__inst_issued.avg.pct_of_peak_sustained_active") or val("smsp__issue_active.avg.pct"),
                    "kernel": r[name_col][:120]}
    raise SystemExit("no launch of %s in %s" % (kernel, report))


out = {"source_stamp": bench.source_stamp(), "kernels": {}}
for arg in sys.argv[1:]:
    kernel, rest = arg.split("=", 1)
    report, _, units = rest.partition(":")
    m = metrics_of(report, kernel)
    m["dram_bytes_per_launch"] = (m["dram_read_bytes"] or 0) + (m["dram_write_bytes"] or 0)
    if units:
        m["units_per_launch"] = int(float(units))
        m["dram_bytes_per_lookup"] = m["dram_bytes_per_launch"] / m["units_per_launch"]
    m["report"] = os.path.basename(report)
    out["kernels"][kernel] = m
path = os.path.join(ROOT, "profiles", "traffic.json")
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
