"""Per-CUDA-source-line executed warp instructions from an ncu report taken with --import-source on (kernels built with -lineinfo):
  python tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

rep, topn = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
fname, lines, ie = "", {}, None
for r in csv.reader(io.StringIO(txt)):
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        ie = r.index("Instructions Executed")
    elif ie is not None and len(r) > ie and r[0].isdigit():
        key = (fname, int(r[0]))
        try:
            n = int(float(r[ie] or 0))
        except ValueError:
            n = 0
        if n:
            cur = lines.get(key, [0, r[1]])
            cur[0] += n
            lines[key] = cur
tot = sum(v[0] for v in lines.values())
print("total warp instructions", tot)
for (f, ln), (n, src) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:topn]:
    print("%10d %5.1f%%  %s:%d  %s" % (n, 100.0 * n / tot, f, ln, src.strip()[:110]))
