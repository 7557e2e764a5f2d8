"""Summarise an `ncu --page source --csv` export: per kernel, the SASS lines with most executed instructions / stall samples."""
import csv
import sys


def num(x):
    try:
        return int(float(x))
    except Exception:
        return 0


def main(path, topn=30):
    rows = list(csv.reader(open(path)))
    kernels, cur, hdr = [], None, None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            kernels.append(cur)
            hdr = None
        elif r and r[0] == "Address":
            hdr = {h: i for i, h in enumerate(r)}
            cur["hdr"] = hdr
        elif cur is not None and hdr is not None and len(r) >= len(hdr) - 1:
            cur["rows"].append(r)
    for kinfo in kernels[:1]:
        ix, data = kinfo["hdr"], kinfo["rows"]
        ti = sum(num(r[ix["Instructions Executed"]]) for r in data)
        ts = sum(num(r[ix["# Samples"]]) for r in data)
        print(kinfo["name"], "instructions", ti, "samples", ts, "sass lines", len(data))
        stall_cols = [h for h in ix if h.startswith("stall_") and "Not Issued" not in h]
        tot = {h: sum(num(r[ix[h]]) for r in data) for h in stall_cols}
        print("stall totals:", {k: v for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v})
        print("---- by instructions executed")
        for r in sorted(data, key=lambda r: -num(r[ix["Instructions Executed"]]))[:topn]:
            print(r[ix["Address"]][-5:], str(num(r[ix["Instructions Executed"]])).rjust(9), str(num(r[ix["# Samples"]])).rjust(6), r[ix["Source"]][:110])
        print("---- by stall samples")
        for r in sorted(data, key=lambda r: -num(r[ix["# Samples"]]))[:topn]:
            why = max(stall_cols, key=lambda h: num(r[ix[h]]))
            print(r[ix["Address"]][-5:], str(num(r[ix["Instructions Executed"]])).rjust(9), str(num(r[ix["# Samples"]])).rjust(6), why.ljust(18), r[ix["Source"]][:100])


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
