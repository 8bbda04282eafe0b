"""Top stall-sample instructions of a kernel from an ncu report (source page, SASS view).
usage: python profiles/ncu_hot_lines.py report.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys


def main(path, top=25):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    kernels, cur, hdr = [], None, None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            kernels.append(cur)
            hdr = None
        elif r and r[0] == "Address":
            hdr = {h: i for i, h in enumerate(r)}
            cur["hdr"] = hdr
        elif cur is not None and hdr is not None and len(r) > 5:
            cur["rows"].append(r)
    for k in kernels[:1]:
        h = k["hdr"]
        tot = sum(int(r[h["# Samples"]] or 0) for r in k["rows"])
        print(f"== {k['name'][:100]}  total samples {tot}, instructions {len(k['rows'])}")
        stall_cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
        agg = {c: sum(int(r[h[c]] or 0) for r in k["rows"]) for c in stall_cols}
        print("   stall mix:", {c[6:]: v for c, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.01 * tot})
        idx = sorted(range(len(k["rows"])), key=lambda i: -int(k["rows"][i][h["# Samples"]] or 0))[:top]
        for i in sorted(idx):
            r = k["rows"][i]
            s = int(r[h["# Samples"]] or 0)
            why = max(stall_cols, key=lambda c: int(r[h[c]] or 0))
            print(f"   #{i:4d} {100.0 * s / max(tot, 1):5.1f}%  {why[6:]:14s} {r[h['Source']].strip()[:90]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
