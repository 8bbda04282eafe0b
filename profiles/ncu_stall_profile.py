"""Stall mix, samples along the instruction stream and top stall instructions of one kernel of an ncu report (SASS view).
usage: python profiles/ncu_stall_profile.py report.ncu-rep <kernel-name-substring> <top_n>"""
import csv, io, subprocess, sys
path, want, top = sys.argv[1], sys.argv[2], int(sys.argv[3])
raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
kernels, cur, hdr = [], None, None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; kernels.append(cur); hdr = None
    elif r and r[0] == "Address":
        hdr = {h: i for i, h in enumerate(r)}; cur["hdr"] = hdr
    elif cur is not None and hdr is not None and len(r) > 5:
        cur["rows"].append(r)
ks = [k for k in kernels if want in k["name"]]
k = ks[-1]
h = k["hdr"]
tot = sum(int(r[h["# Samples"]] or 0) for r in k["rows"])
print(f"== {k['name'][:110]}  total samples {tot}, instructions {len(k['rows'])}")
stall_cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
agg = {c: sum(int(r[h[c]] or 0) for r in k["rows"]) for c in stall_cols}
print("   stall mix:", {c[6:]: round(100.0*v/tot,1) for c, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.01 * tot})
# cumulative profile along the instruction stream in 20 buckets
n=len(k["rows"]); B=24
for bi in range(B):
    lo,hi=bi*n//B,(bi+1)*n//B
    s=sum(int(r[h["# Samples"]] or 0) for r in k["rows"][lo:hi])
    ops={}
    for r in k["rows"][lo:hi]:
        op=r[h['Source']].strip().split()[0] if r[h['Source']].strip() else ''
        if op.startswith('@'): op=r[h['Source']].strip().split()[1]
        ops[op.split('.')[0]]=ops.get(op.split('.')[0],0)+1
    topops=sorted(ops.items(), key=lambda kv:-kv[1])[:4]
    print(f"   instr {lo:5d}-{hi:5d}: {100.0*s/tot:5.1f}%  {topops}")
idx = sorted(range(n), key=lambda i: -int(k["rows"][i][h["# Samples"]] or 0))[:top]
for i in sorted(idx):
    r = k["rows"][i]; s = int(r[h["# Samples"]] or 0)
    why = max(stall_cols, key=lambda c: int(r[h[c]] or 0))
    print(f"   #{i:4d} {100.0 * s / max(tot, 1):5.1f}%  {why[6:]:14s} {r[h['Source']].strip()[:100]}")
