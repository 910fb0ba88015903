"""Group the SASS of one launch in an .ncu-rep into regions of similar execution count:
python scripts/ncu_regions.py rep [launch index] [--all]  -> instructions and stall samples per region"""
import csv, subprocess, sys, io
rep = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
kern = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name": cur = {"name": r[1], "rows": []}; kern.append(cur); continue
    if r and r[0] == "Address": cur["hdr"] = r; continue
    if cur is not None and r: cur["rows"].append(r)
k = kern[which]; h = k["hdr"]; iS = h.index("# Samples"); iI = h.index("Instructions Executed")
R = k["rows"]
tot = sum(int(r[iS]) for r in R); totI = sum(int(r[iI]) for r in R)
print(k["name"], "samples", tot, "inst", totI, "sass", len(R))
if "--all" in sys.argv:
    for i, r in enumerate(R): print(i, r[1][:70].ljust(70), r[iI].rjust(9), r[iS].rjust(5))
    sys.exit()
i = 0
while i < len(R):
    j = i; c = int(R[i][iI]); s = 0; ins = 0; top = (0, "")
    while j < len(R) and abs(int(R[j][iI]) - c) <= 0.15 * max(c, 1) + 1000:
        s += int(R[j][iS]); ins += int(R[j][iI])
        if int(R[j][iS]) > top[0]: top = (int(R[j][iS]), R[j][1][:40])
        j += 1
    if ins > totI * 0.004 or s > tot * 0.01:
        print(f"{i:4d}-{j-1:4d} n={j-i:3d} exec~{c:8d} inst={ins/1e6:6.2f}M ({100*ins/totI:4.1f}%) samples={s:5d} ({100*s/tot:4.1f}%) top: {top[0]} {top[1]}")
    i = j
