"""Stall samples per CUDA source line of an .ncu-rep (cuda,sass source view; -lineinfo + --import-source on).
usage: python tools/ncu_lines.py rep [N]   -- inlined helpers are listed under their own file (tma.cuh ...)"""
import csv, io, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname = None
res = []
h = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        h = r; iS = h.index("# Samples"); iX = h.index("Instructions Executed"); continue
    if h is None or len(r) < len(h) - 5:
        continue
    if r[0].isdigit() and r[iS].isdigit():
        res.append((int(r[iS]), fname, int(r[0]), r[1].strip()[:110], r[iX]))
tot = sum(x[0] for x in res)
print("total samples", tot)
for s, f, ln, src, ix in sorted(res, key=lambda t: -t[0])[:n]:
    print("%7d %5.1f%%  %-18s:%-4d %-110s  inst %s" % (s, 100.0 * s / max(tot, 1), f, ln, src, ix))
