"""Top stalled source lines / SASS instructions of an .ncu-rep (source page; compile with -lineinfo and
capture with --import-source on).  usage: python tools/ncu_src.py rep [N] [sass|cuda]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
view = sys.argv[3] if len(sys.argv) > 3 else "sass"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", view], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
his = [i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r]
hi = his[0]
end = his[1] - 1 if len(his) > 1 else len(rows)      # first captured launch only
h = rows[hi]; data = [r for r in rows[hi + 1:end] if len(r) == len(h) and r[h.index("# Samples")].isdigit()]
iS, isrc = h.index("# Samples"), h.index("Source")
stall = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(int(r[iS]) for r in data)
agg = {}
for r in data:
    for i in stall:
        agg[h[i][6:]] = agg.get(h[i][6:], 0) + int(r[i])
print("total samples", tot, {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for idx, r in enumerate(data):
    r.append(idx)
for r in sorted(data, key=lambda r: -int(r[iS]))[:n]:
    st = {h[i][6:]: int(r[i]) for i in stall if int(r[i]) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print("%6d  #%-5d %-90s %s" % (int(r[iS]), r[-1], r[isrc].strip()[:90], st))
