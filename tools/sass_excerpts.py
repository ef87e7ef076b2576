"""Write profiles/<tag>_sass_<kernel>.txt: the SASS lines of every tensor-core / TMA kernel of
libgwen_b200.so that carry the Blackwell-native instructions (tcgen05.mma = UTC*MMA, tcgen05.ld = LDTM,
tcgen05.commit = UTCBAR, TMA = UTMALDG / UTMASTG / UBLKCP / UTMAPF, packed fp32x2 math, cluster barriers,
system-scope peer accesses), with their addresses, straight from `cuobjdump -sass`.
  python tools/sass_excerpts.py [tag]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
so = os.path.join(ROOT, "gwen_b200", "libgwen_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTC[A-Z]*MMA|LDTM|STTM|UTCBAR|UTMALDG|UTMASTG|UBLKCP|UTMAPF|UTMACCTL|UCGABAR_ARV|UCGABAR_WAIT|"
                 r"FFMA2|FADD2|FMUL2|HMNMX2|SYNCS|FENCE\.VIEW\.ASYNC|LDG\.E[.0-9A-Z]*\.SYS|STG\.E[.0-9A-Z]*\.SYS|"
                 r"RED\.E[.0-9A-Z]*|ATOM[.0-9A-Z]*|LDGSTS[.0-9A-Z]*|LDGSTSBAR[.0-9A-Z]*)\b")
want = ("k_linear_tc3", "k_gcn_fused", "k_wgrad_tc", "k_linear_tf32x3", "k_wgrad_tf32x3", "k_grid_stencil", "k_agg_tiled",
        "k_linear_tc2", "k_linear_b2b")
cur, funcs = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    if cur and re.match(r"\s*/\*[0-9a-f]{4,}\*/", line):
        funcs[cur].append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", line).rstrip())
demangle = subprocess.run(["c++filt"] + list(funcs), capture_output=True, text=True).stdout.splitlines()
done = set()
for (mangled, lines), name in zip(funcs.items(), demangle):
    short = next((w for w in want if w in name), None)
    if short is None:
        continue
    key = short + ("_bf16" if "__nv_bfloat16" in name else ("_f32" if "<float" in name else ""))
    if key in done:
        continue
    done.add(key)
    hits = [l for l in lines if pat.search(l)]
    cnt = collections.Counter(pat.search(l).group(1) for l in hits)
    path = os.path.join(ROOT, "profiles", "%s_sass_%s.txt" % (tag, key))
    with open(path, "w") as f:
        f.write("# %s\n# cuobjdump -sass gwen_b200/libgwen_b200.so (sm_100a), %d instructions; lines below: the %d that are\n"
                "# tcgen05 / TMEM / TMA / mbarrier / packed-fp32 / cluster / system-scope instructions\n# counts: %s\n"
                % (name, len(lines), len(hits), ", ".join("%s x%d" % kv for kv in sorted(cnt.items()))))
        # keep the file readable: every tensor / TMA line, at most 12 of each high-volume arithmetic mnemonic
        seen = collections.Counter()
        for l in hits:
            op = pat.search(l).group(1)
            seen[op] += 1
            if op in ("FFMA2", "FADD2", "FMUL2", "HMNMX2", "SYNCS") and seen[op] > 12:
                continue
            f.write(l + "\n")
    print(path, dict(cnt))
