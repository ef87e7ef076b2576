"""Summarise an .ncu-rep (read on the CPU box with `ncu -i`) into profiles/<name>.{json,md}.

usage: python tools/ncu_summary.py gpurun_out/prof_agg.ncu-rep profiles/r01_k_agg_tiled_cfg2 \
           [--launches gpurun_out/launches.csv] [--alg-bytes N] [--note "..."]
"""
import argparse
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size",
    "launch__block_size", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum",
    "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_global_ld.sum",
    "smsp__sass_inst_executed_op_global_st.sum",
]
STALLS = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
STALL_NAMES = ["short_scoreboard", "long_scoreboard", "mio_throttle", "not_selected", "wait",
               "math_pipe_throttle", "barrier", "lg_throttle", "branch_resolving", "dispatch_stall"]


def read_raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("out")
    ap.add_argument("--launches")
    ap.add_argument("--alg-bytes", type=float)
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    hdr, units, data = read_raw(a.rep)
    kernels = []
    for r in data:
        d = dict(zip(hdr, r))
        k = {"kernel": d.get("Kernel Name"), "metrics": {}, "stalls_per_issue": {}}
        for key in KEYS:
            if key in d and d[key] not in ("", "n/a"):
                k["metrics"][key] = {"value": d[key], "unit": units[hdr.index(key)]}
        for s in STALL_NAMES:
            key = STALLS % s
            if key in d and d[key] not in ("", "n/a"):
                k["stalls_per_issue"][s] = float(d[key])
        kernels.append(k)

    def num(k, key):
        m = k["metrics"].get(key)
        if not m:
            return None
        v = float(m["value"].replace(",", ""))
        u = m["unit"].lower()
        scale = {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0, "usecond": 1e-6, "us": 1e-6,
                 "msecond": 1e-3, "ms": 1e-3, "nsecond": 1e-9, "ns": 1e-9, "second": 1.0, "s": 1.0}.get(u, 1.0)
        return v * scale
    summ = {"report": a.rep, "note": a.note, "kernels": kernels}
    k0 = kernels[0]
    rd, wr, t = num(k0, "dram__bytes_read.sum"), num(k0, "dram__bytes_write.sum"), num(k0, "gpu__time_duration.sum")
    if rd is not None and wr is not None:
        n = len(kernels)
        summ["dram_bytes_per_launch"] = sum((num(k, "dram__bytes_read.sum") or 0) + (num(k, "dram__bytes_write.sum") or 0) for k in kernels) / n
        summ["duration_us_under_ncu"] = sum(num(k, "gpu__time_duration.sum") or 0 for k in kernels) / n * 1e6
    if a.alg_bytes:
        summ["algorithmic_bytes_per_launch"] = a.alg_bytes
    if a.launches:
        per = {}
        txt = [l for l in open(a.launches) if not l.startswith("==")]
        rows = list(csv.DictReader(io.StringIO("".join(txt))))
        for r in rows:
            if r.get("Metric Name") != "gpu__time_duration.sum":
                continue
            v = float(r["Metric Value"].replace(",", ""))
            u = r.get("Metric Unit", "ns")
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(u, 1e-3)
            name = r["Kernel Name"].split("(")[0][-60:]
            c = per.setdefault(name, [0, 0.0])
            c[0] += 1
            c[1] += v
        tot = sum(v[1] for v in per.values()) or 1.0
        summ["launch_list"] = [{"kernel": k, "launches": c, "total_us": round(t_, 1), "share": round(t_ / tot, 4)}
                               for k, (c, t_) in sorted(per.items(), key=lambda kv: -kv[1][1])]
    json.dump(summ, open(a.out + ".json", "w"), indent=1)
    with open(a.out + ".md", "w") as f:
        f.write("# ncu summary: %s\n\n%s\n\n" % (a.rep, a.note))
        for k in kernels[:2]:
            f.write("## %s\n\n| metric | value | unit |\n|---|---|---|\n" % k["kernel"])
            for key, m in k["metrics"].items():
                f.write("| %s | %s | %s |\n" % (key, m["value"], m["unit"]))
            f.write("\nstalls per issue: %s\n\n" % json.dumps(k["stalls_per_issue"]))
        if "dram_bytes_per_launch" in summ:
            f.write("DRAM bytes per launch (read+write, mean over %d captured launches): %.1f MB" % (len(kernels), summ["dram_bytes_per_launch"] / 1e6))
            if a.alg_bytes:
                f.write("; algorithmic bytes %.1f MB (ratio %.2f)" % (a.alg_bytes / 1e6, summ["dram_bytes_per_launch"] / a.alg_bytes))
            f.write("\n\n")
        if "launch_list" in summ:
            f.write("## launch list (ncu --metrics gpu__time_duration.sum, cold-cache serialised: compare shares)\n\n| kernel | launches | total us | share |\n|---|---|---|---|\n")
            for r in summ["launch_list"]:
                f.write("| %s | %d | %.1f | %.1f%% |\n" % (r["kernel"], r["launches"], r["total_us"], 100 * r["share"]))
    print("wrote", a.out + ".json/.md")


if __name__ == "__main__":
    main()
