#!/usr/bin/env python3
"""Turn the ncu artefacts brought back in gpurun_out/ into the committed summaries under profiles/.

    python tools/summarize_ncu.py <round-tag> <launches.csv> <full.ncu-rep> [method]
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    seq = []
    for r in data:
        if len(r) > vi:
            seq.append((r[ki], float(r[vi].replace(",", "")) / 1e3))
    return seq


def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, unit, vals = rows[0], rows[1], rows[2:]
    return hdr, unit, vals


def main():
    tag, lpath, rep = sys.argv[1], sys.argv[2], sys.argv[3]
    method = sys.argv[4] if len(sys.argv) > 4 else "DIRECT7"
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    seq = launches(lpath)
    agg = collections.OrderedDict()
    for k, t in seq:
        a = agg.setdefault(k.split("(")[0][:70], [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(ROOT, "profiles", "%s_launches.md" % tag), "w") as f:
        f.write("# %s — ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare shares)\n\n" % tag)
        f.write("Command: see profiles/README.md.  %d launches captured, %.1f us total.\n\n" % (len(seq), tot))
        f.write("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|\n")
        for k, (c, t) in agg.items():
            f.write("| `%s` | %d | %.1f | %.2f | %.1f%% |\n" % (k, c, t, t / c, 100 * t / tot))
        f.write("\nSequence (kernel, us):\n\n```\n")
        for k, t in seq:
            f.write("%-60s %9.2f\n" % (k.split("(")[0][:60], t))
        f.write("```\n")
    hdr, unit, vals = raw_metrics(rep)
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
            "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
            "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
    stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    with open(os.path.join(ROOT, "profiles", "%s_align_kernel.md" % tag), "w") as f:
        f.write("# %s — ncu --set full of ndt_align_kernel<%s> (%d launches captured)\n\n" % (tag, method, len(vals)))
        f.write("| metric | " + " | ".join("launch %d" % i for i in range(len(vals))) + " | unit |\n|---|" + "---|" * (len(vals) + 1) + "\n")
        for w in want + stalls:
            if w in hdr:
                i = hdr.index(w)
                f.write("| %s | %s | %s |\n" % (w, " | ".join(r[i] for r in vals), unit[i]))
    i_r, i_w = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")

    def to_bytes(v, u):
        return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    reads = [to_bytes(r[i_r], unit[i_r]) for r in vals]
    writes = [to_bytes(r[i_w], unit[i_w]) for r in vals]
    jpath = os.path.join(ROOT, "profiles", "align_kernel_ncu.json")
    js = json.load(open(jpath)) if os.path.exists(jpath) else {}
    def avg(name):
        if name not in hdr:
            return None
        i = hdr.index(name)
        return sum(float(r[i]) for r in vals) / len(vals)
    js[method] = {"dram_bytes_per_launch": sum(reads) / len(reads) + sum(writes) / len(writes),
                  "dram_read_bytes_per_launch": sum(reads) / len(reads), "dram_write_bytes_per_launch": sum(writes) / len(writes),
                  "issue_active_pct": avg("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                  "warps_active_pct": avg("sm__warps_active.avg.pct_of_peak_sustained_active"),
                  "duration_us_alone": avg("gpu__time_duration.sum"),
                  "source": "%s (ncu --set full, %d launches)" % (os.path.basename(rep), len(vals))}
    json.dump(js, open(jpath, "w"), indent=1)
    print("wrote profiles/%s_launches.md, profiles/%s_align_kernel.md, profiles/align_kernel_ncu.json" % (tag, tag))


if __name__ == "__main__":
    main()
