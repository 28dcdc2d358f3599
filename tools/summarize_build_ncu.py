#!/usr/bin/env python3
"""profiles/<tag>_build_launches.md and <tag>_build_kernels.md from the artefacts of `tools/profile_recipe_r02.sh build`.

    python tools/summarize_build_ncu.py <tag> <build_launches.csv> <build_kernels.ncu-rep>
"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    tag, lpath, rep = sys.argv[1:4]
    rows = [r for r in csv.reader(l for l in open(lpath) if l.startswith('"'))]
    h = rows[0]
    ki, mi, vi, idi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
    per = collections.OrderedDict()
    for r in rows[1:]:
        per.setdefault((r[idi], r[ki]), {})[r[mi]] = float(r[vi].replace(",", ""))
    items = [(k, m) for (_, k), m in per.items() if "ndtb200" in k]
    last = []
    for k, m in reversed(items):      # the last build of the run (the timed repetition)
        last.append((k, m))
        if "minmax3d" in k:
            break
    last.reverse()
    with open(os.path.join(ROOT, "profiles", "%s_build_launches.md" % tag), "w") as f:
        f.write("# %s — ncu launch list of ONE 100 M-point map build (`python tools/build_bench.py --points 100000000 --res 1.0 --reps 1`)\n\n" % tag)
        f.write("`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none`; serialised, cold cache.\n\n")
        f.write("| kernel | us | DRAM read MB | DRAM write MB | B/point |\n|---|---|---|---|---|\n")
        tot = tr = tw = 0.0
        for k, m in last:
            t, rd, wr = m["gpu__time_duration.sum"] / 1e3, m["dram__bytes_read.sum"] / 1e6, m["dram__bytes_write.sum"] / 1e6
            tot, tr, tw = tot + t, tr + rd, tw + wr
            name = k.split("(")[0].replace("void ", "").replace("ndtb200::", "")
            f.write("| `%s` | %.1f | %.1f | %.1f | %.1f |\n" % (name, t, rd, wr, (rd + wr) / 100.0))
        f.write("| **total** | **%.1f** | %.1f | %.1f | %.1f |\n" % (tot, tr, tw, (tr + tw) / 100.0))
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, unit, vals = rows[0], rows[1], rows[2:]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
            "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
    stalls = [x for x in hdr if x.startswith("smsp__average_warps_issue_stalled_") and x.endswith("_per_issue_active.ratio")]
    kn = hdr.index("Kernel Name")
    with open(os.path.join(ROOT, "profiles", "%s_build_kernels.md" % tag), "w") as f:
        f.write("# %s — `ncu --set full --clock-control none --import-source on` of the build's two heavy kernels (same command)\n\n" % tag)
        f.write("| metric | " + " | ".join("`%s`" % r[kn].split("(")[0].replace("void ", "")[:40] for r in vals) + " | unit |\n|---|" + "---|" * (len(vals) + 1) + "\n")
        for w in want + stalls:
            if w in hdr:
                i = hdr.index(w)
                f.write("| %s | %s | %s |\n" % (w, " | ".join(r[i] for r in vals), unit[i]))
        f.write("\nPer-source-line shares of the middle payload pass (`python tools/sass_lines.py %s onesweep_payload_kernelILb0 14 'onesweep_payload_kernel<(bool)0>'`):\n\n```\n" % os.path.basename(rep))
        o = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_lines.py"), rep, "onesweep_payload_kernelILb0", "14", "onesweep_payload_kernel<(bool)0>"],
                           capture_output=True, text=True).stdout
        f.write(o + "```\n")
    print("wrote profiles/%s_build_launches.md, profiles/%s_build_kernels.md" % (tag, tag))


if __name__ == "__main__":
    main()
