#!/usr/bin/env python3
"""Per-source-line instruction and stall-sample shares of one profiled kernel: joins the SASS page of an ncu report
(`ncu -i rep --page source --csv`) with the line table of the built library (`nvdisasm -g`).

    python tools/sass_lines.py <report.ncu-rep> <kernel-mangled-substring> [top-N] [demangled-name-substring]
"""
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "toyslam_b200", "lib", "libndt_b200.so")


def line_table(kernel_substr):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, check=True, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    table, cur, active = {}, None, False
    for ln in txt.splitlines():
        if ln.startswith("//--------------------- .text."):
            active = kernel_substr in ln
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m and cur:
            table[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return table


def main():
    rep, ksub = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    # a report may hold several launches: one block per launch ("Kernel Name" row, header row, SASS rows); argv[4] = a
    # substring of the demangled kernel name picks the block (default: the first)
    want = sys.argv[4] if len(sys.argv) > 4 else None
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    hi = starts[0]
    if want is not None:
        hi = next(i for i in starts if i > 0 and rows[i - 1] and rows[i - 1][0] == "Kernel Name" and want in rows[i - 1][1])
    hdr = rows[hi]
    ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    end = next((i for i in starts if i > hi), len(rows) + 1) - 1
    data = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
    base = min(int(r[ia], 16) for r in data)
    tab = line_table(ksub)
    agg = {}
    tot_i = tot_s = 0.0
    for r in data:
        off = int(r[ia], 16) - base
        n_i, n_s = float(r[ii] or 0), float(r[isamp] or 0)
        tot_i += n_i
        tot_s += n_s
        key = tab.get(off, (("?", 0), ""))[0]
        a = agg.setdefault(key, [0.0, 0.0, 0])
        a[0] += n_i
        a[1] += n_s
        a[2] += 1
    src_cache = {}
    print("total warp instructions %.0f, samples %.0f, SASS instructions %d" % (tot_i, tot_s, len(data)))
    print("%-22s %7s %7s %5s  source" % ("file:line", "%inst", "%samp", "sass"))
    for key, (n_i, n_s, cnt) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        f, l = key
        path = os.path.join(ROOT, "toyslam_b200", "csrc", f)
        if f not in src_cache and os.path.exists(path):
            src_cache[f] = open(path).read().split("\n")
        text = src_cache.get(f, [""] * (l + 1))[l - 1].strip()[:100] if l else ""
        print("%-22s %6.2f%% %6.2f%% %5d  %s" % ("%s:%d" % (f, l), 100 * n_i / tot_i, 100 * n_s / max(tot_s, 1), cnt, text))


if __name__ == "__main__":
    main()
