#!/usr/bin/env python3
"""Aggregate an `ncu --page source --csv --print-source sass,cuda` dump per source line:
instructions executed and stall samples.  Usage: ncu_lines.py file.csv [top_n] [kernel_index]"""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40; which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rows = list(csv.reader(open(path)))
# split into kernels: sections start with "File Path"; a new kernel result repeats the first file
sections = []; cur = None
for r in rows:
    if r and r[0] == "File Path":
        cur = {"file": r[1], "rows": []}; sections.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
first = sections[0]["file"]
kernels = []; k = []
for s in sections:
    if s["file"] == first and k:
        kernels.append(k); k = []
    k.append(s)
kernels.append(k)
agg = collections.OrderedDict(); total_i = 0; total_s = 0
for s in kernels[which]:
    hdr = None
    for r in s["rows"]:
        if r and r[0] == "Line No":
            hdr = r; continue
        if hdr is None or len(r) < 8 or not r[0]:
            continue
        try:
            inst = int(r[7]); samp = int(r[6])
        except ValueError:
            continue
        key = (s["file"].split("/")[-1], int(r[0]))
        a = agg.setdefault(key, [0, 0, r[1]])
        a[0] += inst; a[1] += samp
        total_i += inst; total_s += samp
print("kernel", which, "of", len(kernels), "total warp-instructions", total_i, "samples", total_s)
for (f, ln), (i, smp, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%6.2f%% inst %5.2f%% stall  %s:%d  %s" % (100.0 * i / max(total_i, 1), 100.0 * smp / max(total_s, 1), f, ln, src.strip()[:110]))
