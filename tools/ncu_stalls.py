"""Per-instruction stall summary of an ncu --page source --csv export (ncu -i X.ncu-rep --page source --csv > f.csv).
usage: python tools/ncu_stalls.py f.csv [kernel-substring] [min-percent]
Prints the kernel's stall-reason totals and every SASS instruction holding at least min-percent of the samples."""
import csv
import sys

path = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
rows = list(csv.reader(open(path)))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None:
        cur["data"].append(r)
seen = set()
for b in blocks:
    if want not in b["name"] or b["name"] in seen or not b["data"]:
        continue
    seen.add(b["name"])
    hdr = b["hdr"]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[ix["# Samples"]] or 0) for r in b["data"] if len(r) >= len(hdr))
    inst = sum(int(r[ix["Instructions Executed"]] or 0) for r in b["data"] if len(r) >= len(hdr))
    print("==", b["name"][:90], "samples", tot, "warp-instructions", inst)
    agg = {s: sum(int(r[ix[s]] or 0) for r in b["data"] if len(r) >= len(hdr)) for s in stalls}
    print("   " + "  ".join("%s %.1f%%" % (s[6:], 100.0 * v / max(tot, 1)) for s, v in sorted(agg.items(), key=lambda x: -x[1]) if v * 200 > tot))
    for i, r in enumerate(b["data"]):
        if len(r) < len(hdr):
            continue
        n = int(r[ix["# Samples"]] or 0)
        if n * 100.0 >= minpct * tot:
            st = {s[6:]: int(r[ix[s]] or 0) for s in stalls if int(r[ix[s]] or 0) > 0}
            print("  %5d %5.2f%% ie=%8s  %-64s %s" % (i, 100.0 * n / tot, r[ix["Instructions Executed"]], r[ix["Source"]].strip()[:64], st))
