#!/usr/bin/env python3
"""SASS evidence for profiles/: per kernel of libtsdf_b200.so the instruction count and the mnemonic histogram
(cuobjdump -sass), plus the full listing of the hot kernels, gzip-compressed.
    python tools/sass_summary.py [tag]      (no GPU needed)"""
import collections, gzip, os, re, subprocess, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"
lib = os.path.join(R, "tracking_sdf_b200", "_lib", "libtsdf_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", txt)), capture_output=True, text=True).stdout.splitlines()
blocks = re.split(r"\n\s*Function : ", txt)[1:]
HOT = ("k_linearize<true>", "k_fuse_cert<0>", "k_fuse_exact<0, 0, false>", "k_fuse_plan", "k_mc_sweep<false>", "k0_normals")
out = ["# cuobjdump -sass tracking_sdf_b200/_lib/libtsdf_b200.so (sm_100a), per kernel: instructions, top mnemonics",
       "# no tensor-core (UTC*MMA / HMMA) and no TMA (UTMALDG) instructions by design: neither stage is a contraction and the access",
       "# patterns are gathers (tracker) and 32-byte read-modify-write sectors (fusion); LDG.E.*.CONSTANT = read-only path (data from the",
       "# preprocessing stream), plain LDG.E = coherent path (data written by the programmatic-dependent-launch predecessor)", ""]
hot_txt = []
for name, b in zip(names, blocks):
    ins = re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", b, flags=re.M)
    h = collections.Counter(i.split(".")[0] for i in ins)
    full = collections.Counter(ins)
    short = name.split("(")[0].replace("void ", "").replace("tsdf::", "")
    flags = []
    for key in ("UTMALDG", "UTCHMMA", "HMMA", "LDGSTS"):
        if any(key in i for i in ins):
            flags.append(key)
    nc = sum(v for k, v in full.items() if k.startswith("LDG") and "CONSTANT" in k); coh = sum(v for k, v in full.items() if k.startswith("LDG") and "CONSTANT" not in k)
    out.append("%-46s %6d instr | LDG nc %3d coherent %3d | STG %3d | DFMA/DADD/DMUL %4d | MUFU %3d | SHFL %3d | %s%s" % (
        short[:46], len(ins), nc, coh, h.get("STG", 0), h.get("DFMA", 0) + h.get("DADD", 0) + h.get("DMUL", 0), h.get("MUFU", 0), h.get("SHFL", 0),
        " ".join("%s %d" % kv for kv in h.most_common(6)), ("  [" + ",".join(flags) + "]") if flags else ""))
    if any(short.startswith(x.split("<")[0]) and (("<" not in x) or x in short.replace("(bool)1", "true").replace("(bool)0", "false")) for x in HOT):
        hot_txt.append("Function : " + name + "\n" + b)
open(os.path.join(R, "profiles", "%s_sass_summary.txt" % TAG), "w").write("\n".join(out) + "\n")
with gzip.open(os.path.join(R, "profiles", "%s_sass_hot_kernels.txt.gz" % TAG), "wt") as f:
    f.write("\n\n".join(hot_txt))
print("\n".join(out[:60]))
