"""Build tuning variants of libtsdf_b200.so into gpurun_out-independent scratch dir build/variants (git-ignored)."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools import build as b
out_dir = os.path.join(b.ROOT, "tracking_sdf_b200", "_lib", "variants")
os.makedirs(out_dir, exist_ok=True)
cu, _ = b.cuda_sources()
for spec in sys.argv[1:]:
    name, *defs = spec.split(",")
    out = os.path.join(out_dir, "libtsdf_%s.so" % name)
    cmd = [b.NVCC] + b.NVCC_FLAGS + ["-D" + d for d in defs] + ["-I", os.path.join(b.ROOT, "include"), "-shared", "-o", out] + cu
    r = subprocess.run(cmd + ["-Xptxas", "-v"], capture_output=True, text=True)
    if r.returncode: print(r.stderr); raise SystemExit(1)
    lines = r.stderr.splitlines()
    for i, l in enumerate(lines):
        if "Compiling entry" in l and ("k_fuse_exact" in l or "k_linearize" in l or "k_fuse_cert" in l):
            kn = "k_fuse_exact" if "k_fuse_exact" in l else "k_linearize" if "k_linearize" in l else "k_fuse_cert"
            print(name, kn, "|", lines[i + 1].split(":")[-1].strip() if i + 1 < len(lines) else "", "|", lines[i + 2].split(":")[-1].strip() if i + 2 < len(lines) else "")
