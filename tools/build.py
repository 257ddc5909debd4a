"""Build recipes for every native artefact in the repo (all in-tree, all git-ignored .so files).

  build_cuda()   -> tracking_sdf_b200/_lib/libtsdf_b200.so   (nvcc, sm_100a; THE product)
  build_oracle() -> oracle/_build/liboracle.so               (g++; test infrastructure)
  build_synth()  -> tools/_build/libsynth.so                 (g++; synthetic depth frames)
  build_emul()   -> tests/_build/libcore_emul.so             (g++; host compile of the kernels'
                                                              __host__ __device__ core, CPU tests only)

  build_ref()    -> oracle/_ref/libtsdf_ref.so               (g++; THE REFERENCE ITSELF: its four hot-path
                                                              translation units compiled unmodified from
                                                              /root/reference against oracle/shim/; only where
                                                              the reference sources are present; pins the oracle)
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
NVCC = "/usr/local/cuda/bin/nvcc" if os.path.exists("/usr/local/cuda/bin/nvcc") else "nvcc"

# No fast-math, no FMA contraction, no -march: the fp32/fp64 rounding sequence is the spec.
HOST_FLAGS = ["-O2", "-std=c++17", "-fopenmp", "-ffp-contract=off", "-fPIC", "-Wall", "-Wno-unknown-pragmas"]
# -fmad=false: device arithmetic must round exactly like the reference's non-contracted x86 code.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-fmad=false", "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "-Xcompiler", "-fno-strict-aliasing"]


def _newer(out, srcs):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(s) > t for s in srcs)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build failed: " + " ".join(cmd))
    if verbose and r.stdout.strip():
        print(r.stdout)
    return r.stdout


def build_oracle(force=False, verbose=False):
    d = os.path.join(ROOT, "oracle")
    out = os.path.join(d, "_build", "liboracle.so")
    srcs = [os.path.join(d, "oracle.cpp"), os.path.join(d, "oracle.h")]
    if force or _newer(out, srcs):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        _run([GXX] + HOST_FLAGS + ["-shared", "-o", out, srcs[0]], verbose)
    return out


def build_ref(force=False, verbose=False):
    """oracle/Makefile target `ref`.  Returns the library path, or None when neither the reference sources
    (/root/reference, build container only) nor a prebuilt library are present."""
    d = os.path.join(ROOT, "oracle")
    out = os.path.join(d, "_ref", "libtsdf_ref.so")
    if not os.path.exists("/root/reference/src/src/sdf.cpp"):
        return out if os.path.exists(out) else None
    _run(["make", "-C", d, "ref"] + (["-B"] if force else []), verbose)
    return out


def build_synth(force=False, verbose=False):
    d = os.path.join(ROOT, "tools")
    out = os.path.join(d, "_build", "libsynth.so")
    srcs = [os.path.join(d, "synth.cpp")]
    if force or _newer(out, srcs):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        _run([GXX] + HOST_FLAGS + ["-shared", "-o", out, srcs[0]], verbose)
    return out


def cuda_sources():
    d = os.path.join(ROOT, "tracking_sdf_b200", "csrc")
    cu = sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cu"))
    hdr = sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith((".cuh", ".h", ".hpp")))
    hdr += [os.path.join(ROOT, "include", f) for f in sorted(os.listdir(os.path.join(ROOT, "include")))]
    return cu, hdr


def build_cuda(force=False, verbose=False, extra=()):
    cu, hdr = cuda_sources()
    out = os.path.join(ROOT, "tracking_sdf_b200", "_lib", "libtsdf_b200.so")
    if force or _newer(out, cu + hdr):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        cmd = [NVCC] + NVCC_FLAGS + list(extra) + ["-I", os.path.join(ROOT, "include"), "-shared", "-o", out] + cu
        _run(cmd, verbose)
    return out


def build_emul(force=False, verbose=False):
    d = os.path.join(ROOT, "tests", "host_emul")
    out = os.path.join(ROOT, "tests", "_build", "libcore_emul.so")
    srcs = [os.path.join(d, "core_emul.cpp")]
    _, hdr = cuda_sources()
    if force or _newer(out, srcs + hdr):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        _run([GXX] + HOST_FLAGS + ["-I", os.path.join(ROOT, "tracking_sdf_b200", "csrc"),
                                   "-I", os.path.join(ROOT, "include"), "-shared", "-o", out, srcs[0]], verbose)
    return out


if __name__ == "__main__":
    v = True
    what = sys.argv[1:] or ["oracle", "synth", "cuda"]
    for w in what:
        print(w, "->", {"oracle": build_oracle, "synth": build_synth, "cuda": build_cuda, "emul": build_emul, "ref": build_ref}[w](force=True, verbose=v))
