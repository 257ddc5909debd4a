"""bench.py contract checks that need no GPU: the reference arm (oracle/_ref — the reference's own translation
units — timed on the host cores; the oracle port only where that library is absent) prints exactly one JSON line
with the keys the driver reads, with the SAME `config` as the CUDA arm would print, and with an explicit OpenMP
thread count even under a launcher that exports OMP_NUM_THREADS=1; the product arm refuses to run without a
CUDA device."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args, timeout=600, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=e)


def test_reference_arm_prints_one_contract_line():
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the arm must not inherit it
    r = run(["--impl", "reference", "--grid", "64", "--steps", "3", "--warmup", "1", "--ref-budget", "5"], env={"OMP_NUM_THREADS": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1                                        # ONE JSON line on stdout, everything else on stderr
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches", "impl"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    from oracle import pyref as pr
    assert cb["kind"] == ("reference" if pr.available() else "port") and cb["value"] == d["value"] and "sample" in cb
    assert cb["cores"] == len(os.sched_getaffinity(0))            # all host cores, stated
    import bench
    assert d["config"] == bench.workload_config(64, 1)            # the one definition both arms print
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["gpu_launches"] == 0


def test_product_arm_fails_loudly_without_a_gpu():
    import tracking_sdf_b200 as T
    if T.load_library().tsdf_device_count() > 0:
        import pytest
        pytest.skip("a GPU is present")
    r = run(["--steps", "2", "--warmup", "1", "--no-cpu", "--grid", "64"], timeout=300)
    assert r.returncode != 0                                     # no CPU fallback, no fabricated line
    assert not [l for l in r.stdout.splitlines() if l.strip().startswith("{")]
