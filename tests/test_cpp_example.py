"""The C++ host side: include/tracking_sdf_b200.hpp + examples/sdf_reconstruction_loop.cpp (the
reference node's per-frame loop on the library).  CPU: it compiles, links, and fails loudly
without a GPU.  GPU: it runs the loop and writes a TUM-format trajectory (sdf_reconstruction.cpp:4-17)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "_build", "sdf_loop")
EXE_SHARDED = os.path.join(ROOT, "tests", "_build", "sharded_loop")


def build_exe(src="sdf_reconstruction_loop.cpp", exe=EXE):
    import tracking_sdf_b200 as T
    T.load_library()
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    lib_dir = os.path.join(ROOT, "tracking_sdf_b200", "_lib")
    cmd = ["/usr/bin/g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", src), os.path.join(ROOT, "tools", "synth.cpp"),
           "-L" + lib_dir, "-ltsdf_b200", "-Wl,-rpath," + lib_dir, "-fopenmp", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_cpp_example_builds_and_has_no_fallback(tmp_path):
    import tracking_sdf_b200 as T
    build_exe()
    if T.load_library().tsdf_device_count() > 0:
        pytest.skip("a GPU is present")
    r = subprocess.run([EXE, os.path.join(ROOT, "data", "fr1_plant_gt_every4.txt"), "3", "64", str(tmp_path / "t.txt")],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_cpp_example_runs_the_node_loop(tmp_path, gpu_lib):
    build_exe()
    out = tmp_path / "trajectory.txt"
    ply = tmp_path / "mesh.ply"
    r = subprocess.run([EXE, os.path.join(ROOT, "data", "fr1_plant_gt_every4.txt"), "8", "128", str(out), str(ply)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    head = open(ply).read(400).split("\n")
    nv = int([l for l in head if l.startswith("element vertex")][0].split()[-1])
    nf = int([l for l in head if l.startswith("element face")][0].split()[-1])
    assert nv > 3000 and nv == 3 * nf                # the visualisation thread's coloured triangle soup (sdf.cpp:352-385)
    rows = np.loadtxt(out)
    assert rows.shape == (7, 8)                       # frame 1 is not tracked (sdf_reconstruction.cpp:69)
    gt = np.loadtxt(os.path.join(ROOT, "data", "fr1_plant_gt_every4.txt"))[1:8]
    assert np.allclose(rows[:, 0], gt[:, 0], atol=1e-3)
    assert np.linalg.norm(rows[:, 1:4] - gt[:, 1:4], axis=1).max() < 0.06
    assert np.allclose(np.linalg.norm(rows[:, 4:8], axis=1), 1.0, atol=1e-3)


def test_sharded_cpp_host_builds_and_has_no_fallback():
    """examples/sharded_reconstruction.cpp: the multi-GPU host in C++ (b200::ShardedSDF over tsdf_shard_attach_local /
    tsdf_group_*)."""
    import tracking_sdf_b200 as T
    build_exe("sharded_reconstruction.cpp", EXE_SHARDED)
    if T.load_library().tsdf_device_count() > 0:
        pytest.skip("a GPU is present")
    r = subprocess.run([EXE_SHARDED, os.path.join(ROOT, "data", "fr1_plant_gt_every4.txt"), "3", "64", "2", "1"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("shards", [2, 3])
def test_sharded_cpp_host_matches_the_unsharded_volume(gpu_lib, shards):
    """One C++ host thread drives `shards` z slabs (over every visible GPU; on a single-GPU box they share the device)
    through the node's loop and checks itself against the unsharded volume: poses to rounding, owned voxels."""
    build_exe("sharded_reconstruction.cpp", EXE_SHARDED)
    r = subprocess.run([EXE_SHARDED, os.path.join(ROOT, "data", "fr1_plant_gt_every4.txt"), "6", "128", str(shards), "1"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "verify:" in r.stdout, r.stdout
    print(r.stdout)
