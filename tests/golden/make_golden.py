#!/usr/bin/env python3
"""Generate tests/golden/{track_fuse,color_mesh}_m32.npz FROM THE REFERENCE ITSELF.

The reference ships no golden vectors.  Its four hot-path translation units are compiled
unmodified against oracle/shim/ (oracle/Makefile `ref` -> oracle/_ref/libtsdf_ref.so, bound by oracle/pyref.py)
and every point-to-plane fixture below (keys m0_*, exp_*, the colour and mesh file) is an OUTPUT OF THAT LIBRARY
on the deterministic synthetic frames (tools/synth) — with one OpenMP thread where the reference's result depends
on its thread count (the normal-equation sums).  The point-to-point fixtures (m1_*) come from the oracle: the
reference's point-to-point call is commented out (sdf.cpp:267), only the formula (sdf.h:169-172) exists.
/root/reference does not exist on the GPU box, so these committed files are how the reference's results travel.
Run in the build container only:
    python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po   # noqa: E402
from oracle import pyref as pr      # noqa: E402
from tools import synth             # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "track_fuse_m32.npz")
OUT_CM = os.path.join(ROOT, "tests", "golden", "color_mesh_m32.npz")


def build():
    depth, Rs, ts = synth.render_sequence(6)
    out = {"K": synth.K_DEFAULT, "R_gt": Rs, "t_gt": ts,
           "depth_sha256": np.frombuffer(hashlib.sha256(depth.tobytes()).digest(), np.uint8),
           "depth_probe": depth[:, ::60, ::80].copy()}
    pr.set_num_threads(1)
    for metric in (0, 1):
        kw = dict(m=32, metric=metric, gauss_newton_max_iteration=10, maximum_twist_diff=float("-inf"))
        o = pr.Reference(**kw) if metric == 0 else po.Oracle(use_coord_table=0, **kw)
        o.set_intrinsics(synth.K_DEFAULT)
        nupd = []
        for f in range(3):                       # fuse three frames at ground-truth poses
            o.set_pose(Rs[f], ts[f])
            nupd.append(o.fuse(depth[f]))
        out["m%d_n_updated" % metric] = np.array(nupd)
        out["m%d_D" % metric] = o.D.copy()
        out["m%d_W" % metric] = o.W.copy()
        o.set_pose(Rs[3], ts[3])                 # one linearisation at the GT pose of frame 3
        J, psi, flag = o.linearize_pixels(depth[3])
        if metric == 0:
            A, b = o.linearize(depth[3])
            ok = flag == 1
            st = {"n_valid": int(ok.sum()), "n_oob": int((flag == 2).sum()),
                  "residual": float(np.sum(psi[ok].astype(np.float64) ** 2))}   # not a reference quantity: psi is
        else:
            A, b, st = o.linearize(depth[3])
        out["m%d_A" % metric] = A; out["m%d_b" % metric] = b
        out["m%d_lin_stats" % metric] = np.array([st["n_valid"], st["n_oob"], st["residual"]])
        out["m%d_flag" % metric] = flag; out["m%d_J" % metric] = J[::16].copy(); out["m%d_psi" % metric] = psi[::16].copy()
        out["m%d_J_sha256" % metric] = np.frombuffer(hashlib.sha256(J.tobytes() + psi.tobytes()).digest(), np.uint8)
        o.set_pose(Rs[2], ts[2])                 # track frame 3 from frame 2's pose, 10 fixed iterations
        st = o.track(depth[3])
        R, t = o.get_pose()
        out["m%d_R_tracked" % metric] = R; out["m%d_t_tracked" % metric] = t
        out["m%d_track_stats" % metric] = np.array([st["iterations"]])
        pts = np.random.default_rng(7).uniform(-1, 33, (512, 3))
        v, ok = o.interpolate_distance(pts)
        out["m%d_sample_pts" % metric] = pts; out["m%d_sample_val" % metric] = v; out["m%d_sample_ok" % metric] = ok
        o.close()
    tw = np.random.default_rng(11).uniform(-0.5, 0.5, (16, 6)); tw[0] = 0; tw[1, 3:] = 1e-5; tw[2, 3:] = 1e-9
    out["exp_twist"] = tw
    out["exp_R"] = np.stack([pr.exp_map(x)[0] for x in tw]); out["exp_t"] = np.stack([pr.exp_map(x)[1] for x in tw])
    pr.set_num_threads(po.num_threads())
    return out


def build_color_mesh():
    """Colour fusion (sdf.cpp:294-304), colour sampling (sdf.cpp:164-217) and the mesher
    (marching_cubes_sdf.cpp:243-287 + sdf.cpp:352-385) at m = 32 on three synthetic frames."""
    depth, Rs, ts = synth.render_sequence(3)
    o = pr.Reference(m=32, metric=0)
    o.set_intrinsics(synth.K_DEFAULT)
    sha = hashlib.sha256()
    nupd = []
    for f in range(3):
        rgb = synth.synth_rgb(depth[f], Rs[f], ts[f])
        sha.update(rgb.tobytes())
        o.set_pose(Rs[f], ts[f])
        nupd.append(o.fuse_rgb(depth[f], rgb))
    CW, R, G, B = o.color()
    out = {"rgb_sha256": np.frombuffer(sha.digest(), np.uint8), "n_updated": np.array(nupd),
           "Color_W": CW.copy(), "R": R.copy(), "G": G.copy(), "B": B.copy()}
    pts = np.random.default_rng(17).uniform([-3.1, -3.1, -0.6], [3.1, 3.1, 3.1], (2048, 3))
    out["sample_pts"] = pts; out["sample_rgba"] = o.interpolate_color(pts)
    for iso, tag in ((0.0, "iso0"), (0.1, "iso01")):
        xyz = o.mesh(iso)
        out["mesh_%s_xyz" % tag] = xyz
        if iso == 0.0:                               # SDF::visualize extracts at iso 0 only (sdf.cpp:44)
            world, rgba = o.visualize()
        else:                                        # marker arithmetic of sdf.cpp:354-356 + interpolate_color per vertex
            origin = np.array([o.cfg.origin[0], o.cfg.origin[1], o.cfg.origin[2]])
            world = xyz.astype(np.float64) + origin
            rgba = o.interpolate_color(world)
        out["mesh_%s_world" % tag] = world; out["mesh_%s_rgba" % tag] = rgba
    o.close()
    return out


if __name__ == "__main__":
    np.savez_compressed(OUT, **build())
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
    np.savez_compressed(OUT_CM, **build_color_mesh())
    print("wrote", OUT_CM, os.path.getsize(OUT_CM), "bytes")
