#!/usr/bin/env python3
"""bench.py — frames/s of the track + fuse hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]          this repo's CUDA path
  python bench.py --impl reference [...]                       the reference's CPU algorithm (oracle port)

A step = one 640x480 depth frame tracked (10 fixed Gauss-Newton iterations) and fused into a
512^3 grid (BASELINE.json configs[1]).  Frames are synthetic: raycast from an analytic scene
along the bundled fr1/plant ground-truth camera path (tools/synth).

N = 1   one sequence on one B200.
N > 1   launched by torch.distributed.run, one process per GPU: N independent sequences of the
        same shape, one per GPU (BASELINE.json configs[4]); no data-path collective; value is the
        aggregate frames/s; scaling "weak".  (The z-slab sharded volumes of configs[2,3] are run
        with --workload sharded.)

Numbers in the JSON line
  value      frames/s with every depth frame already resident in HBM, CUDA-event timed on the
             library's stream, max over ranks.  The 1 GiB grid is > L2 (126 MB), so no L2 flush.
  e2e        the same frames through tsdf_track_and_fuse() with HOST (pinned) depth buffers: each
             step includes the H2D copy of its frame and the D2H read of pose + stats; wall clock.
  roofline   the fusion kernel (k_fuse): 16 B per updated voxel / CUDA-event duration vs the
             measured HBM copy peak; dense_fuse = the same kernel on the dense micro-benchmark
             (every voxel of the 512^3 grid updated, 2.147 GB per launch) where north_star's
             ">= 70 % of HBM peak" is defined; roofline_track = the linearisation kernel's
             gather bytes (832 B per valid pixel-iteration, L2-resident) for information.
  cpu_baseline  the oracle port of the reference (all host threads) on the first frames of the
             same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "frames/s (track+fuse) 640x480 @512^3"
GN_ITERS = 10


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def ncu_traffic(key):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) from the committed ncu --set full
    capture summary (profiles/ncu_traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return json.load(open(p)).get(key)
    except Exception:
        return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def dist_env():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def init_dist(world, local, backend):
    if world <= 1:
        return None
    import torch
    import torch.distributed as dist
    if backend == "nccl":
        torch.cuda.set_device(local)
    dist.init_process_group(backend=backend)
    return dist


def barrier_max(dist, value, device=None):
    """barrier, then max over ranks of `value` (plumbing only: torch.distributed)."""
    if dist is None:
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def render_frames(n, start, pinned):
    from tools import synth
    import tracking_sdf_b200.capi as capi
    t0 = time.time()
    buf = capi.pinned_empty((n, synth.HEIGHT, synth.WIDTH), np.float32) if pinned else np.empty((n, synth.HEIGHT, synth.WIDTH), np.float32)
    depth, Rs, ts = synth.render_sequence(n, start=start, out=buf)
    log("[bench] rendered %d synthetic frames in %.1f s" % (n, time.time() - t0))
    return depth, Rs, ts


def host_threads():
    """Host cores this process may use.  Launchers such as torchrun export OMP_NUM_THREADS=1, so the CPU arms
    set their OpenMP thread count to this explicitly instead of inheriting the environment."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def workload_config(m, n_gpus):
    """`config` of the JSON line — built in ONE place so the CUDA arm and the reference arm describe the
    workload with identical strings (the driver compares them)."""
    return {"workload": "%d^3 grid, 640x480 synthetic depth along fr1/plant GT path, %d GN iterations/frame, point-to-plane fusion "
                        "(BASELINE.json configs[1]%s)" % (m, GN_ITERS, "; one independent sequence per GPU, configs[4]" if n_gpus > 1 else ""),
            "grid_bytes": 8 * m ** 3, "l2": "inputs larger than L2 (1 GiB grid vs 126 MB); no flush",
            "gn_iterations": GN_ITERS, "pixel_stride": 3, "frames_source": "tools/synth.py + data/fr1_plant_gt_every4.txt"}


def run_cpu_baseline(depth, Rs, ts, m, budget_s, max_frames):
    """The reference's CPU implementation on the first frames of the workload, all host threads, timed where the
    reference prints its own timings (camera_tracking.cpp:68,243; sdf.cpp:225,306).
    kind "reference": oracle/_ref — the reference's own sdf.cpp / camera_tracking.cpp / eigen_utils.cpp compiled
    unmodified over oracle/shim (its SDF::update includes the colour running mean and the global_coords table);
    kind "port": the oracle restatement, only when that library is absent."""
    from oracle import pyoracle as po
    from oracle import pyref as pr
    from tools import synth
    nthr = host_threads()
    po.set_num_threads(nthr)
    use_ref = os.path.exists(pr.PATH) or pr.sources_present()
    kw = dict(m=m, gauss_newton_max_iteration=GN_ITERS, maximum_twist_diff=float("-inf"))
    t_setup = time.time()
    if use_ref:
        pr.set_num_threads(nthr)
        o = pr.Reference(**kw)
    else:
        o = po.Oracle(use_coord_table=1, **kw)
    o.set_intrinsics(synth.K_DEFAULT)
    o.set_pose(Rs[0], ts[0])
    if use_ref:
        o.fuse(depth[0], count=False); o.timers(reset=True)
    else:
        o.fuse(depth[0])
    setup = time.time() - t_setup
    t_track = t_fuse = 0.0
    n = 0
    n_upd = None
    t_begin = time.time()
    for f in range(1, len(depth)):
        if use_ref:
            cloud, normals = o.backproject(depth[f])           # upstream of the reference (ROS/PCL): not timed
            o.track_cloud(cloud)
            nu = o.fuse_cloud(cloud, normals, None, count=(f == 1))
            if f == 1:
                n_upd = nu
            t_track, t_fuse = o.timers()
        else:
            t0 = time.time(); o.track(depth[f]); t1 = time.time(); nu = o.fuse(depth[f]); t2 = time.time()
            t_track += t1 - t0; t_fuse += t2 - t1
            if f == 1:
                n_upd = nu
        n += 1
        if n >= max_frames or (time.time() - t_begin) > budget_s:
            break
    R, t = o.get_pose()
    err = float(np.linalg.norm(t - ts[n]))
    o.close()
    tot = t_track + t_fuse
    kind = "reference" if use_ref else "port"
    what = ("the reference's own translation units (sdf.cpp, camera_tracking.cpp, eigen_utils.cpp) compiled unmodified over oracle/shim "
            "(oracle/_ref); SDF::update incl. its colour running mean and global_coords table; clouds + normals from the shared K1 definition, not timed"
            if use_ref else "oracle port of the reference (oracle/_ref absent), OpenMP, precomputed global_coords table like sdf.cpp:11")
    return {"value": n / tot, "unit": "frames/s", "cores": nthr, "kind": kind,
            "sample": "frames 1..%d of the same %d^3 workload (frame 0 fused untimed), %d GN iterations each; %s; "
                      "time = the spans the reference prints itself (estimate_new_position + update)" % (n, m, GN_ITERS, what),
            "ms_per_frame_track": 1e3 * t_track / n, "ms_per_frame_fuse": 1e3 * t_fuse / n,
            "voxels_visited_per_s": (m ** 3) * n / t_fuse, "voxels_updated_frame1": n_upd,
            "frames": n, "setup_s": setup, "final_pos_err_m": err}


def main_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores."""
    rank, world, local = dist_env()
    if rank != 0:
        return 0
    n_gpus = max(world, 1, args.gpus)
    n_need = min(args.steps + args.warmup, 64) + 1
    depth, Rs, ts = render_frames(n_need, 0, pinned=False)
    t0 = time.time()
    cb = run_cpu_baseline(depth, Rs, ts, args.m, budget_s=args.ref_budget, max_frames=n_need - 1)
    out = {"metric": METRIC if args.m == 512 else METRIC.replace("512", str(args.m)), "value": cb["value"], "unit": "frames/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / cb["value"],
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 values / f64 geometry",
           "data": "synthetic", "impl": "reference",
           "config": workload_config(args.m, n_gpus),
           "note": "rate measured on a bounded sample of the requested steps (%d frames, %.0f s budget); host CPU only, %d threads; "
                   "one sequence (the CPU arm does not replicate per GPU)" % (cb["frames"], args.ref_budget, cb["cores"]),
           "cpu_baseline": cb,
           "e2e": {"value": cb["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0, "wall_s": time.time() - t0}
    emit(out)
    return 0


def dense_fuse_bench(T, m, K, reps, hbm):
    """Dense micro-benchmark: camera outside the volume looking in, constant depth
    behind the whole volume -> every voxel is in view and in front of the surface -> every voxel
    is updated: traffic is exactly 16 B x m^3 per launch."""
    g = T.Tsdf(T.default_config(m=m, gauss_newton_max_iteration=GN_ITERS, maximum_twist_diff=float("-inf")))
    g.set_intrinsics(K)
    R = np.array([[1, 0, 0], [0, 0, 1], [0, -1, 0]], float)
    t = np.array([0.0, -12.0, 1.25])
    depth = np.full((480, 640), 40.0, np.float32)
    dev = g.dev_alloc(depth.nbytes); g.dev_upload(dev, depth)
    g.set_pose(R, t)
    for _ in range(3):
        g.enqueue_frame(dev, track=0, slot=0)
    g.sync(); g.total_updates(reset=True)
    g.stage_timing_begin(reps)
    for _ in range(reps):
        g.enqueue_frame(dev, track=0, slot=0)
        g.sync()      # the fusion launch is timed alone: without this the NEXT call's preprocessing (second stream) overlaps it
    ms = g.stage_timing_end()
    upd = g.total_updates()
    rmw_ms = g.debug_stream_rmw(10)           # plain RMW stream over the store: the practical ceiling
    g.dev_free(dev); g.close()
    t_fuse = float(ms[:, 2].mean()) * 1e-3
    per_launch = upd / reps
    ach = 16.0 * per_launch / t_fuse / 1e9
    return {"bound": "hbm", "kernel": "k_fuse_cert (row-certified free space: pure read-modify-write of the store)",
            "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": ncu_traffic("fuse_dense"),
            "rmw_stream_ceiling_ms": rmw_ms, "rmw_stream_ceiling_gbs": 16.0 * m ** 3 / (rmw_ms * 1e-3) / 1e9,
            "ms_per_launch": t_fuse * 1e3, "voxels_updated_per_launch": per_launch, "all_voxels_updated": bool(per_launch == m ** 3),
            "voxel_updates_per_s": per_launch / t_fuse}


def color_fuse_bench(T, m, K, reps, hbm, depth_frame, R, t):
    """Colour fusion (sdf.cpp:294-304): (a) the dense pose of dense_fuse_bench with a
    constant image: every voxel gets D/W and Color_W/R/G/B updated = 48 B per voxel; (b) one trajectory
    frame.  Timed per launch with the library's stage events (the fusion stage alone)."""
    from tools import synth
    g = T.Tsdf(T.default_config(m=m, gauss_newton_max_iteration=GN_ITERS, maximum_twist_diff=float("-inf")))
    g.set_intrinsics(K)
    g.enable_color()
    Rd = np.array([[1, 0, 0], [0, 0, 1], [0, -1, 0]], float); td = np.array([0.0, -12.0, 1.25])
    depth = np.full((480, 640), 40.0, np.float32)
    rgb = np.full((480, 640, 3), 128, np.uint8)
    g.fuse_rgb(depth, rgb, Rd, td)
    ts_ = []
    n = 0
    for _ in range(reps):
        n = g.fuse_rgb(depth, rgb, Rd, td)
        ts_.append(float(g.last_stage_ms()[2]))
    t_dense = float(np.mean(ts_)) * 1e-3
    g.reset(); g.set_intrinsics(K)
    rgb_t = synth.synth_rgb(depth_frame, R, t)
    g.fuse_rgb(depth_frame, rgb_t, R, t)
    tt = []
    nt = 0
    for _ in range(reps):
        nt = g.fuse_rgb(depth_frame, rgb_t, R, t)
        tt.append(float(g.last_stage_ms()[2]))
    g.close()
    ach = 48.0 * n / t_dense / 1e9
    return {"bound": "hbm", "kernel": "k_fuse_cert (skip certificates; rows certified as free space bypass the unit queue) + k_fuse_exact<colour> "
                                 "(every updated voxel needs its pixel: normal + rgb)",
            "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "bytes_per_updated_voxel": 48,
            "dense_ms_per_launch": t_dense * 1e3, "dense_voxels_updated": int(n),
            "trajectory_ms_per_launch": float(np.mean(tt)), "trajectory_voxels_updated": int(nt)}


def k0_bench(T, m, K, n_frames, hbm):
    """the node's pre-processing (K0: fast bilateral filter + average-3D-gradient normals,
    sdf_reconstruction.cpp:37-49) on NOISY synthetic depth (tools/synth.add_sensor_noise, seed 1234): frames/s and
    ATE of the tracked path with and without it, same frames, device-resident."""
    from tools import synth, evaluate_ate
    depth, Rs, ts = synth.render_sequence(n_frames)
    noisy = synth.add_sensor_noise(depth, seed=1234)
    out = {"frames": n_frames, "noise": "sigma_z = 0.0012 + 0.0019 (z - 0.4)^2 m, 0.2 % dropped pixels, seed 1234"}
    for tag, pre in (("without_k0", 0), ("with_k0", 1)):
        g = T.Tsdf(T.default_config(m=m, preprocess=pre, gauss_newton_max_iteration=GN_ITERS, maximum_twist_diff=float("-inf")))
        g.set_intrinsics(K)
        ring = g.pose_ring_capacity()
        dev = g.dev_alloc(noisy.nbytes); g.dev_upload(dev, noisy)
        fb = noisy[0].nbytes
        g.set_pose(Rs[0], ts[0])
        g.enqueue_frame(dev, track=0, slot=0)
        for f in range(1, 5):
            g.enqueue_frame(dev + f * fb, track=1, slot=f)
        g.sync()
        g.stage_timing_begin(n_frames - 5)
        g.timer_begin()
        for f in range(5, n_frames):
            g.enqueue_frame(dev + f * fb, track=1, slot=f % ring)
        ms = g.timer_end()
        g.sync()
        stage = g.stage_timing_end()
        lost = 0
        est = []
        for f in range(1, n_frames):
            try:
                est.append(g.read_pose_ring(f % ring)[1])
            except T.TsdfError:
                lost += 1
                est.append(np.full(3, np.nan))
        est = np.array(est)
        okf = np.isfinite(est[:, 0])
        ate, _ = evaluate_ate.ate_rmse(est[okf], ts[1:n_frames][okf], do_align=True)
        out[tag] = {"frames_per_s": (n_frames - 5) / (ms * 1e-3), "ate_rmse_m": ate, "frames_lost": lost,
                    "final_pos_err_vs_gt_m": float(np.linalg.norm(est[okf][-1] - ts[1:n_frames][okf][-1])),
                    "stage_ms": {"prep": float(stage[:, 0].mean()), "track": float(stage[:, 1].mean()), "fuse": float(stage[:, 2].mean())}}
        g.dev_free(dev); g.close()
    out["note"] = ("K0 = 12 launches on the preprocessing stream (bilateral grid: min/max, bins, splat, 6 blur passes, slice; gradients + "
                   "discontinuity map; window normals), overlapped with the previous frame; parity with PCL unpinned (own definition, "
                   "bit-equal to the oracle's: tests/test_gpu_k0.py)")
    return out


def ragged_bench(T, m, K, n_frames):
    """The benchmark frames are complete (every pixel hits the analytic scene).  This sub-record pays for the invalid
    paths too: the same frames with 5 % NaN speckle, a NaN block, a zero block and the outermost 12 pixel rings
    missing (seed 99) — frames/s and the tracked path's error, device-resident."""
    from tools import synth, evaluate_ate
    depth, Rs, ts = synth.render_sequence(n_frames)
    rng = np.random.default_rng(99)
    d = depth.copy()
    d[rng.random(d.shape) < 0.05] = np.nan
    d[:, 100:160, 200:330] = np.nan
    d[:, 300:330, 10:120] = 0.0
    d[:, :12, :] = np.nan; d[:, -12:, :] = np.nan; d[:, :, :12] = np.nan; d[:, :, -12:] = np.nan
    g = T.Tsdf(T.default_config(m=m, gauss_newton_max_iteration=GN_ITERS, maximum_twist_diff=float("-inf")))
    g.set_intrinsics(K)
    ring = g.pose_ring_capacity()
    dev = g.dev_alloc(d.nbytes); g.dev_upload(dev, d)
    fb = d[0].nbytes
    g.set_pose(Rs[0], ts[0])
    g.enqueue_frame(dev, track=0, slot=0)
    for f in range(1, 5):
        g.enqueue_frame(dev + f * fb, track=1, slot=f)
    g.sync()
    g.timer_begin()
    for f in range(5, n_frames):
        g.enqueue_frame(dev + f * fb, track=1, slot=f % ring)
    ms = g.timer_end()
    g.sync()
    est = np.array([g.read_pose_ring(f % ring)[1] for f in range(1, n_frames)])
    st = g.read_pose_ring((n_frames - 1) % ring)[2]
    ate, _ = evaluate_ate.ate_rmse(est, ts[1:n_frames], do_align=True)
    g.dev_free(dev); g.close()
    return {"frames": n_frames, "invalid_pixel_fraction": float((~(np.isfinite(d) & (d > 0))).mean()),
            "frames_per_s": (n_frames - 5) / (ms * 1e-3), "ate_rmse_m": ate, "n_valid_last": int(st["n_valid"]),
            "note": "same trajectory frames with NaN speckle, NaN / zero blocks and a missing border: what the invalid-pixel paths cost"}


def main_cuda(args):
    rank, world, local = dist_env()
    if world != args.gpus and world > 1:
        log("[bench] WORLD_SIZE %d != --gpus %d; using WORLD_SIZE" % (world, args.gpus))
    n_gpus = max(world, 1)
    import tracking_sdf_b200 as T
    from tracking_sdf_b200 import capi
    from tools import synth
    L = T.load_library()                       # raises if the CUDA library is missing: no fallback
    ndev = L.tsdf_device_count()
    if ndev < 1:
        raise SystemExit("bench.py: no CUDA device visible (the product has no CPU path)")
    device = local % ndev
    dist = init_dist(world, local, "nccl")
    tdev = None
    if dist is not None:
        import torch
        tdev = torch.device("cuda", device)
    hbm, peak_src = peaks()
    K = synth.K_DEFAULT
    m = args.m
    W, Ksteps = args.warmup, args.steps
    n_frames = W + Ksteps
    # config 5: independent sequences — each rank starts elsewhere on the path
    depth, Rs, ts = render_frames(n_frames, start=rank * 37, pinned=True)
    frame_bytes = depth[0].nbytes

    cfg = T.default_config(m=m, device=device, gauss_newton_max_iteration=GN_ITERS, maximum_twist_diff=float("-inf"))
    g = T.Tsdf(cfg)
    g.set_intrinsics(K)
    ring = g.pose_ring_capacity()

    # ---------------- value: frames resident in HBM --------------------------------------------
    dev = g.dev_alloc(depth.nbytes)
    g.dev_upload(dev, depth)
    g.set_pose(Rs[0], ts[0])
    g.enqueue_frame(dev, track=0, slot=0)                       # frame 0: fuse at the GT pose
    for f in range(1, W):
        g.enqueue_frame(dev + f * frame_bytes, track=1, slot=f % ring)
    g.sync()
    g.total_updates(reset=True)
    launches0 = g.kernel_launch_count()
    sampler = ClockSampler(device); sampler.start()
    if dist is not None:
        import torch
        dist.barrier(); torch.cuda.synchronize()
    g.timer_begin()
    for f in range(W, n_frames):
        g.enqueue_frame(dev + f * frame_bytes, track=1, slot=f % ring)
    ms_total = g.timer_end()
    g.sync()
    if dist is not None:
        import torch
        torch.cuda.synchronize()
    ms_total = barrier_max(dist, ms_total, tdev)
    clocks = sampler.stop()
    launches = g.kernel_launch_count() - launches0
    n_upd = g.total_updates()
    first_kept = max(W, n_frames - min(ring, Ksteps))
    poses_dev = [g.read_pose_ring(f % ring) for f in range(first_kept, n_frames)]
    last_R, last_t, last_st = poses_dev[-1]
    from tools import evaluate_ate
    est_xyz = np.array([pp[1] for pp in poses_dev]); gt_xyz = ts[first_kept:n_frames]
    ate_aligned, _ = evaluate_ate.ate_rmse(est_xyz, gt_xyz, do_align=True)
    ate_raw, _ = evaluate_ate.ate_rmse(est_xyz, gt_xyz, do_align=False)
    mesh = None
    if rank == 0 and n_gpus == 1 and not args.no_mesh:
        # the visualisation thread's product (sdf.cpp:327, marching_cubes_sdf.cpp:243-287) on the volume just built
        # (all K timed frames fused)
        nv = 0
        tm = []
        for _ in range(3):
            t0m = time.perf_counter()
            nv = g.mesh_extract(0.0)
            tm.append(time.perf_counter() - t0m)
        t_mesh = min(tm)
        mesh = {"kernel": "k_mc_sweep (count + surface-cell list) + cub scan + k_mc_emit_list", "ms_per_extraction": t_mesh * 1e3, "triangles": nv // 3,
                "store_read_gbs": 8.0 * m ** 3 / t_mesh / 1e9, "peak": hbm, "frac": 8.0 * m ** 3 / t_mesh / 1e9 / hbm,
                "note": "host wall clock around tsdf_mesh_extract (ONE sweep over the 8 B/voxel store + scan + list emit enqueued behind it + one sync); "
                        "algorithmic bytes = 8 B x m^3 (every voxel read once); cub::DeviceScan is library code, off the frame path"}
    value = n_gpus * Ksteps / (ms_total * 1e-3)
    # stage breakdown: a SECOND pass over the same frames with the per-stage CUDA events on (an event record between
    # two kernels ends the programmatic-dependent-launch chain there, so the timed region above runs without them)
    g.reset(); g.set_intrinsics(K); g.set_pose(Rs[0], ts[0])
    g.enqueue_frame(dev, track=0, slot=0)
    for f in range(1, W):
        g.enqueue_frame(dev + f * frame_bytes, track=1, slot=f % ring)
    g.sync(); g.total_updates(reset=True)
    n_stage = min(Ksteps, 200)
    g.stage_timing_begin(n_stage)
    g.timer_begin()
    for f in range(W, W + n_stage):
        g.enqueue_frame(dev + f * frame_bytes, track=1, slot=f % ring)
    ms_staged = g.timer_end()
    g.sync()
    stage = g.stage_timing_end()
    n_upd_staged = g.total_updates()
    t_prep, t_track, t_fuse = [float(x) * 1e-3 for x in stage.mean(axis=0)]
    upd_per_frame = n_upd_staged / n_stage
    ach = 16.0 * upd_per_frame / t_fuse / 1e9
    n_valid = last_st["n_valid"]
    gather = 832.0 * n_valid * GN_ITERS / t_track / 1e9
    roof_fuse = {"bound": "hbm", "kernel": "fusion stage (k_fuse_tables + k_fuse_plan + k_fuse_cert + k_fuse_exact)", "achieved": ach, "peak": hbm,
                 "unit": "GB/s", "frac": ach / hbm, "traffic": ncu_traffic("fuse_trajectory"), "peak_source": peak_src, "ms_per_launch": t_fuse * 1e3,
                 "voxels_updated_per_launch": upd_per_frame, "voxel_updates_per_s": upd_per_frame / t_fuse,
                 "voxels_visited_per_s": m ** 3 / t_fuse,
                 "note": "algorithmic bytes = 16 B x voxels updated (read+write D,W); skipped voxels move no bytes. On this workload only ~5 % of the "
                         "134 M voxels are in view, so the stage is bound by certificate/fp64 latency on in-view voxels, not by HBM: see dense_fuse for the "
                         "HBM-bound case (every voxel updated), which is where north_star's >= 70 % target is defined"}
    # floor of one GN iteration.  Measured this round: with every gather forced into the same four lines ("perfect
    # memory", -DTSDF_X_HOTLOAD) the pixel loop loses only 1.9 us, and four alternative voxel layouts gain 9-11 %: the loop is bound by
    # instruction issue and dependent-instruction latency, not by the L1 data pipe as round 1 assumed.  Hard floor of the loop =
    # 7.11 M warp-instructions (ncu) / (148 SMs x 4 schedulers x 1 instruction per clock) = 6.1 us at 1965 MHz.  Serial tail: three
    # dependent global round trips (block partial + fence -> ticket -> read of the 148 partials, ~0.7 us each), a ~2 us chain of
    # dependent fp64 divisions (6x6 LU pivots, exp map, 3x3 inverse), the dependent launch and the reload of the pose (~1.6 us).
    floor_us = 6.1 + 2.1 + 2.0 + 1.6
    roof_track = {"bound": "instruction issue + dependent-latency chain", "kernel": "k_linearize", "achieved": gather, "peak": hbm, "unit": "GB/s",
                  "frac": gather / hbm, "ms_per_launch": t_track * 1e3 / GN_ITERS, "launches_per_frame": GN_ITERS,
                  "floor_us_per_launch": floor_us, "frac_of_floor": floor_us / (t_track * 1e6 / GN_ITERS),
                  "note": "832 B gathered per valid pixel-iteration (13 samples x 8 neighbours x {D,W}); the working set is L2-resident, so the HBM "
                          "peak (achieved/peak/frac) is only a yardstick; the bound that applies is floor_us_per_launch: issue slots of the pixel "
                          "loop (6.1 us for 7.11 M warp-instructions) + the serial reduction/solve/launch tail (5.7 us); frac_of_floor = floor / measured"}
    share = {"prep": t_prep, "track": t_track, "fuse": t_fuse}
    tot = sum(share.values())
    share = {k: v / tot for k, v in share.items()}
    g.dev_free(dev)
    color = None
    if rank == 0 and n_gpus == 1 and not args.no_color:
        color = color_fuse_bench(T, m, K, 5, hbm, depth[W], Rs[W], ts[W])

    # ---------------- e2e: host buffers through the public calls -------------------------------
    # (a) streaming: tsdf_submit_frame — every step does the H2D copy of its frame (pinned host
    #     memory -> device ring, on a copy stream) and the D2H of its pose record; frames are
    #     pipelined, results read after the final sync.  (b) synchronous: tsdf_track_and_fuse.
    g.reset(); g.set_intrinsics(K)
    g.set_pose(Rs[0], ts[0])
    g.submit_frame(depth[0], track=0, slot=0)
    for f in range(1, W):
        g.submit_frame(depth[f], track=1, slot=f % ring)
    g.sync()
    if dist is not None:
        import torch
        dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for f in range(W, n_frames):
        g.submit_frame(depth[f], track=1, slot=f % ring)
    g.sync()
    R_s, t_s, st_s = g.read_pose_ring((n_frames - 1) % ring)
    e2e_stream_s = time.perf_counter() - t0
    e2e_stream_s = barrier_max(dist, e2e_stream_s, tdev)

    g.reset(); g.set_intrinsics(K)
    g.fuse(depth[0], Rs[0], ts[0])
    for f in range(1, W):
        g.track_and_fuse(depth[f])
    if dist is not None:
        import torch
        dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for f in range(W, n_frames):
        R_e, t_e, st_e, nu_e = g.track_and_fuse(depth[f])
    e2e_s = time.perf_counter() - t0
    e2e_s = barrier_max(dist, e2e_s, tdev)
    e2e = {"value": n_gpus * Ksteps / e2e_stream_s, "unit": "frames/s", "h2d_bytes_per_step": int(frame_bytes),
           "d2h_bytes_per_step": 496,          # sizeof(PoseState): pose, twist, normal equations, counters
           "ms_per_step": 1e3 * e2e_stream_s / Ksteps,
           "timing": "host wall clock around K tsdf_submit_frame(HOST pinned depth) calls + final tsdf_sync and pose read; "
                     "H2D of frame n+1 overlaps track+fuse of frame n",
           "sync_value": n_gpus * Ksteps / e2e_s, "sync_ms_per_step": 1e3 * e2e_s / Ksteps,
           "sync_note": "K synchronous tsdf_track_and_fuse(HOST depth) calls: H2D, compute and D2H (504 B) serialised per frame",
           "stream_vs_sync_pose_diff_m": float(np.linalg.norm(t_s - t_e))}
    pose_agree = float(np.linalg.norm(t_e - last_t))
    track_err = float(np.linalg.norm(last_t - ts[n_frames - 1]))
    g.close()

    out = {"metric": METRIC if m == 512 else METRIC.replace("512", str(m)), "value": value, "unit": "frames/s", "n_gpus": n_gpus,
           "steps": Ksteps, "warmup": W, "ms_per_step": ms_total / Ksteps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32 values / f64 geometry", "data": "synthetic",
           "config": workload_config(m, n_gpus),
           "roofline": roof_fuse, "roofline_track": roof_track, "stage_share": share,
           "stage_ms": {"prep": t_prep * 1e3, "track": t_track * 1e3, "fuse": t_fuse * 1e3,
                        "measured_on": "a second pass over the first %d timed frames with per-stage CUDA events (%.4f ms/frame with them)" % (n_stage, ms_staged / n_stage)},
           "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
           "tracking": {"final_pos_err_vs_gt_m": track_err, "n_valid_last": int(n_valid), "e2e_vs_resident_pose_diff_m": pose_agree,
                        "ate_rmse_m": ate_aligned, "ate_rmse_unaligned_m": ate_raw, "ate_frames": int(len(poses_dev)),
                        "note": "ATE of the tracked poses vs the ground-truth path the frames were rendered from (tools/evaluate_ate.py); "
                                "drift is the reference algorithm's (frame-to-model tracking on a 1.2 cm grid), identical on the CPU oracle"}}

    if rank == 0 and n_gpus == 1 and not args.no_dense:
        out["dense_fuse"] = dense_fuse_bench(T, m, K, reps=20, hbm=hbm)
    if color is not None:
        out["color_fuse"] = color
    if mesh is not None:
        out["mesh"] = mesh
    if rank == 0 and n_gpus == 1 and not args.no_k0:
        out["preprocess_k0"] = k0_bench(T, m, K, 200, hbm)
        out["ragged_depth"] = ragged_bench(T, m, K, 200)
    if rank == 0 and n_gpus == 1 and not args.no_cpu:
        nb = min(n_frames, 40)
        out["cpu_baseline"] = run_cpu_baseline(depth[:nb], Rs[:nb], ts[:nb], m, budget_s=args.cpu_budget, max_frames=nb - 1)
    if not args.no_sharded:
        # BASELINE.json configs[2]: the SAME run also measures one 1024^3 volume z-slab sharded over the N ranks
        # (N = 1: the whole volume on one GPU — the 1-GPU point of the strong-scaling curve)
        sh = run_sharded(dist, tdev, device, n_gpus, args.sharded_grid, W, min(Ksteps, args.sharded_steps), False, hbm, peak_src)
        out["sharded"] = {k: sh[k] for k in ("metric", "value", "unit", "n_gpus", "steps", "ms_per_step", "scaling", "config", "roofline",
                                             "stage_ms", "e2e", "shard_check", "tracking")}
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(out)
    return 0


def run_sharded(dist, tdev, device, n_gpus, m, W, Ksteps, equal_slabs, hbm, peak_src):
    """BASELINE.json configs[2,3]: ONE m^3 volume cut into z-slabs, one slab per GPU (strong scaling: the frame
    sequence and the total voxel count are fixed as N grows).  Fusion needs no exchange (halos are fused
    redundantly); tracking exchanges 30 doubles per Gauss-Newton iteration in-kernel over NVLink peer stores.
    Returns the record (same on every rank)."""
    import tracking_sdf_b200 as T
    from tracking_sdf_b200 import sharding, capi
    from tools import synth
    K = synth.K_DEFAULT
    n_frames = W + Ksteps
    depth, Rs, ts = render_frames(n_frames, start=0, pinned=True)
    frame_bytes = depth[0].nbytes
    kw = dict(m=m, gauss_newton_max_iteration=GN_ITERS, maximum_twist_diff=float("-inf"))
    bounds = None
    halo = 0
    if dist is not None:
        halo = capi.slab_plan(capi.default_config(n_shards=n_gpus, shard_rank=min(1, n_gpus - 1), **kw))["halo"]
    if dist is not None and not equal_slabs:
        # work-balanced slabs: per-layer cost profile from 8 poses spread over the run (same frames on every rank,
        # so every rank derives the same partition), halo layers included in each slab's cost
        idx = list(range(0, n_frames, max(1, n_frames // 8)))
        # cost of a slab = what it fuses (halo included) + the tracked pixels it owns, in milliseconds of a frame
        # (single-GPU fusion time of this volume ~ 0.85 ms x (m/1024)^3; pixel loop of a frame's tracking 0.16 ms)
        wts, wown = sharding.frame_cost_weights(m, K, [(Rs[i], ts[i]) for i in idx], [depth[i] for i in idx],
                                                fuse_ms=0.85 * (m / 1024.0) ** 3, track_px_ms=0.16)
        bounds = capi.balanced_slabs(wts, n_gpus, min_layers=16, halo=halo, weights_own=wown)
    g = sharding.ShardedTsdf(dist, device, bounds=bounds, **kw) if dist is not None else T.Tsdf(T.default_config(device=device, **kw))
    g.set_intrinsics(K)
    ring = g.pose_ring_capacity()
    ks0, ks1, ko0, ko1 = g.stored_range()
    dev = g.dev_alloc(depth.nbytes); g.dev_upload(dev, depth)
    g.set_pose(Rs[0], ts[0])
    g.enqueue_frame(dev, track=0, slot=0)
    for f in range(1, W):
        g.enqueue_frame(dev + f * frame_bytes, track=1, slot=f % ring)
    g.sync(); g.total_updates(reset=True)
    launches0 = g.kernel_launch_count()
    sampler = ClockSampler(device); sampler.start()
    if dist is not None:
        import torch
        dist.barrier(); torch.cuda.synchronize()
    g.stage_timing_begin(Ksteps)
    g.timer_begin()
    for f in range(W, n_frames):
        g.enqueue_frame(dev + f * frame_bytes, track=1, slot=f % ring)
    ms_total = g.timer_end()
    g.sync()
    ms_total = barrier_max(dist, ms_total, tdev)
    clocks = sampler.stop()
    launches = g.kernel_launch_count() - launches0
    stage = g.stage_timing_end()
    n_upd_local = g.total_updates()
    R_l, t_l, st_l = g.read_pose_ring((n_frames - 1) % ring)      # raises on lost / halo / peer-timeout records
    t_prep, t_track, t_fuse = [float(x) * 1e-3 for x in stage.mean(axis=0)]
    t_fuse_max = barrier_max(dist, t_fuse, tdev)
    t_track_max = barrier_max(dist, t_track, tdev)
    n_upd = n_upd_local
    pose_spread = 0.0
    per_rank = None
    if dist is not None:
        # per-rank, per-frame stage times: the ranks run in lock step (every GN iteration waits for all of them), so a
        # frame lasts as long as its slowest fuser plus the tracking chain; which rank is slowest changes along the path
        import torch
        mine = torch.tensor(stage[:, 1:3].astype(np.float64), device=tdev)            # [frames, (track, fuse)]
        allr = [torch.empty_like(mine) for _ in range(n_gpus)]
        dist.all_gather(allr, mine)
        A = torch.stack(allr).cpu().numpy()                                           # [ranks, frames, 2]
        per_rank = {"track_ms_mean": [float(x) for x in A[:, :, 0].mean(axis=1)], "fuse_ms_mean": [float(x) for x in A[:, :, 1].mean(axis=1)],
                    "fuse_ms_max_over_ranks_mean_over_frames": float(A[:, :, 1].max(axis=0).mean()),
                    "track_ms_min_over_ranks_mean_over_frames": float(A[:, :, 0].min(axis=0).mean()),
                    "note": "track includes the wait for the slowest rank's fusion of the previous frame (the first GN iteration "
                            "cannot complete before every rank has joined); min over ranks = the tracking chain itself"}
    if dist is not None:
        import torch
        tt = torch.tensor([float(n_upd_local)], dtype=torch.float64, device=tdev)
        dist.all_reduce(tt); n_upd = float(tt.item())
        # every rank solved the same summed system: the final poses must be the same bits on all ranks
        pv = torch.tensor(np.concatenate([R_l.ravel(), t_l]), dtype=torch.float64, device=tdev)
        lo = pv.clone(); hi = pv.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        pose_spread = float((hi - lo).abs().max().item())
    # e2e: host buffers streamed through tsdf_submit_frame on every rank (H2D of frame n+1 overlaps frame n)
    g.reset(); g.set_intrinsics(K)
    g.set_pose(Rs[0], ts[0])
    g.submit_frame(depth[0], track=0, slot=0)
    for f in range(1, W):
        g.submit_frame(depth[f], track=1, slot=f % ring)
    g.sync()
    if dist is not None:
        import torch
        dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for f in range(W, n_frames):
        g.submit_frame(depth[f], track=1, slot=f % ring)
    g.sync()
    R_e, t_e, st_e = g.read_pose_ring((n_frames - 1) % ring)
    e2e_s = barrier_max(dist, time.perf_counter() - t0, tdev)
    upd_per_frame = n_upd / Ksteps
    ach = 16.0 * upd_per_frame / t_fuse_max / 1e9
    out = {"metric": METRIC.replace("512", str(m)) + " z-slab sharded", "value": Ksteps / (ms_total * 1e-3), "unit": "frames/s",
           "n_gpus": n_gpus, "steps": Ksteps, "warmup": W, "ms_per_step": ms_total / Ksteps, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f32 values / f64 geometry", "data": "synthetic",
           "config": {"workload": "%d^3 grid z-slab sharded over %d GPU(s), 640x480 synthetic depth along fr1/plant GT path, %d GN iterations/frame "
                                  "(BASELINE.json configs[2]/[3])" % (m, n_gpus, GN_ITERS),
                      "slab_rank0": {"own": [ko0, ko1], "stored": [ks0, ks1]}, "slab_bounds": bounds if bounds is not None else "equal thickness",
                      "halo_layers": halo, "grid_bytes_total": 8 * m ** 3,
                      "l2": "inputs larger than L2; no flush", "exchange": "30 doubles per GN iteration, in-kernel NVLink peer stores, rank-order sum"},
           "roofline": {"bound": "hbm", "kernel": "fusion stage (k_fuse_tables + k_fuse_plan + k_fuse_cert + k_fuse_exact)", "achieved": ach, "peak": hbm * n_gpus, "unit": "GB/s", "frac": ach / (hbm * n_gpus),
                        "traffic": None, "peak_source": peak_src + " x n_gpus", "ms_per_launch": t_fuse_max * 1e3,
                        "voxels_updated_per_launch": upd_per_frame},
           "stage_ms": {"prep": t_prep * 1e3, "track_max_over_ranks": t_track_max * 1e3, "fuse_max_over_ranks": t_fuse_max * 1e3,
                        "per_rank": per_rank},
           "e2e": {"value": Ksteps / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": int(frame_bytes) * n_gpus, "d2h_bytes_per_step": 496 * n_gpus,
                   "ms_per_step": 1e3 * e2e_s / Ksteps, "timing": "host wall clock, K tsdf_submit_frame(HOST pinned depth) calls + final tsdf_sync on every rank "
                                                                  "(every rank copies the whole frame: the depth image is replicated)"},
           "gpu_launches": int(launches), "clocks": clocks,
           "shard_check": {"pose_spread_across_ranks": pose_spread, "ok": bool(pose_spread == 0.0 and st_l["halo_miss"] == 0 and st_l["iterations"] == GN_ITERS),
                           "e2e_vs_resident_pose_diff_m": float(np.linalg.norm(t_e - t_l)),
                           "note": "every rank solves the same rank-order sum, so the final pose must be the same bits everywhere; slab contents vs the "
                                   "unsharded volume are compared by tools/shard_check.py (logs under profiles/)"},
           "tracking": {"final_pos_err_vs_gt_m": float(np.linalg.norm(t_l - ts[n_frames - 1])), "n_valid_last": int(st_l["n_valid"])}}
    g.dev_free(dev)
    g.close()
    return out


def main_sharded(args):
    rank, world, local = dist_env()
    n_gpus = max(world, 1)
    import tracking_sdf_b200 as T
    L = T.load_library()
    if L.tsdf_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device visible (the product has no CPU path)")
    device = local % L.tsdf_device_count()
    dist = init_dist(world, local, "nccl")
    tdev = None
    if dist is not None:
        import torch
        tdev = torch.device("cuda", device)
    hbm, peak_src = peaks()
    out = run_sharded(dist, tdev, device, n_gpus, args.m, args.warmup, args.steps, args.equal_slabs, hbm, peak_src)
    if dist is not None:
        dist.destroy_process_group()
    if rank == 0:
        emit(out)
    return 0


_REAL_STDOUT = None


def protect_stdout():
    """Libraries (NCCL's version banner) write to fd 1; the contract is ONE JSON line on stdout.
    Point fd 1 at stderr for the whole run and keep the real stdout for the final line."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(obj):
    _REAL_STDOUT.write(json.dumps(obj) + "\n")
    _REAL_STDOUT.flush()


def main():
    protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--grid", dest="m", type=int, default=None, help="voxels per axis (default 512; 1024 for --workload sharded)")
    ap.add_argument("--workload", default="sequence", choices=["sequence", "sharded"],
                    help="sequence: 512^3 per GPU, one independent sequence per GPU (default); sharded: one m^3 volume z-slab sharded over the GPUs")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-dense", action="store_true", help="skip the dense fusion micro-benchmark")
    ap.add_argument("--no-color", action="store_true", help="skip the colour fusion measurement")
    ap.add_argument("--no-mesh", action="store_true", help="skip the marching-cubes measurement")
    ap.add_argument("--no-k0", action="store_true", help="skip the noisy-depth K0 (pre-processing) measurement")
    ap.add_argument("--no-sharded", action="store_true", help="skip the z-slab sharded sub-record (1024^3 over the N ranks)")
    ap.add_argument("--sharded-grid", type=int, default=1024)
    ap.add_argument("--sharded-steps", type=int, default=200)
    ap.add_argument("--equal-slabs", action="store_true", help="sharded workload: equal-thickness z slabs instead of the work-balanced partition")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--ref-budget", type=float, default=90.0)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.m is None:
        args.m = 1024 if args.workload == "sharded" else 512
    if args.impl == "reference":
        return main_reference(args)
    if args.workload == "sharded":
        return main_sharded(args)
    return main_cuda(args)


if __name__ == "__main__":
    sys.exit(main())
