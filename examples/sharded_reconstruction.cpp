/*
 * sharded_reconstruction.cpp — the reference node's per-frame loop on a volume cut into z slabs, one slab per
 * shard, shards spread over every visible GPU, driven by ONE C++ host thread (BASELINE configs[2],[3]).
 *
 * The C++ counterpart of tracking_sdf_b200/sharding.py: b200::ShardedSDF (include/tracking_sdf_b200.hpp) creates the
 * shards, attaches their mailboxes (tsdf_shard_attach_local: in-kernel all-reduce of the 6x6 normal equations over
 * NVLink peer memory when the shards sit on different devices) and issues the node's call sequence
 * (sdf_reconstruction.cpp:61-74) to all of them in lock step.  With `verify` it also runs the same frames on one
 * unsharded volume and compares: every pose, and every OWNED layer of every slab against the unsharded D/W.
 *
 *   build: g++ -O2 -std=c++17 -Iinclude examples/sharded_reconstruction.cpp tools/synth.cpp \
 *              -Ltracking_sdf_b200/_lib -ltsdf_b200 -Wl,-rpath,$PWD/tracking_sdf_b200/_lib -fopenmp -o sharded_loop
 *   run:   ./sharded_loop data/fr1_plant_gt_every4.txt 12 256 4 1        (frames, m, shards, verify)
 */
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "tracking_sdf_b200.hpp"

extern "C" void synth_render_depth(const double R[9], const double t[3], const double K[9], int w, int h, float* depth);

struct Pose { double stamp, t[3], R[9]; };

static std::vector<Pose> load_trajectory(const char* path) {
    std::vector<Pose> out;
    std::ifstream f(path);
    std::string line;
    while (std::getline(f, line)) {
        if (line.empty() || line[0] == '#') continue;
        std::istringstream ss(line);
        Pose p; double q[4];
        ss >> p.stamp >> p.t[0] >> p.t[1] >> p.t[2] >> q[0] >> q[1] >> q[2] >> q[3];
        const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
        const double x = q[0] / n, y = q[1] / n, z = q[2] / n, w = q[3] / n;
        const double R[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                             2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                             2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)};
        for (int i = 0; i < 9; i++) p.R[i] = R[i];
        out.push_back(p);
    }
    return out;
}

int main(int argc, char** argv) {
    const char* traj = argc > 1 ? argv[1] : "data/fr1_plant_gt_every4.txt";
    const int n_frames = argc > 2 ? std::atoi(argv[2]) : 12;
    const int m = argc > 3 ? std::atoi(argv[3]) : 256;
    const int n_shards = argc > 4 ? std::atoi(argv[4]) : 2;
    const bool verify = argc > 5 && std::atoi(argv[5]) != 0;
    const std::vector<Pose> gt = load_trajectory(traj);
    if ((int)gt.size() < n_frames || n_frames < 2) { std::fprintf(stderr, "trajectory too short\n"); return 2; }
    const double K[9] = {525.0, 0, 319.5, 0, 525.0, 239.5, 0, 0, 1};
    const int W = 640, H = 480, GN = 10;
    try {
        const int ndev = tsdf_device_count();
        if (ndev < 1) throw b200::Error(TSDF_ERR_CUDA, "no CUDA device: tsdf_b200 has no CPU fallback");
        std::vector<int> devices;
        for (int d = 0; d < ndev && d < n_shards; d++) devices.push_back(d);
        const double origin[3] = {-3.0, -3.0, -0.5};
        /* fixed iteration count (maximum_twist_diff = -inf): sharded and unsharded runs execute the same launches */
        b200::ShardedSDF vol(n_shards, devices, m, 6.0f, 6.0f, 3.5f, origin, 0.3f, 0.025f, GN, -INFINITY, 1.0f, 0.01f, W, H);
        vol.camera_info_cb(K);
        std::vector<std::vector<float>> depth((size_t)n_frames, std::vector<float>((size_t)W * H));
        for (int f = 0; f < n_frames; f++) synth_render_depth(gt[f].R, gt[f].t, K, W, H, depth[(size_t)f].data());
        std::vector<std::array<double, 12>> poses((size_t)n_frames);
        vol.set_camera_transformation(gt[0].R, gt[0].t);
        vol.update(depth[0].data());
        const auto t0 = std::chrono::steady_clock::now();
        int64_t n_upd = 0;
        for (int f = 1; f < n_frames; f++) {
            double R[9], t[3];
            tsdf_track_stats st;
            n_upd = vol.estimate_new_position_and_update(depth[(size_t)f].data(), R, t, &st);
            if (st.halo_miss) { std::fprintf(stderr, "frame %d: a sample needed a voxel outside slab + halo\n", f); return 3; }
            for (int q = 0; q < 9; q++) poses[(size_t)f][(size_t)q] = R[q];
            for (int q = 0; q < 3; q++) poses[(size_t)f][(size_t)(9 + q)] = t[q];
        }
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        const auto& pl = poses[(size_t)n_frames - 1];
        const double err = std::sqrt((pl[9] - gt[n_frames - 1].t[0]) * (pl[9] - gt[n_frames - 1].t[0]) + (pl[10] - gt[n_frames - 1].t[1]) * (pl[10] - gt[n_frames - 1].t[1]) +
                                     (pl[11] - gt[n_frames - 1].t[2]) * (pl[11] - gt[n_frames - 1].t[2]));
        std::printf("shards %d on %d device(s)  grid %d^3  frames %d  %.1f frames/s (synchronous host loop)  last frame: %lld voxels updated, "
                    "position error vs ground truth %.4f m\n", n_shards, (int)devices.size(), m, n_frames, (n_frames - 1) / secs, (long long)n_upd, err);
        if (verify) {
            b200::SDF sdf(m, 6.0f, 6.0f, 3.5f, origin, 0.3f, 0.025f);
            b200::CameraTracking cam(GN, -INFINITY, 1.0f, 0.01f, &sdf, W, H, 0);
            cam.camera_info_cb(K);
            cam.set_camera_transformation(gt[0].R, gt[0].t);
            sdf.update(&cam, depth[0].data());
            double dpose = 0.0;
            for (int f = 1; f < n_frames; f++) {
                double R[9], t[3];
                cam.estimate_new_position_and_update(depth[(size_t)f].data(), R, t);
                for (int q = 0; q < 9; q++) dpose = std::fmax(dpose, std::fabs(R[q] - poses[(size_t)f][(size_t)q]));
                for (int q = 0; q < 3; q++) dpose = std::fmax(dpose, std::fabs(t[q] - poses[(size_t)f][(size_t)(9 + q)]));
            }
            /* owned layers of every slab against the unsharded volume (x-fastest layout: layer k is one m*m block) */
            const size_t layer = (size_t)m * m;
            std::vector<float> D((size_t)m * layer), Wt(D.size());
            b200::check(tsdf_download(sdf.handle(), D.data(), Wt.data(), TSDF_LAYOUT_XFASTEST));
            size_t differing = 0, compared = 0;
            double dmax = 0.0;
            for (int r = 0; r < vol.n_shards(); r++) {
                int32_t kb, ke, ob, oe;
                b200::check(tsdf_stored_range(vol.shard(r), &kb, &ke, &ob, &oe));
                std::vector<float> Ds((size_t)(ke - kb) * layer), Ws(Ds.size());
                b200::check(tsdf_download(vol.shard(r), Ds.data(), Ws.data(), TSDF_LAYOUT_XFASTEST));
                for (int k = ob; k < oe; k++)
                    for (size_t q = 0; q < layer; q++) {
                        const float a = Ds[(size_t)(k - kb) * layer + q], b = D[(size_t)k * layer + q];
                        const float wa = Ws[(size_t)(k - kb) * layer + q], wb = Wt[(size_t)k * layer + q];
                        compared++;
                        if (a != b || wa != wb) { differing++; dmax = std::fmax(dmax, std::fmax(std::fabs((double)a - b), std::fabs((double)wa - wb))); }
                    }
            }
            /* the normal equations are summed in a different order (slab by slab), so poses agree to rounding, not bits */
            std::printf("verify: max |pose difference| vs the unsharded volume %.3e; %zu of %zu owned voxels differ (max %.3e)\n", dpose, differing, compared, dmax);
            if (!(dpose < 1e-6) || !(dmax < 1e-3)) { std::fprintf(stderr, "sharded and unsharded runs disagree\n"); return 4; }
        }
    } catch (const b200::Error& e) {
        std::fprintf(stderr, "tsdf_b200 error %d: %s\n", (int)e.status, e.what());
        return 1;
    }
    return 0;
}
