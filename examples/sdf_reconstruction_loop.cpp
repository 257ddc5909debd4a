/*
 * sdf_reconstruction_loop.cpp — the reference node's per-frame loop on the B200 library, in C++.
 *
 * Mirrors SDF_Reconstruction (sdf_reconstruction.cpp:82-91 construction, :21-80 kinect_callback,
 * :4-17 writePoseToFile) with b200::SDF / b200::CameraTracking from include/tracking_sdf_b200.hpp.
 * ROS, PCL and tf are replaced by the synthetic frame source (tools/synth.cpp) and the bundled
 * fr1/plant ground-truth path; everything between "frame arrives" and "pose written" is the
 * same call sequence the node makes.
 *
 *   build: g++ -O2 -std=c++17 -Iinclude examples/sdf_reconstruction_loop.cpp tools/synth.cpp \
 *              -Ltracking_sdf_b200/_lib -ltsdf_b200 -Wl,-rpath,$PWD/tracking_sdf_b200/_lib -fopenmp -o sdf_loop
 *   run:   ./sdf_loop data/fr1_plant_gt_every4.txt 20 256 trajectory.txt
 */
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iomanip>
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

/* sdf_reconstruction.cpp:4-17: `timestamp tx ty tz qx qy qz qw`, fixed, 4 decimals (TUM format) */
static void writePoseToFile(std::ofstream& f, double stamp, const std::array<double, 3>& t, const std::array<double, 9>& R) {
    double qw, qx, qy, qz;
    const double tr = R[0] + R[4] + R[8];
    if (tr > 0) { double s = std::sqrt(tr + 1.0) * 2; qw = 0.25 * s; qx = (R[7] - R[5]) / s; qy = (R[2] - R[6]) / s; qz = (R[3] - R[1]) / s; }
    else if (R[0] > R[4] && R[0] > R[8]) { double s = std::sqrt(1.0 + R[0] - R[4] - R[8]) * 2; qw = (R[7] - R[5]) / s; qx = 0.25 * s; qy = (R[1] + R[3]) / s; qz = (R[2] + R[6]) / s; }
    else if (R[4] > R[8]) { double s = std::sqrt(1.0 + R[4] - R[0] - R[8]) * 2; qw = (R[2] - R[6]) / s; qx = (R[1] + R[3]) / s; qy = 0.25 * s; qz = (R[5] + R[7]) / s; }
    else { double s = std::sqrt(1.0 + R[8] - R[0] - R[4]) * 2; qw = (R[3] - R[1]) / s; qx = (R[2] + R[6]) / s; qy = (R[5] + R[7]) / s; qz = 0.25 * s; }
    f << std::fixed << std::setprecision(4) << stamp << " " << t[0] << " " << t[1] << " " << t[2] << " " << qx << " " << qy << " " << qz << " " << qw << "\n";
}

int main(int argc, char** argv) {
    const char* traj = argc > 1 ? argv[1] : "data/fr1_plant_gt_every4.txt";
    const int n_frames = argc > 2 ? std::atoi(argv[2]) : 20;
    const int m = argc > 3 ? std::atoi(argv[3]) : 256;
    const char* out_path = argc > 4 ? argv[4] : "trajectory.txt";
    const char* mesh_path = argc > 5 && argv[5][0] != '-' ? argv[5] : nullptr;   /* optional: the visualisation thread's mesh as a PLY file ("-" = none) */
    const bool k0 = argc > 6 && std::atoi(argv[6]) != 0;         /* optional: 1 = the node's bilateral filter + normal estimation (:37-49) on the device */
    const std::vector<Pose> gt = load_trajectory(traj);
    if ((int)gt.size() < n_frames) { std::fprintf(stderr, "trajectory too short\n"); return 2; }
    const double K[9] = {525.0, 0, 319.5, 0, 525.0, 239.5, 0, 0, 1};
    const int W = 640, H = 480;
    try {
        /* sdf_reconstruction.cpp:83-88 */
        const double sdf_origin[3] = {-3.0, -3.0, -0.5};
        b200::SDF sdf(m, 6.0f, 6.0f, 3.5f, sdf_origin, 0.3f, 0.025f);
        b200::CameraTracking camera_tracking(20, 0.001f, 1.0f, 0.01f, &sdf, W, H, 0, k0);
        camera_tracking.camera_info_cb(K);                                       /* :90-91 */
        std::ofstream myfile(out_path, std::ios::out | std::ios::trunc);
        std::vector<float> depth((size_t)W * H);
        /* a registered colour image (the r,g,b of the XYZRGB cloud): a fixed test pattern */
        std::vector<uint8_t> rgb((size_t)W * H * 3);
        for (int v = 0; v < H; v++)
            for (int u = 0; u < W; u++) {
                uint8_t* c = &rgb[3 * ((size_t)v * W + u)];
                c[0] = (uint8_t)(255 * u / W); c[1] = (uint8_t)(255 * v / H); c[2] = (uint8_t)((((u / 32) + (v / 32)) & 1) ? 220 : 60);
            }
        int frame_num = 0;
        double err = 0;
        for (int f = 0; f < n_frames; f++) {                                     /* kinect_callback, :21-80 */
            synth_render_depth(gt[f].R, gt[f].t, K, W, H, depth.data());
            frame_num++;
            if (frame_num == 1) {
                camera_tracking.set_camera_transformation(gt[0].R, gt[0].t);     /* _useGroundTruth seed, :61-66 */
            } else {
                camera_tracking.estimate_new_position(&sdf, depth.data());       /* :70 */
                writePoseToFile(myfile, gt[f].stamp, camera_tracking.trans(), camera_tracking.rot());   /* :71 */
            }
            if (mesh_path) sdf.update(&camera_tracking, depth.data(), rgb.data());   /* :74 with the colour part, sdf.cpp:294-304 */
            else sdf.update(&camera_tracking, depth.data());                     /* :74 */
            const auto t = camera_tracking.trans();
            err = std::sqrt((t[0] - gt[f].t[0]) * (t[0] - gt[f].t[0]) + (t[1] - gt[f].t[1]) * (t[1] - gt[f].t[1]) + (t[2] - gt[f].t[2]) * (t[2] - gt[f].t[2]));
        }
        std::printf("frames %d  grid %d^3  final position error vs ground truth %.4f m  -> %s\n", n_frames, m, err, out_path);
        if (mesh_path) {
            /* SDF::visualize (sdf.cpp:317-391): marching cubes, marker points = vertices + sdf_origin, one
             * interpolate_color per vertex - here written as an ASCII PLY triangle soup instead of a ROS marker */
            std::vector<float> xyz, rgba;
            std::vector<double> world;
            const int64_t nv = sdf.mesh(xyz, &world, &rgba);
            std::ofstream ply(mesh_path, std::ios::out | std::ios::trunc);
            ply << "ply\nformat ascii 1.0\nelement vertex " << nv << "\nproperty float x\nproperty float y\nproperty float z\n"
                << "property uchar red\nproperty uchar green\nproperty uchar blue\nelement face " << nv / 3
                << "\nproperty list uchar int vertex_indices\nend_header\n";
            auto to8 = [](float c) { const float v = c * 255.0f; return (int)(v != v ? 0 : v < 0 ? 0 : v > 255 ? 255 : v); };   /* interpolated colours are 0..1 (sdf.cpp:210-213) */
            for (int64_t q = 0; q < nv; q++)
                ply << world[3 * q] << " " << world[3 * q + 1] << " " << world[3 * q + 2] << " " << to8(rgba[4 * q]) << " " << to8(rgba[4 * q + 1]) << " " << to8(rgba[4 * q + 2]) << "\n";
            for (int64_t q = 0; q < nv / 3; q++) ply << "3 " << 3 * q << " " << 3 * q + 1 << " " << 3 * q + 2 << "\n";
            std::printf("mesh: %lld triangles -> %s\n", (long long)(nv / 3), mesh_path);
        }
    } catch (const b200::Error& e) {
        std::fprintf(stderr, "tsdf_b200 error %d: %s\n", (int)e.status, e.what());
        return 1;
    }
    return 0;
}
