/*
 * tracking_sdf_b200.hpp — header-only C++ mirror of the reference's two classes over the C ABI
 * (include/tsdf_b200.h).  Same class names, method names and argument meaning as
 * /root/reference/src/include/sdf_3d_reconstruction/{sdf.h,camera_tracking.h}, minus the
 * ROS/PCL/Eigen types: matrices are row-major double arrays, the organised point cloud +
 * normals arguments become the depth image they were computed from (back-projection and
 * normals are part of the device path).
 *
 *   reference                                                   here
 *   SDF(m, width, height, depth, origin, delta, eps)            b200::SDF(...)                  sdf.h:78-79
 *   CameraTracking(max_iter, max_twist_diff, v_h, w_h, sdf)     b200::CameraTracking(...)       camera_tracking.cpp:3-4
 *   camera_info_cb(msg)                                         camera_info_cb(K[9])            camera_tracking.cpp:22-36
 *   set_camera_transformation(rot, trans)                       set_camera_transformation(R,t)  camera_tracking.cpp:59-65
 *   estimate_new_position(sdf, cloud)                           estimate_new_position(sdf, depth)   camera_tracking.cpp:66-245
 *   sdf->update(camera_tracking, cloud, normals)                update(camera_tracking, depth)      sdf.cpp:224-315
 *   interpolate_distance(voxel_pt, is_interpolated)             interpolate_distance(...)       sdf.cpp:127-163
 *   sdf->update(...) colour part (cloud's r,g,b)                update(camera_tracking, depth, rgb) sdf.cpp:294-304
 *   interpolate_color(global_coords, color)                     interpolate_color(global, rgba) sdf.cpp:164-217
 *   mc->performReconstruction(cloud) + marker fill              mesh(xyz[, world, rgba])        marching_cubes_sdf.cpp:243-287, sdf.cpp:327-385
 *   get_array_index / get_voxel_coordinates / get_global_coordinates / get_number_of_voxels     sdf.h:107-157
 *   public rot, trans, rot_inv, rot_inv_trans, K, isKFilled     rot(), trans(), ... accessors
 *
 * Errors: the reference returns void, exit(0)s without intrinsics and lets NaN poses propagate;
 * here every failing call throws b200::Error carrying the tsdf_status and message.
 */
#ifndef TRACKING_SDF_B200_HPP_
#define TRACKING_SDF_B200_HPP_

#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "tsdf_b200.h"

namespace b200 {

struct Error : std::runtime_error {
    tsdf_status status;
    Error(tsdf_status s, const std::string& what) : std::runtime_error(what), status(s) {}
};
inline void check(tsdf_status s) {
    if (s != TSDF_OK) throw Error(s, tsdf_last_error());
}

class CameraTracking;

/* The voxel store.  Owns nothing until a CameraTracking is constructed on it: the reference's
 * two objects point at each other (sdf.cpp:245,250 call the tracker; camera_tracking.cpp:260,269
 * call the volume); here both are views of one device handle created from both constructors'
 * arguments. */
class SDF {
public:
    int m;
    float m_div_height, m_div_width, m_div_depth;             /* sdf.h:69-72 (fp32, sdf.cpp:19-21) */

    SDF(int m_, float width, float height, float depth, const double sdf_origin[3], float distance_delta,
        float distance_epsilon)
        : m(m_) {
        tsdf_default_config(&cfg_);
        cfg_.m = m_; cfg_.width = width; cfg_.height = height; cfg_.depth = depth;
        for (int q = 0; q < 3; q++) cfg_.origin[q] = sdf_origin[q];
        cfg_.distance_delta = distance_delta; cfg_.distance_epsilon = distance_epsilon;
        m_div_height = m_ / height; m_div_width = m_ / width; m_div_depth = m_ / depth;
    }
    ~SDF() { if (h_) tsdf_destroy(h_); }
    SDF(const SDF&) = delete;
    SDF& operator=(const SDF&) = delete;

    int64_t get_number_of_voxels() const { return (int64_t)m * m * m; }                         /* sdf.h:107 */
    int64_t get_array_index(const int32_t ijk[3]) const { return tsdf_get_array_index(handle(), ijk[0], ijk[1], ijk[2]); }
    void get_voxel_coordinates(int64_t array_idx, int32_t ijk[3]) const { tsdf_get_voxel_coordinates_idx(handle(), array_idx, ijk); }
    void get_voxel_coordinates(const double global[3], double voxel[3]) const { tsdf_get_voxel_coordinates(handle(), global, voxel); }
    void get_global_coordinates(const int32_t ijk[3], double global[3]) const { tsdf_get_global_coordinates(handle(), ijk, global); }

    /* sdf.cpp:127-163 for one point in continuous voxel coordinates */
    float interpolate_distance(const double voxel_coordinates[3], bool& is_interpolated) const {
        float v; uint8_t ok;
        check(tsdf_interpolate_distance(handle(), 1, voxel_coordinates, &v, &ok));
        is_interpolated = ok != 0;
        return v;
    }
    /* batched form: pts n x 3 */
    void interpolate_distance(int64_t n, const double* pts, float* out, uint8_t* ok) const {
        check(tsdf_interpolate_distance(handle(), n, pts, out, ok));
    }

    /* sdf.cpp:224-315: integrate `depth` (host, float32 metres, row-major) at the tracker's current pose */
    inline int64_t update(CameraTracking* camera_tracking, const float* depth);

    /* the same with the colour running mean (sdf.cpp:294-304); rgb = height*width*3 bytes, the r,g,b of the
     * organised XYZRGB cloud the reference receives */
    inline int64_t update(CameraTracking* camera_tracking, const float* depth, const uint8_t* rgb);

    /* sdf.cpp:164-217: colour at a WORLD point; rgba[3] = 1 */
    void interpolate_color(const double global_coords[3], float rgba[4]) const {
        check(tsdf_interpolate_color(handle(), 1, global_coords, rgba));
    }
    void interpolate_color(int64_t n, const double* global_pts, float* rgba) const {
        check(tsdf_interpolate_color(handle(), n, global_pts, rgba));
    }
    /* Color_W, R, G, B (sdf.h:49-52), reference layout */
    void download_color(std::vector<float>& CW, std::vector<float>& R, std::vector<float>& G, std::vector<float>& B) const {
        const size_t n = (size_t)get_number_of_voxels();
        CW.resize(n); R.resize(n); G.resize(n); B.resize(n);
        check(tsdf_download_color(handle(), CW.data(), R.data(), G.data(), B.data(), TSDF_LAYOUT_REFERENCE));
    }

    /* the visualisation thread's product (sdf.cpp:327-385) without the 1-64 GiB download: marching cubes on the
     * device (iso level 0, sdf.cpp:44), triangle soup in the reference's order; world = marker points
     * (+ sdf_origin), rgba = interpolate_color per vertex (needs colour fusion to have run) */
    int64_t mesh(std::vector<float>& xyz, std::vector<double>* world = nullptr, std::vector<float>* rgba = nullptr, float iso_level = 0.0f) const {
        int64_t n = 0;
        check(tsdf_mesh_extract(handle(), iso_level, &n));
        xyz.resize((size_t)n * 3);
        if (world) world->resize((size_t)n * 3);
        if (rgba) rgba->resize((size_t)n * 4);
        check(tsdf_mesh_download(handle(), xyz.data(), world ? world->data() : nullptr, rgba ? rgba->data() : nullptr));
        return n;
    }

    /* the raw arrays the reference hands to its mesher (sdf.cpp:47-48), reference (z-fastest) layout */
    void download(std::vector<float>& D, std::vector<float>& W) const {
        D.resize((size_t)get_number_of_voxels()); W.resize(D.size());
        check(tsdf_download(handle(), D.data(), W.data(), TSDF_LAYOUT_REFERENCE));
    }
    tsdf_handle handle() const {
        if (!h_) throw Error(TSDF_ERR_BAD_ARG, "SDF is not attached to a CameraTracking yet");
        return h_;
    }

private:
    friend class CameraTracking;
    tsdf_config cfg_;
    tsdf_handle h_ = nullptr;
};

class CameraTracking {
public:
    bool isKFilled = false;                                   /* camera_tracking.h:59 */

    /* camera_tracking.cpp:3-4 — NB the definition's parameter order (max_iter, max_twist_diff, v_h, w_h) */
    CameraTracking(int gauss_newton_max_iteration, float maximum_twist_diff, float v_h, float w_h, SDF* sdf,
                   int image_width = 640, int image_height = 480, int device = 0,
                   bool preprocess = false /* the node's bilateral filter + normal estimation, sdf_reconstruction.cpp:37-49, on the device (K0) */)
        : sdf_(sdf) {
        tsdf_config c = sdf->cfg_;
        c.gauss_newton_max_iteration = gauss_newton_max_iteration;
        c.maximum_twist_diff = maximum_twist_diff;
        c.v_h = v_h; c.w_h = w_h;
        c.image_width = image_width; c.image_height = image_height; c.device = device;
        c.preprocess = preprocess ? 1 : 0;
        /* one device volume per SDF: a second tracker on the same SDF would orphan the first handle */
        if (sdf->h_) throw Error(TSDF_ERR_BAD_ARG, "this SDF is already attached to a CameraTracking");
        check(tsdf_create(&c, &sdf->h_));
        h_ = sdf->h_;
    }

    void camera_info_cb(const double K_row_major[9]) {        /* camera_tracking.cpp:22-36 */
        check(tsdf_set_intrinsics(h_, K_row_major));
        for (int q = 0; q < 9; q++) K[q] = K_row_major[q];
        isKFilled = true;
    }
    void set_camera_transformation(const double rot[9], const double trans[3]) { check(tsdf_set_pose(h_, rot, trans)); }

    /* camera_tracking.cpp:66-245; the new pose is also returned through rot()/trans() like the
     * reference's public members */
    void estimate_new_position(const SDF*, const float* depth, tsdf_track_stats* stats = nullptr) {
        double R[9], t[3];
        check(tsdf_track(h_, depth, TSDF_HOST, R, t, stats));
    }
    /* kinect_callback's track-then-update in one device pass (sdf_reconstruction.cpp:69-74) */
    void estimate_new_position_and_update(const float* depth, double R_out[9], double t_out[3],
                                          tsdf_track_stats* stats = nullptr, int64_t* n_updated = nullptr) {
        check(tsdf_track_and_fuse(h_, depth, TSDF_HOST, R_out, t_out, stats, n_updated));
    }

    std::array<double, 9> rot() const { std::array<double, 9> R; double t[3]; check(tsdf_get_pose(h_, R.data(), t)); return R; }
    std::array<double, 3> trans() const { double R[9]; std::array<double, 3> t; check(tsdf_get_pose(h_, R, t.data())); return t; }
    std::array<double, 9> rot_inv() const { std::array<double, 9> R; double t[3]; check(tsdf_get_pose_inv(h_, R.data(), t)); return R; }
    std::array<double, 3> rot_inv_trans() const { double R[9]; std::array<double, 3> t; check(tsdf_get_pose_inv(h_, R, t.data())); return t; }

    /* camera_tracking.cpp:40-58 (host-side conveniences, same formulas) */
    void project_camera_to_image_plane(const double cam[3], double img[2]) const {
        const double i0 = K[0] * cam[0] + K[1] * cam[1] + K[2] * cam[2], i1 = K[3] * cam[0] + K[4] * cam[1] + K[5] * cam[2],
                     i2 = K[6] * cam[0] + K[7] * cam[1] + K[8] * cam[2];
        img[0] = i0 / i2; img[1] = i1 / i2;
    }
    void project_world_to_camera(const double w[3], double c[3]) const {
        const auto Ri = rot_inv(); const auto ti = rot_inv_trans();
        for (int r = 0; r < 3; r++) c[r] = (Ri[3 * r] * w[0] + Ri[3 * r + 1] * w[1]) + Ri[3 * r + 2] * w[2] + ti[r];
    }
    void project_camera_to_world(const double c[3], double w[3]) const {
        const auto R = rot(); const auto t = trans();
        for (int r = 0; r < 3; r++) w[r] = (R[3 * r] * c[0] + R[3 * r + 1] * c[1]) + R[3 * r + 2] * c[2] + t[r];
    }

    double K[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    tsdf_handle handle() const { return h_; }

private:
    SDF* sdf_;
    tsdf_handle h_ = nullptr;
};

inline int64_t SDF::update(CameraTracking* camera_tracking, const float* depth) {
    if (!camera_tracking->isKFilled) throw Error(TSDF_ERR_NO_INTRINSICS, "Camera Matrix not received");   /* sdf.cpp:227-229 */
    int64_t n = 0;
    check(tsdf_fuse(handle(), depth, TSDF_HOST, nullptr, nullptr, &n));
    return n;
}

inline int64_t SDF::update(CameraTracking* camera_tracking, const float* depth, const uint8_t* rgb) {
    if (!camera_tracking->isKFilled) throw Error(TSDF_ERR_NO_INTRINSICS, "Camera Matrix not received");   /* sdf.cpp:227-229 */
    int64_t n = 0;
    check(tsdf_fuse_rgb(handle(), depth, rgb, TSDF_HOST, nullptr, nullptr, &n));
    return n;
}

/* One volume cut into z slabs, one slab per shard, driven from ONE host thread (the C++ counterpart of
 * tracking_sdf_b200/sharding.py, which does the same across processes with torch.distributed for the rendezvous).
 * Shards are placed round-robin on `devices`; shards on different devices all-reduce the 6x6 normal equations inside
 * the tracking kernel over NVLink peer memory (tsdf_shard_attach_local), shards sharing a device (a single-GPU box)
 * use the deferred same-device combine.  The per-frame call sequence is the node's (sdf_reconstruction.cpp:61-74):
 * set_camera_transformation on the first frame, estimate_new_position_and_update afterwards.  Every shard holds the
 * same pose; n_updated sums the voxels each shard owns.
 *   cost_weights: optional per-z-layer fusion cost (m values) for work-balanced slabs (tsdf_balanced_slabs); nullptr =
 *   equal thickness. */
class ShardedSDF {
public:
    ShardedSDF(int n_shards, const std::vector<int>& devices, int m, float width, float height, float depth,
               const double sdf_origin[3], float distance_delta, float distance_epsilon,
               int gauss_newton_max_iteration, float maximum_twist_diff, float v_h, float w_h,
               int image_width = 640, int image_height = 480, const double* cost_weights = nullptr) {
        if (n_shards < 1 || devices.empty()) throw Error(TSDF_ERR_BAD_ARG, "ShardedSDF: need at least one shard and one device");
        tsdf_config c;
        tsdf_default_config(&c);
        c.m = m; c.width = width; c.height = height; c.depth = depth;
        for (int q = 0; q < 3; q++) c.origin[q] = sdf_origin[q];
        c.distance_delta = distance_delta; c.distance_epsilon = distance_epsilon;
        c.gauss_newton_max_iteration = gauss_newton_max_iteration; c.maximum_twist_diff = maximum_twist_diff;
        c.v_h = v_h; c.w_h = w_h; c.image_width = image_width; c.image_height = image_height;
        c.n_shards = n_shards;
        std::vector<int32_t> bounds;
        if (cost_weights && n_shards > 1) {
            int32_t plan[5];
            c.shard_rank = 0;
            check(tsdf_slab_plan(&c, plan));
            bounds.resize((size_t)n_shards + 1);
            check(tsdf_balanced_slabs(m, n_shards, cost_weights, 8, plan[4], bounds.data()));
        }
        try {
            for (int r = 0; r < n_shards; r++) {
                c.shard_rank = r;
                c.device = devices[(size_t)r % devices.size()];
                if (!bounds.empty()) { c.slab_k_begin = bounds[(size_t)r]; c.slab_k_end = bounds[(size_t)r + 1]; }
                tsdf_handle h = nullptr;
                check(tsdf_create(&c, &h));
                h_.push_back(h);
            }
            if (n_shards > 1) check(tsdf_shard_attach_local(h_.data(), n_shards));
        } catch (...) {
            for (tsdf_handle h : h_) tsdf_destroy(h);
            throw;
        }
    }
    ~ShardedSDF() { for (tsdf_handle h : h_) tsdf_destroy(h); }
    ShardedSDF(const ShardedSDF&) = delete;
    ShardedSDF& operator=(const ShardedSDF&) = delete;

    int n_shards() const { return (int)h_.size(); }
    void camera_info_cb(const double K_row_major[9]) { check(tsdf_group_set_intrinsics(h_.data(), n_shards(), K_row_major)); }
    void set_camera_transformation(const double rot[9], const double trans[3]) { check(tsdf_group_set_pose(h_.data(), n_shards(), rot, trans)); }
    /* sdf->update at the current pose (frame 1 of the node) */
    int64_t update(const float* depth, double R_out[9] = nullptr, double t_out[3] = nullptr) {
        int64_t n = 0;
        check(tsdf_group_frame(h_.data(), n_shards(), depth, TSDF_HOST, 0, 1, R_out, t_out, nullptr, &n));
        return n;
    }
    /* estimate_new_position + update (sdf_reconstruction.cpp:69-74) on every shard in lock step */
    int64_t estimate_new_position_and_update(const float* depth, double R_out[9], double t_out[3], tsdf_track_stats* stats = nullptr) {
        int64_t n = 0;
        check(tsdf_group_frame(h_.data(), n_shards(), depth, TSDF_HOST, 1, 1, R_out, t_out, stats, &n));
        return n;
    }
    /* owned z layers [begin, end) of shard r */
    void owned_layers(int r, int32_t& begin, int32_t& end) const {
        int32_t kb, ke;
        check(tsdf_stored_range(h_.at((size_t)r), &kb, &ke, &begin, &end));
    }
    /* the slab of shard r in the reference's layout (its STORED layers, halo included): see tsdf_download */
    tsdf_handle shard(int r) const { return h_.at((size_t)r); }

private:
    std::vector<tsdf_handle> h_;
};

}  // namespace b200
#endif
