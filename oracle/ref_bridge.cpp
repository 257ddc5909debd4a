/*
 * ref_bridge.cpp — C ABI over the REFERENCE ITSELF (test infrastructure; same rule as oracle.h).
 *
 * Linked with the four hot-path translation units of the reference, compiled UNMODIFIED from where they lie
 * (/root/reference/src/src/{sdf,camera_tracking,eigen_utils,marching_cubes_sdf}.cpp) against the shim headers
 * in oracle/shim/ (Eigen subset, PCL containers, ROS stubs — the image has none of the real ones).  The
 * result, oracle/_ref/libtsdf_ref.so, is what pins oracle/oracle.cpp: tests/test_oracle_vs_ref.py compares the
 * two function by function, tests/golden/ is generated from it, and bench.py --impl reference times it.
 *
 * This file contains no algorithm: it only builds the objects the node builds
 * (sdf_reconstruction.cpp:83-88), converts flat arrays to pcl::PointCloud, calls the reference's own member
 * functions and copies results out.  `private`/`protected` are redefined for THIS translation unit only so the
 * bridge can read D, W, A, b, ... (class layout does not depend on access specifiers with g++); the reference's
 * own translation units are compiled without any define.
 */
#include <algorithm>
#include <chrono>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>
#include <omp.h>

#include "oracle.h"

#define private public
#define protected public
#include "sdf_3d_reconstruction/sdf.h"
#include "sdf_3d_reconstruction/camera_tracking.h"
#include "sdf_3d_reconstruction/eigen_utils.h"
#undef private
#undef protected

namespace {

struct Ref {
    orc_config cfg;
    SDF* sdf;
    CameraTracking* tracker;
    pcl::PointCloud<pcl::PointXYZRGB>::Ptr cloud;
    pcl::PointCloud<pcl::Normal>::Ptr normals;
    pcl::PointCloud<pcl::PointXYZ> mesh;
    std::vector<float> w_before;
    double t_track = 0, t_update = 0;          /* seconds inside the reference's own calls (conversions excluded) */
};

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

/* the reference reports through std::cout; capture it so tests stay quiet and the GN stop step can be read */
struct CoutCapture {
    std::ostringstream ss;
    std::streambuf* old;
    CoutCapture() : old(std::cout.rdbuf(ss.rdbuf())) {}
    ~CoutCapture() { std::cout.rdbuf(old); }
};

void fill_clouds(Ref* r, const float* cloud, const float* normals, const uint8_t* rgb) {
    const int Wd = r->cfg.image_width, Hd = r->cfg.image_height;
    const size_t n = (size_t)Wd * Hd;
    if (!r->cloud) r->cloud = boost::make_shared<pcl::PointCloud<pcl::PointXYZRGB> >();
    if (!r->normals) r->normals = boost::make_shared<pcl::PointCloud<pcl::Normal> >();
    r->cloud->points.resize(n); r->cloud->width = Wd; r->cloud->height = Hd;
    r->normals->points.resize(n); r->normals->width = Wd; r->normals->height = Hd;
#pragma omp parallel for
    for (int64_t p = 0; p < (int64_t)n; p++) {
        pcl::PointXYZRGB& q = r->cloud->points[p];
        q.x = cloud[3 * p]; q.y = cloud[3 * p + 1]; q.z = cloud[3 * p + 2];
        if (rgb) { q.r = rgb[3 * p]; q.g = rgb[3 * p + 1]; q.b = rgb[3 * p + 2]; } else { q.r = q.g = q.b = 0; }
        pcl::Normal& m = r->normals->points[p];
        if (normals) { m.normal_x = normals[3 * p]; m.normal_y = normals[3 * p + 1]; m.normal_z = normals[3 * p + 2]; }
        else { m.normal_x = m.normal_y = m.normal_z = std::nanf(""); }
    }
}

void get_rot(const Eigen::Matrix3d& M, double out[9]) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) out[3 * i + j] = M(i, j);
}

}  // namespace

extern "C" {

/* sdf_reconstruction.cpp:83-88: new SDF(m, width, height, depth, origin, delta, epsilon);
 * new CameraTracking(max_iter, max_twist_diff, v_h, w_h, sdf) — definition order, camera_tracking.cpp:3-4 */
void* ref_create(const orc_config* cfg) {
    CoutCapture cap;
    Ref* r = new Ref();
    r->cfg = *cfg;
    Eigen::Vector3d origin(cfg->origin[0], cfg->origin[1], cfg->origin[2]);
    r->sdf = new SDF(cfg->m, cfg->width, cfg->height, cfg->depth, origin, cfg->distance_delta, cfg->distance_epsilon);
    r->tracker = new CameraTracking(cfg->gauss_newton_max_iteration, cfg->maximum_twist_diff, cfg->v_h, cfg->w_h, r->sdf);
    r->tracker->isKFilled = false;          /* the reference leaves it uninitialised until camera_info_cb */
    return r;
}
void ref_destroy(void* h) {
    Ref* r = (Ref*)h;
    /* ~SDF is empty in the reference (sdf.cpp:52-54): release its arrays here */
    delete[] r->sdf->D; delete[] r->sdf->W; delete[] r->sdf->global_coords; delete[] r->sdf->voxel_coords;
    delete[] r->sdf->Color_W; delete[] r->sdf->R; delete[] r->sdf->G; delete[] r->sdf->B;
    delete r->sdf->mc;
    delete r->tracker;
    delete r->sdf;
    delete r;
}

/* camera_tracking.cpp:22-36 through the reference's own callback */
void ref_set_intrinsics(void* h, const double K[9]) {
    Ref* r = (Ref*)h;
    CoutCapture cap;
    auto msg = boost::make_shared<sensor_msgs::CameraInfo>();
    for (int q = 0; q < 9; q++) msg->K[q] = K[q];
    r->tracker->camera_info_cb(msg);
}
void ref_set_pose(void* h, const double R[9], const double t[3]) {
    Ref* r = (Ref*)h;
    Eigen::Matrix3d rot;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) rot(i, j) = R[3 * i + j];
    Eigen::Vector3d trans(t[0], t[1], t[2]);
    r->tracker->set_camera_transformation(rot, trans);
}
void ref_get_pose(void* h, double R[9], double t[3]) {
    Ref* r = (Ref*)h;
    get_rot(r->tracker->rot, R);
    for (int q = 0; q < 3; q++) t[q] = r->tracker->trans(q);
}
void ref_get_pose_inv(void* h, double Rinv[9], double tinv[3]) {
    Ref* r = (Ref*)h;
    get_rot(r->tracker->rot_inv, Rinv);
    for (int q = 0; q < 3; q++) tinv[q] = r->tracker->rot_inv_trans(q);
}
void ref_set_gn(void* h, int max_iter, float max_twist_diff) {
    Ref* r = (Ref*)h;
    r->tracker->gauss_newton_max_iteration = max_iter;
    r->tracker->maximum_twist_diff = max_twist_diff;
}

/* SDF::update (sdf.cpp:224-315) on a cloud + normals (+ rgb).  count != 0: return the number of voxels whose
 * W changed (the reference keeps no counter); count == 0: return 0 (timed runs). */
int64_t ref_fuse_cloud(void* h, const float* cloud, const float* normals, const uint8_t* rgb, int count) {
    Ref* r = (Ref*)h;
    if (!r->tracker->isKFilled) return -1;                  /* the reference would exit(0), sdf.cpp:227-229 */
    fill_clouds(r, cloud, normals, rgb);
    const int64_t nv = r->sdf->number_of_voxels;
    if (count) r->w_before.assign(r->sdf->W, r->sdf->W + nv);
    {
        CoutCapture cap;
        const double t0 = now_s();
        r->sdf->update(r->tracker, r->cloud, r->normals);
        r->t_update += now_s() - t0;
    }
    int64_t n = 0;
    if (count)
        for (int64_t i = 0; i < nv; i++) n += (r->sdf->W[i] != r->w_before[i]);
    return n;
}

/* CameraTracking::estimate_new_position (camera_tracking.cpp:66-245) */
void ref_track_cloud(void* h, const float* cloud, orc_track_stats* st) {
    Ref* r = (Ref*)h;
    fill_clouds(r, cloud, nullptr, nullptr);
    std::string log;
    {
        CoutCapture cap;
        const double t0 = now_s();
        r->tracker->estimate_new_position(r->sdf, r->cloud);
        r->t_track += now_s() - t0;
        log = cap.ss.str();
    }
    if (st) {
        memset(st, 0, sizeof *st);
        const char* key = "STOP Gauss Newton at step: ";
        size_t p = log.find(key);
        if (p != std::string::npos) { st->stopped = 1; st->iterations = atoi(log.c_str() + p + strlen(key)) + 1; }
        else st->iterations = r->tracker->gauss_newton_max_iteration;
        bool finite = true;
        for (int q = 0; q < 3; q++) finite = finite && std::isfinite(r->tracker->trans(q));
        st->singular = finite ? 0 : 1;
    }
}

/* One GN iteration's normal equations as the reference builds them (camera_tracking.cpp:146-189): run
 * estimate_new_position with max_iter = 1, read the members A and b, restore the pose.  After the call the
 * perturbed rotations r1p..r3m belong to the restored pose, so get_partial_derivative can be called per pixel. */
void ref_linearize_cloud(void* h, const float* cloud, double A[36], double b[6]) {
    Ref* r = (Ref*)h;
    fill_clouds(r, cloud, nullptr, nullptr);
    Eigen::Matrix3d rot = r->tracker->rot;
    Eigen::Vector3d trans = r->tracker->trans;
    const int it = r->tracker->gauss_newton_max_iteration;
    r->tracker->gauss_newton_max_iteration = 1;
    {
        CoutCapture cap;
        r->tracker->estimate_new_position(r->sdf, r->cloud);
    }
    r->tracker->gauss_newton_max_iteration = it;
    r->tracker->set_camera_transformation(rot, trans);
    for (int i = 0; i < 6; i++) {
        for (int j = 0; j < 6; j++) A[6 * i + j] = r->tracker->A(i, j);
        b[i] = r->tracker->b(i);
    }
}

/* the solve + exp map + pose update of one iteration, by the reference's own expressions
 * (camera_tracking.cpp:191-192, 237-239) — these four lines are the only place the bridge restates code,
 * because the reference has no function boundary there. */
void ref_apply_update(void* h, const double A[36], const double b[6], double twist_out[6]) {
    Ref* r = (Ref*)h;
    CameraTracking* c = r->tracker;
    for (int i = 0; i < 6; i++) {
        for (int j = 0; j < 6; j++) c->A(i, j) = A[6 * i + j];
        c->b(i) = b[i];
    }
    c->twist_diff = c->A.inverse() * c->b;
    Eigen::Affine3d aff = eigen_utils::direct_exponential_map(c->twist_diff, 1.0);
    c->rot = aff.rotation().transpose() * c->rot;
    c->trans = c->trans - aff.rotation().transpose() * aff.translation();
    c->set_camera_transformation(c->rot, c->trans);
    for (int q = 0; q < 6; q++) twist_out[q] = c->twist_diff(q);
}

/* per strided pixel, in the reference's loop order (i outer, j inner, camera_tracking.cpp:162-163):
 * CameraTracking::get_partial_derivative called with FRESH per-pixel state.  flag: 0 NaN point, 1 ok,
 * 2 centre out of the volume (the early return left is_interpolated untouched), 3 a sample not interpolated. */
int32_t ref_linearize_pixels(void* h, const float* cloud, float* J, float* psi, uint8_t* flag) {
    Ref* r = (Ref*)h;
    double A[36], b[6];
    ref_linearize_cloud(h, cloud, A, b);                    /* sets r1p..r3m for the current pose */
    const int Wd = r->cfg.image_width, Hd = r->cfg.image_height, s = 3;   /* the stride is hard-coded, :162-163 */
    const int nj = (Hd + s - 1) / s;
    int n = 0;
    for (int i = 0; i < Wd; i += s)
        for (int j = 0; j < Hd; j += s, n++) {
            const int p = (i / s) * nj + (j / s);
            for (int a = 0; a < 6; a++) J[6 * p + a] = 0;
            psi[p] = 0; flag[p] = 0;
            pcl::PointXYZRGB point = r->cloud->at(i, j);
            if (std::isnan(point.x) || std::isnan(point.y) || std::isnan(point.z)) continue;
            Eigen::Vector3d cp(point.x, point.y, point.z);
            Eigen::Matrix<double, 6, 1> d = Eigen::Matrix<double, 6, 1>::Zero();
            bool is_interpolated = false;
            double sdf_val = 0;
            Eigen::Vector3d world, vox;
            r->tracker->project_camera_to_world(cp, world);
            r->sdf->get_voxel_coordinates(world, vox);
            const int m = r->sdf->m;
            const bool oob = vox(0) < 0 || vox(1) < 0 || vox(2) < 0 || vox(0) >= m || vox(1) >= m || vox(2) >= m;
            r->tracker->get_partial_derivative(r->sdf, cp, d, is_interpolated, sdf_val);
            if (oob) { flag[p] = 2; continue; }
            if (!is_interpolated) { flag[p] = 3; continue; }
            flag[p] = 1;
            for (int a = 0; a < 6; a++) J[6 * p + a] = (float)d(a);
            psi[p] = (float)sdf_val;
        }
    return n;
}

/* SDF::interpolate_distance (sdf.cpp:127-163) */
void ref_interpolate(void* h, int64_t n, const double* pts, float* out, uint8_t* ok) {
    Ref* r = (Ref*)h;
    for (int64_t q = 0; q < n; q++) {
        Eigen::Vector3d v(pts[3 * q], pts[3 * q + 1], pts[3 * q + 2]);
        bool is_interp = false;
        out[q] = r->sdf->interpolate_distance(v, is_interp);
        ok[q] = is_interp ? 1 : 0;
    }
}

/* eigen_utils::direct_exponential_map (eigen_utils.cpp:85-128), delta_t = 1 */
void ref_exp_map(const double twist[6], double R[9], double t[3]) {
    Eigen::Matrix<double, 6, 1> v;
    for (int q = 0; q < 6; q++) v(q) = twist[q];
    Eigen::Affine3d aff = eigen_utils::direct_exponential_map(v, 1.0);
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) R[3 * i + j] = aff(i, j);
        t[i] = aff(i, 3);
    }
}

/* sdf.h:113-157 */
int64_t ref_get_array_index(void* h, int32_t i, int32_t j, int32_t k) {
    Eigen::Vector3i v(i, j, k);
    return ((Ref*)h)->sdf->get_array_index(v);
}
void ref_get_voxel_coordinates_idx(void* h, int64_t idx, int32_t ijk[3]) {
    Eigen::Vector3i v;
    ((Ref*)h)->sdf->get_voxel_coordinates((int)idx, v);
    for (int q = 0; q < 3; q++) ijk[q] = v(q);
}
void ref_get_voxel_coordinates(void* h, const double g[3], double v[3]) {
    Eigen::Vector3d gg(g[0], g[1], g[2]), vv;
    ((Ref*)h)->sdf->get_voxel_coordinates(gg, vv);
    for (int q = 0; q < 3; q++) v[q] = vv(q);
}
void ref_get_global_coordinates(void* h, const int32_t ijk[3], double g[3]) {
    Eigen::Vector3i v(ijk[0], ijk[1], ijk[2]);
    Eigen::Vector3d gg;
    ((Ref*)h)->sdf->get_global_coordinates(v, gg);
    for (int q = 0; q < 3; q++) g[q] = gg(q);
}

float* ref_D(void* h) { return ((Ref*)h)->sdf->D; }
float* ref_W(void* h) { return ((Ref*)h)->sdf->W; }
float* ref_color(void* h, int which) {
    SDF* s = ((Ref*)h)->sdf;
    return which == 0 ? s->Color_W : which == 1 ? s->R : which == 2 ? s->G : s->B;
}
int64_t ref_number_of_voxels(void* h) { return ((Ref*)h)->sdf->get_number_of_voxels(); }

/* SDF::create_circle (sdf.cpp:99-126) */
void ref_create_circle(void* h, float radius, float cx, float cy, float cz) { ((Ref*)h)->sdf->create_circle(radius, cx, cy, cz); }

/* SDF::interpolate_color (sdf.cpp:164-217) */
void ref_interpolate_color(void* h, int64_t n, const double* global_pts, float* rgba) {
    Ref* r = (Ref*)h;
    for (int64_t q = 0; q < n; q++) {
        geometry_msgs::Point p;
        p.x = global_pts[3 * q]; p.y = global_pts[3 * q + 1]; p.z = global_pts[3 * q + 2];
        std_msgs::ColorRGBA c;
        r->sdf->interpolate_color(p, c);
        rgba[4 * q] = c.r; rgba[4 * q + 1] = c.g; rgba[4 * q + 2] = c.b; rgba[4 * q + 3] = c.a;
    }
}

/* pcl::MarchingCubesSDF::performReconstruction (marching_cubes_sdf.cpp:243-287) -> #vertices */
int64_t ref_mesh(void* h, float iso_level) {
    Ref* r = (Ref*)h;
    r->sdf->mc->setIsoLevel(iso_level);
    r->sdf->mc->performReconstruction(r->mesh);
    r->sdf->mc->setIsoLevel(0.0f);                          /* sdf.cpp:44 */
    return (int64_t)r->mesh.size();
}
void ref_mesh_copy(void* h, float* xyz) {
    Ref* r = (Ref*)h;
    for (size_t q = 0; q < r->mesh.size(); q++) {
        xyz[3 * q] = r->mesh.points[q].x; xyz[3 * q + 1] = r->mesh.points[q].y; xyz[3 * q + 2] = r->mesh.points[q].z;
    }
}

/* One pass of SDF::visualize's loop (sdf.cpp:317-391): mesh at iso 0, marker points + per-vertex colours.
 * The shim's ros::ok() lets the loop body run once and its Publisher keeps the marker.  -> #marker points */
int64_t ref_visualize(void* h) {
    Ref* r = (Ref*)h;
    r->sdf->initial_update_done = true;                     /* the condvar gate of sdf.cpp:321-323 */
    r->sdf->finish_visualization_thread = false;
    ros::shim::ok_budget() = 1;
    {
        CoutCapture cap;
        r->sdf->visualize(1000.0);
    }
    auto* mk = reinterpret_cast<visualization_msgs::Marker*>(ros::shim::last_marker_slot());
    return mk ? (int64_t)mk->points.size() : 0;
}
void ref_marker_copy(void* h, double* world, float* rgba) {
    auto* mk = reinterpret_cast<visualization_msgs::Marker*>(ros::shim::last_marker_slot());
    if (!mk) return;
    for (size_t q = 0; q < mk->points.size(); q++) {
        world[3 * q] = mk->points[q].x; world[3 * q + 1] = mk->points[q].y; world[3 * q + 2] = mk->points[q].z;
        rgba[4 * q] = mk->colors[q].r; rgba[4 * q + 1] = mk->colors[q].g; rgba[4 * q + 2] = mk->colors[q].b; rgba[4 * q + 3] = mk->colors[q].a;
    }
}

void ref_get_constants(void* h, float out[6]) {
    Ref* r = (Ref*)h;
    out[0] = r->sdf->m_div_width; out[1] = r->sdf->m_div_height; out[2] = r->sdf->m_div_depth;
    out[3] = r->tracker->v_h2_width; out[4] = r->tracker->v_h2_height; out[5] = r->tracker->v_h2_depth;
}

/* accumulated wall time inside estimate_new_position / update — exactly the spans the reference itself
 * prints ("camera estimation method", camera_tracking.cpp:68,243; "update method", sdf.cpp:225,306) */
void ref_timers(void* h, double out[2], int reset) {
    Ref* r = (Ref*)h;
    out[0] = r->t_track; out[1] = r->t_update;
    if (reset) { r->t_track = 0; r->t_update = 0; }
}

int ref_num_threads(void) { return omp_get_max_threads(); }
void ref_set_num_threads(int n) { omp_set_num_threads(n); }

}  // extern "C"
