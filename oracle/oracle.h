/*
 * oracle.h — C ABI of the CPU ORACLE for the track+fuse hot path of mees/tracking_sdf.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (tracking_sdf_b200/, include/) may
 * import, link or call this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and there only as the checker / the
 * timed CPU baseline, never as a fallback.
 *
 * PINNED AGAINST THE REFERENCE ITSELF: the reference has no tests, golden vectors or fixtures
 *  and its build needs ROS, PCL, Eigen and Boost, none of which exist in this
 * image — but its four hot-path translation units (sdf.cpp, camera_tracking.cpp, eigen_utils.cpp,
 * marching_cubes_sdf.cpp) compile UNMODIFIED against the small shim headers in oracle/shim/
 * (oracle/Makefile target `ref` -> oracle/_ref/libtsdf_ref.so, C ABI in ref_bridge.cpp).
 * tests/test_oracle_vs_ref.py compares this restatement with that library function by function:
 * bit-equal for the index maps, interpolate_distance, update (D, W, colour), get_partial_derivative
 * (J, psi per pixel), the normal equations (one thread), the exp map, set_camera_transformation,
 * marching cubes, interpolate_color; a few ulp for A.inverse()*b (Eigen's 6x6 LU inverse, stated
 * in oracle/shim/eigen_shim.h, not verifiable without Eigen).  tests/golden/ is generated from
 * that library.  STILL UNPINNED (third-party code that is not under /root/reference): Eigen's
 * own rounding inside Matrix<6,6>::inverse() and Affine3d::rotation(); the ROS depth_image_proc
 * back-projection and the PCL normals upstream of the boundary (orc_backproject is one definition
 * shared by oracle and GPU).  Each function in oracle.cpp cites the reference file:line it follows.
 */
#ifndef TSDF_ORACLE_H_
#define TSDF_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_config {
    int32_t m;                 /* voxels per axis                      sdf_reconstruction.cpp:85 */
    float width, height, depth;/* metric extents (x, y, z)             sdf_reconstruction.cpp:85 */
    double origin[3];          /* world position of the grid corner    sdf_reconstruction.cpp:83 */
    float distance_delta;      /* truncation delta                     sdf_reconstruction.cpp:85 */
    float distance_epsilon;    /* weight plateau epsilon               sdf_reconstruction.cpp:85 */
    int32_t gauss_newton_max_iteration; /*                             sdf_reconstruction.cpp:88 */
    float maximum_twist_diff;  /* signed stop threshold                sdf_reconstruction.cpp:88 */
    float v_h;                 /* translational step, voxels (1.0)     camera_tracking.cpp:3-4,11 */
    float w_h;                 /* rotational step, rad (0.01)          camera_tracking.cpp:3-4,12 */
    int32_t pixel_stride;      /* 3                                    camera_tracking.cpp:162-163 */
    int32_t metric;            /* 0 = point-to-plane (sdf.cpp:272), 1 = point-to-point (sdf.h:169) */
    int32_t image_width, image_height;
    int32_t use_coord_table;   /* 1 = precompute global_coords[] like sdf.cpp:11,40-41 */
    int32_t preprocess;        /* 1 = K0 before K1 (the node's pre-processing, sdf_reconstruction.cpp:37-49): fast bilateral
                                * depth filter + average-3D-gradient normals over an adaptive window; 0 = K1's own normals */
} orc_config;

typedef struct orc_track_stats {
    int32_t iterations;        /* GN iterations executed                                   */
    int32_t stopped;           /* 1 if the signed stop test fired (camera_tracking.cpp:216) */
    int32_t n_valid;           /* pixels that contributed in the LAST iteration             */
    int32_t n_oob;             /* pixels whose centre left the volume (TRAP 5), last iter   */
    int32_t singular;          /* 1 if the 6x6 system had a zero pivot / non-finite twist   */
    int32_t pad;
    double residual;           /* sum psi^2 over valid pixels, last iteration (addition)    */
} orc_track_stats;

void orc_default_config(orc_config* cfg);

void* orc_create(const orc_config* cfg);
void  orc_destroy(void* h);

void orc_set_intrinsics(void* h, const double K[9]);                 /* camera_tracking.cpp:22-36 */
void orc_set_pose(void* h, const double R[9], const double t[3]);    /* camera_tracking.cpp:59-65 */
void orc_get_pose(void* h, double R[9], double t[3]);
void orc_get_pose_inv(void* h, double Rinv[9], double tinv[3]);

/* K0 (call sites sdf_reconstruction.cpp:37-49).  The node filters the organised cloud with
 * pcl::FastBilateralFilter (defaults sigma_s = 15 px, sigma_r = 0.05 m) and estimates normals with
 * pcl::IntegralImageNormalEstimation (AVERAGE_3D_GRADIENT, MaxDepthChangeFactor 0.02, NormalSmoothingSize 10).
 * PCL is not under /root/reference and not in this image: PARITY UNPINNED at this boundary.  What is restated
 * here is PCL's published algorithm (bilateral grid: splat, [1 2 1]/4 blur twice per axis, trilinear slice;
 * normals: cross product of the summed central-difference 3-D gradients over a window whose size is the chamfer
 * distance to the nearest depth discontinuity, capped at the smoothing size), as ONE definition shared with the
 * device kernels (tsdf_k0.cu), which must reproduce it bit for bit.  depth_out: filtered depth [h*w];
 * normals: [h*w*3], NaN = none. */
void orc_preprocess(void* h, const float* depth, float* depth_out, float* normals);
void orc_k0_normals(void* h, const float* depth_filtered, float* normals);   /* the normal stage alone */

/* K1: depth -> organised cloud (x,y,z) + normals; both [h*w*3] floats, NaN = invalid */
void orc_backproject(void* h, const float* depth, float* cloud, float* normals);

/* sdf.cpp:224-305 (D/W part). Uses the current pose + K.  Returns #voxels updated. */
int64_t orc_fuse(void* h, const float* depth);
int64_t orc_fuse_cloud(void* h, const float* cloud, const float* normals);
/* sdf.cpp:224-305 including the colour running mean (:294-304); rgb = h*w*3 bytes registered to the depth */
int64_t orc_fuse_rgb(void* h, const float* depth, const uint8_t* rgb);
void    orc_enable_color(void* h);                                   /* sdf.cpp:14-17,30-34: Color_W = 0, R = G = B = 0.4 */
float*  orc_color(void* h, int which);                               /* 0 Color_W, 1 R, 2 G, 3 B (reference layout) */
/* sdf.cpp:164-217, n world points -> n x (r,g,b,a) */
void    orc_interpolate_color(void* h, int64_t n, const double* global_pts, float* rgba);
/* pcl::MarchingCubesSDF::performReconstruction (marching_cubes_sdf.cpp:243-287) -> #vertices; then copy out */
int64_t orc_mesh(void* h, float iso_level);
void    orc_mesh_copy(void* h, float* xyz, double* world, float* rgba);

/* camera_tracking.cpp:66-245 */
void orc_track(void* h, const float* depth, orc_track_stats* stats);

/* One linearisation at the current pose, no pose update.  A row-major 6x6, b 6. */
void orc_linearize(void* h, const float* depth, double A[36], double b[6], orc_track_stats* stats);

/* Per-pixel linearisation record for the strided pixels in reference loop order
 * (column i outer, row j inner): J [n*6] float, psi [n] float, flag [n] uint8
 * (0 = NaN point, 1 = ok, 2 = out of volume, 3 = some sample not interpolated).
 * Returns n = number of strided pixels. */
int32_t orc_linearize_pixels(void* h, const float* depth, float* J, float* psi, uint8_t* flag);

/* Solve + exp map + pose update for given A, b (camera_tracking.cpp:191-239).
 * Returns 1 if singular. twist_out 6 doubles. */
int32_t orc_apply_update(void* h, const double A[36], const double b[6], double twist_out[6]);

/* sdf.cpp:127-163.  pts: n x 3 doubles, continuous VOXEL coordinates. */
void orc_interpolate(void* h, int64_t n, const double* pts, float* out, uint8_t* ok);

/* eigen_utils.cpp:85-128.  twist (v, w); R row-major 3x3, t 3. */
void orc_exp_map(const double twist[6], double R[9], double t[3]);

/* index / coordinate maps, sdf.h:113-157 */
int64_t orc_get_array_index(void* h, int32_t i, int32_t j, int32_t k);
void    orc_get_voxel_coordinates_idx(void* h, int64_t idx, int32_t ijk[3]);
void    orc_get_voxel_coordinates(void* h, const double g[3], double v[3]);
void    orc_get_global_coordinates(void* h, const int32_t ijk[3], double g[3]);

/* raw grids in REFERENCE layout (z-fastest: idx = m*m*i + m*j + k, sdf.h:120) */
float*  orc_D(void* h);
float*  orc_W(void* h);
int64_t orc_number_of_voxels(void* h);
void    orc_reset(void* h);                                         /* sdf.cpp:29-31 */

/* sdf.cpp:99-126 analytic sphere filler (known-answer tests) */
void orc_create_circle(void* h, float radius, float cx, float cy, float cz);

/* derived fp32 constants, for tests */
void orc_get_constants(void* h, float out[6]); /* m_div_width,height,depth, v_h2_width,height,depth */

int orc_num_threads(void);
void orc_set_num_threads(int n);   /* launchers such as torchrun export OMP_NUM_THREADS=1: benches set the count explicitly */

#ifdef __cplusplus
}
#endif
#endif
