/*
 * tsdf_b200.h — C ABI of the B200-native track + fuse hot path (libtsdf_b200.so).
 *
 * Drop-in boundary for the per-frame path of mees/tracking_sdf (Bylow et al. 2013).  The
 * reference has no FFI/plugin layer: the boundary is the public surface of its two C++
 * classes, `SDF` and `CameraTracking`, as called from `SDF_Reconstruction` (the ROS node).
 * Each entry point below names the reference interface it replaces; paths are relative to
 * /root/reference/src/.  The C++ classes with the reference's own method names are in
 * include/tracking_sdf_b200.hpp (header-only, over this ABI); INTEGRATION.md shows the
 * binding a maintainer of the ROS node would add.
 *
 * Conventions
 *  - plain pointers and sizes only; no exceptions cross the ABI; every call returns a
 *    tsdf_status (0 = ok) and tsdf_last_error() gives the text of the last failure.
 *  - R is a row-major 3x3 camera->world rotation, t the camera centre in world (the
 *    reference's `rot` / `trans`, camera_tracking.h:42-47).  K is row-major 3x3.
 *  - depth images are float32 metres along the camera z axis, row-major [height][width],
 *    NaN / inf / <= 0 = invalid.  `mem` says whether the pointer is host or device memory.
 *  - voxel (i,j,k) = (x,y,z) index exactly as in the reference (sdf.h:113-157).  The device
 *    layout is private (x-fastest, {D,W} interleaved); grids cross the ABI in one of the
 *    two documented layouts of tsdf_layout.
 *  - there is NO CPU fallback: every compute entry point needs a CUDA device and fails
 *    with TSDF_ERR_CUDA otherwise.
 */
#ifndef TSDF_B200_H_
#define TSDF_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TSDF_ABI_VERSION 1

typedef enum tsdf_status {
    TSDF_OK = 0,
    TSDF_ERR_BAD_ARG = 1,
    TSDF_ERR_NO_INTRINSICS = 2,   /* the reference exit(0)s here, sdf.cpp:227-229 */
    TSDF_ERR_CUDA = 3,
    TSDF_ERR_TRACKING_LOST = 4,   /* singular normal equations / non-finite twist (unguarded at camera_tracking.cpp:191) */
    TSDF_ERR_HALO = 5,            /* sharded: a tracking sample needed a voxel outside slab+halo */
    TSDF_ERR_NOMEM = 6,
    TSDF_ERR_PEER = 7             /* sharded: a peer rank did not deliver its normal equations in time; pose kept */
} tsdf_status;

typedef enum tsdf_metric {
    TSDF_POINT_TO_PLANE = 0,      /* active in the reference, sdf.cpp:272, sdf.h:177-181 */
    TSDF_POINT_TO_POINT = 1       /* defined, call commented out, sdf.cpp:267, sdf.h:169-172 */
} tsdf_metric;

typedef enum tsdf_mem { TSDF_HOST = 0, TSDF_DEVICE = 1 } tsdf_mem;

typedef enum tsdf_layout {
    TSDF_LAYOUT_REFERENCE = 0,    /* separate D[], W[]; z-fastest idx = m*m*i + m*j + k (sdf.h:120) */
    TSDF_LAYOUT_XFASTEST = 1      /* separate D[], W[]; x-fastest idx = (k*m + j)*m + i */
} tsdf_layout;

/* Replaces the ctor arguments of SDF (sdf.h:78-79) and CameraTracking
 * (camera_tracking.cpp:3-4), which the node hard-codes at sdf_reconstruction.cpp:83-88. */
typedef struct tsdf_config {
    int32_t m;                          /* voxels per axis (256)                       */
    float width, height, depth;         /* metric extents x,y,z (6, 6, 3.5)             */
    double origin[3];                   /* grid corner in world (-3,-3,-0.5)            */
    float distance_delta;               /* truncation delta (0.3)                       */
    float distance_epsilon;             /* weight plateau epsilon (0.025)               */
    int32_t gauss_newton_max_iteration; /* 20                                           */
    float maximum_twist_diff;           /* signed stop threshold 0.001; -INFINITY = fixed iteration count */
    float v_h;                          /* translational step in voxels (1.0)           */
    float w_h;                          /* rotational step in rad (0.01)                */
    int32_t pixel_stride;               /* 3 (camera_tracking.cpp:162-163)              */
    int32_t metric;                     /* tsdf_metric                                  */
    int32_t image_width, image_height;  /* 640 x 480                                    */
    int32_t device;                     /* CUDA device ordinal                          */
    /* z-slab sharding.  n_shards = 1: the whole volume on `device`.     */
    int32_t n_shards;                   /* slabs the volume is cut into along z         */
    int32_t shard_rank;                 /* which slab this handle owns                  */
    int32_t halo;                       /* extra z layers kept (and fused) on each side; <0 = auto */
    /* explicit owned layers [slab_k_begin, slab_k_end) of this slab; 0,0 = equal thickness.  Equal
     * slabs are not equal work (the view frustum is not uniform in z): tsdf_balanced_slabs cuts the
     * volume by a per-layer cost profile instead.  Every rank must use the same partition. */
    int32_t slab_k_begin, slab_k_end;
    /* 1: run K0 on every frame before anything else — the pre-processing the reference's node applies at
     * sdf_reconstruction.cpp:37-49 (pcl::FastBilateralFilter with PCL's defaults, then
     * pcl::IntegralImageNormalEstimation AVERAGE_3D_GRADIENT, 0.02, 10): the FILTERED depth feeds tracking and
     * fusion (cloud_filtered at :70, :74) and its normals replace K1's own.  PCL is un-vendored: the definition
     * is this library's (shared with the oracle), parity unpinned at that boundary.  0 (default): no filter,
     * K1's 4-neighbour normals — the noise-free benchmark configuration. */
    int32_t preprocess;
    int32_t reserved[1];
} tsdf_config;

typedef struct tsdf_track_stats {
    int32_t iterations;                 /* GN iterations executed                       */
    int32_t stopped;                    /* the signed stop test fired                   */
    int32_t n_valid;                    /* pixels in the last iteration's sums (all shards) */
    int32_t n_oob;                      /* pixels whose centre sample left the volume   */
    int32_t singular;                   /* normal equations singular / twist non-finite */
    int32_t halo_miss;                  /* sharded: samples that needed a voxel we do not hold */
    double residual;                    /* sum psi^2 of the last iteration              */
    double A[36];                       /* last iteration's J^T J (row-major, symmetric) */
    double b[6];                        /* last iteration's J^T r                       */
    double twist[6];                    /* last solved twist (v, w)                     */
} tsdf_track_stats;

typedef struct tsdf_handle_s* tsdf_handle;

int32_t     tsdf_abi_version(void);
const char* tsdf_last_error(void);
int32_t     tsdf_device_count(void);

void        tsdf_default_config(tsdf_config* cfg);          /* sdf_reconstruction.cpp:83-88 */

/* SDF::SDF + CameraTracking::CameraTracking (sdf.cpp:8-51, camera_tracking.cpp:3-18):
 * allocates the grid in HBM, D = width+height+depth, W = 0, pose = the reference's initial pose. */
tsdf_status tsdf_create(const tsdf_config* cfg, tsdf_handle* out);
tsdf_status tsdf_destroy(tsdf_handle h);
tsdf_status tsdf_reset(tsdf_handle h);                      /* re-run the grid init of sdf.cpp:28-31 */
tsdf_status tsdf_get_config(tsdf_handle h, tsdf_config* cfg);

/* CameraTracking::camera_info_cb (camera_tracking.cpp:22-36) */
tsdf_status tsdf_set_intrinsics(tsdf_handle h, const double K[9]);

/* CameraTracking::set_camera_transformation (camera_tracking.cpp:59-65) and the public
 * rot / trans / rot_inv / rot_inv_trans members the node reads (sdf_reconstruction.cpp:71). */
tsdf_status tsdf_set_pose(tsdf_handle h, const double R[9], const double t[3]);
tsdf_status tsdf_get_pose(tsdf_handle h, double R[9], double t[3]);
tsdf_status tsdf_get_pose_inv(tsdf_handle h, double Rinv[9], double tinv[3]);

/* CameraTracking::estimate_new_position (camera_tracking.cpp:66-245): Gauss-Newton on the
 * SDF from the current pose; returns the new pose.  stats may be NULL. */
tsdf_status tsdf_track(tsdf_handle h, const float* depth, int32_t mem,
                       double R_out[9], double t_out[3], tsdf_track_stats* stats);

/* SDF::update (sdf.cpp:224-305, D/W part).  R,t = NULL,NULL: fuse at the current pose
 * (what the node does, sdf_reconstruction.cpp:74); otherwise set the pose first (the
 * _useGroundTruth branch, sdf_reconstruction.cpp:61-66).  n_updated may be NULL. */
tsdf_status tsdf_fuse(tsdf_handle h, const float* depth, int32_t mem,
                      const double R[9], const double t[3], int64_t* n_updated);

/* kinect_callback's else-branch + update in one call (sdf_reconstruction.cpp:69-74):
 * track then fuse at the tracked pose; the pose never visits the host in between. */
tsdf_status tsdf_track_and_fuse(tsdf_handle h, const float* depth, int32_t mem,
                                double R_out[9], double t_out[3],
                                tsdf_track_stats* stats, int64_t* n_updated);

/* ---- colour (the second half of SDF::update, sdf.cpp:294-304, and SDF::interpolate_color,
 * sdf.cpp:164-217).  The colour store (Color_W, R, G, B: 16 B per voxel, sdf.cpp:14-17, initial
 * values 0 / 0.4 / 0.4 / 0.4, sdf.cpp:30-34) is allocated by tsdf_enable_color; the *_rgb calls
 * enable it implicitly.  rgb: height*width*3 bytes (the r,g,b of pcl::PointXYZRGB at (col,row)),
 * registered to the depth image, in the same memory space as `depth`.  Only the point-to-plane
 * metric has a colour update in the reference (it needs the normal); TSDF_ERR_BAD_ARG otherwise. */
tsdf_status tsdf_enable_color(tsdf_handle h);
tsdf_status tsdf_fuse_rgb(tsdf_handle h, const float* depth, const uint8_t* rgb, int32_t mem,
                          const double R[9], const double t[3], int64_t* n_updated);
tsdf_status tsdf_track_and_fuse_rgb(tsdf_handle h, const float* depth, const uint8_t* rgb, int32_t mem,
                                    double R_out[9], double t_out[3],
                                    tsdf_track_stats* stats, int64_t* n_updated);
/* n WORLD points (host, n x 3 doubles) -> n x (r,g,b,a) floats (host), exactly as
 * SDF::interpolate_color: interpolated values are scaled by 1/255, an exact voxel hit is not
 * (sdf.cpp:193-198), nothing in range gives NaN. */
tsdf_status tsdf_interpolate_color(tsdf_handle h, int64_t n, const double* global_pts, float* rgba);
/* the four colour arrays of this handle's stored z range, same layouts as tsdf_download */
tsdf_status tsdf_download_color(tsdf_handle h, float* color_w, float* r, float* g, float* b, int32_t layout);

/* ---- mesh (the visualisation thread's consumer of D/W: pcl::MarchingCubesSDF::performReconstruction,
 * marching_cubes_sdf.cpp:243-287, called from SDF::visualize, sdf.cpp:327).  Marching cubes over the
 * interior cells (m-2)^3 (sdf.cpp:36-39), a cell takes part only when all eight corners have W > 0
 * (marching_cubes_sdf.cpp:221); output = unindexed triangle soup, three vertices per triangle, cells in
 * the reference's index order, in the mesher's own frame: x,y,z in [0,width]x[0,height]x[0,depth] with
 * vertex = extent * index / m (no half-voxel offset, :123-125).  iso_level outside [0,1) gives an empty
 * mesh like the reference (:248-254).  The mesh stays on the device until the next extract / destroy.
 * On a z-slab handle only the cells of the owned layers are meshed. */
/* Work-balanced z partition (host only): weights[m] = relative fusion cost of each layer (e.g. the
 * number of in-view voxels for a few representative poses); bounds[n_shards + 1] receives the cuts
 * (bounds[0] = 0, bounds[n_shards] = m, every slab at least min_layers thick) that minimise the
 * largest per-slab cost, where a slab's cost is the weight of every layer it fuses, i.e. its own
 * layers plus `halo` layers on each side (use tsdf_slab_plan's halo).  Slab r = [bounds[r], bounds[r+1]). */
tsdf_status tsdf_balanced_slabs(int32_t m, int32_t n_shards, const double* weights, int32_t min_layers, int32_t halo, int32_t* bounds);
/* The same with a second per-layer profile that counts for the OWNER of a layer only (no halo): the tracked pixels
 * whose centre cell lies in the layer, in the same cost unit as `weights` — a rank linearises only the pixels it
 * owns, so a slab's frame time is (what it fuses, halo included) + (what it tracks).  weights_own may be NULL. */
tsdf_status tsdf_balanced_slabs2(int32_t m, int32_t n_shards, const double* weights, const double* weights_own,
                                 int32_t min_layers, int32_t halo, int32_t* bounds);

tsdf_status tsdf_mesh_extract(tsdf_handle h, float iso_level, int64_t* n_vertices);
/* copy the last mesh to the host; any pointer may be NULL.  xyz: n*3 floats as above; world: n*3
 * doubles = (double)xyz + sdf_origin, the marker points of sdf.cpp:354-356; rgba: n*4 floats =
 * SDF::interpolate_color at those points (sdf.cpp:380-385; needs the colour store). */
tsdf_status tsdf_mesh_download(tsdf_handle h, float* xyz, double* world, float* rgba);

/* Asynchronous variant for streaming: enqueue track+fuse of a DEVICE-resident frame; the
 * pose of frame `slot` lands in an internal pinned ring (capacity tsdf_pose_ring_capacity)
 * and is read back after tsdf_sync with tsdf_read_pose_ring.  track = 0: fuse only.
 * The frame's preprocessing (back-projection, normals, certificates) runs on a second stream and
 * overlaps the previous frame's tracking and fusion, so depth_dev must be completely written when
 * the call is made and must not change until tsdf_sync (or until two later frames were enqueued:
 * the call waits on the host for the preprocessing of the frame enqueued two calls earlier). */
tsdf_status tsdf_enqueue_frame(tsdf_handle h, const float* depth_dev, int32_t track, int32_t slot);
/* The same with a HOST (preferably pinned) depth buffer: the H2D copy goes through a copy stream
 * into a small ring of device frames, so the copy and the preprocessing of frame n+1 overlap
 * track+fuse of frame n.
 * The host buffer must stay valid until tsdf_sync (or until 4 later submissions returned: the call
 * waits on the host for the H2D copy issued four submissions earlier). */
tsdf_status tsdf_submit_frame(tsdf_handle h, const float* depth_host, int32_t track, int32_t slot);
tsdf_status tsdf_sync(tsdf_handle h);
int32_t     tsdf_pose_ring_capacity(void);
/* Returns the frame's tracking status exactly like tsdf_track: TSDF_ERR_TRACKING_LOST, TSDF_ERR_HALO or
 * TSDF_ERR_PEER when that frame's record says so (R, t, stats are filled in either way). */
tsdf_status tsdf_read_pose_ring(tsdf_handle h, int32_t slot, double R[9], double t[3], tsdf_track_stats* stats);

/* One linearisation at the current pose with NO pose update (the body of the loop at
 * camera_tracking.cpp:81-189): A row-major 6x6, b 6.  For parity tests. */
tsdf_status tsdf_linearize(tsdf_handle h, const float* depth, int32_t mem,
                           double A[36], double b[6], tsdf_track_stats* stats);
/* per-pixel records of that linearisation in the reference's loop order (column outer, row
 * inner): J [n*6], psi [n], flag [n] (0 NaN point, 1 ok, 2 out of volume, 3 not interpolated).
 * n must equal tsdf_num_strided_pixels(). */
int32_t     tsdf_num_strided_pixels(tsdf_handle h);
tsdf_status tsdf_linearize_pixels(tsdf_handle h, const float* depth, int32_t mem,
                                  float* J, float* psi, uint8_t* flag);

/* K1 (not in the reference; upstream ROS depth_image_proc + PCL normals): organised cloud
 * and normals, each [height*width*3] floats on the host, NaN = invalid. normals may be NULL. */
tsdf_status tsdf_backproject(tsdf_handle h, const float* depth, int32_t mem, float* cloud, float* normals);

/* K0 alone (needs tsdf_config.preprocess = 1): the filtered depth image [height*width] and the normals
 * [height*width*3, NaN = none] the frame paths would use; host buffers.  normals may be NULL.
 * Replaces the node's calls at sdf_reconstruction.cpp:37-49. */
tsdf_status tsdf_preprocess(tsdf_handle h, const float* depth, int32_t mem, float* depth_filtered, float* normals);

/* SDF::interpolate_distance (sdf.cpp:127-163) evaluated on the device: pts n x 3 doubles in
 * continuous voxel coordinates (host), out n floats, ok n bytes (host). */
tsdf_status tsdf_interpolate_distance(tsdf_handle h, int64_t n, const double* pts, float* out, uint8_t* ok);

/* Grid accessors: SDF::get_number_of_voxels (sdf.h:107) and the raw D/W arrays the reference
 * hands to its mesher (sdf.cpp:47-48).  Host buffers hold this handle's stored z range
 * [k_begin, k_end) (the whole grid when n_shards = 1) in the requested layout. */
int64_t     tsdf_number_of_voxels(tsdf_handle h);
tsdf_status tsdf_stored_range(tsdf_handle h, int32_t* k_begin, int32_t* k_end, int32_t* k_own_begin, int32_t* k_own_end);
tsdf_status tsdf_download(tsdf_handle h, float* D, float* W, int32_t layout);
tsdf_status tsdf_upload(tsdf_handle h, const float* D, const float* W, int32_t layout);
/* device pointer to the private interleaved store: float2{D,W} at ((k-k_begin)*m + j)*m + i */
tsdf_status tsdf_device_grid(tsdf_handle h, void** dw_interleaved, int64_t* n_voxels_stored);

/* Pure index / coordinate maps, identical (i,j,k) semantics to sdf.h:113-157 */
int64_t     tsdf_get_array_index(tsdf_handle h, int32_t i, int32_t j, int32_t k);        /* reference (z-fastest) index or -1 */
void        tsdf_get_voxel_coordinates_idx(tsdf_handle h, int64_t idx, int32_t ijk[3]);
void        tsdf_get_voxel_coordinates(tsdf_handle h, const double g[3], double v[3]);
void        tsdf_get_global_coordinates(tsdf_handle h, const int32_t ijk[3], double g[3]);

/* eigen_utils::direct_exponential_map (eigen_utils.cpp:85-128) evaluated on the device */
tsdf_status tsdf_exp_map(tsdf_handle h, const double twist[6], double R[9], double t[3]);

/* Device-memory helpers so callers without a CUDA runtime binding (ctypes, cgo) can keep
 * frames resident in HBM. */
tsdf_status tsdf_dev_alloc(tsdf_handle h, int64_t bytes, void** dev_ptr);
tsdf_status tsdf_dev_free(tsdf_handle h, void* dev_ptr);
tsdf_status tsdf_dev_upload(tsdf_handle h, void* dev_dst, const void* host_src, int64_t bytes);
tsdf_status tsdf_host_alloc_pinned(int64_t bytes, void** host_ptr);
tsdf_status tsdf_host_free_pinned(void* host_ptr);

/* Timing of the last enqueued work, CUDA events on the handle's stream (milliseconds):
 * out[0] prep (back-projection+normals), out[1] tracking (all GN iterations), out[2] fusion. */
tsdf_status tsdf_last_stage_ms(tsdf_handle h, float out[3]);
tsdf_status tsdf_event_timer_begin(tsdf_handle h);              /* record start on the stream   */
tsdf_status tsdf_event_timer_end(tsdf_handle h, float* ms);     /* record stop, sync, elapsed   */
int64_t     tsdf_kernel_launch_count(tsdf_handle h);            /* kernels launched so far      */
/* per-frame stage times over a run of enqueued frames: begin(n) arms n x 4 events, every
 * following frame records into them, end() syncs and returns ms[n][3] = {prep, track, fuse} */
tsdf_status tsdf_stage_timing_begin(tsdf_handle h, int32_t n_frames);
tsdf_status tsdf_stage_timing_end(tsdf_handle h, int32_t* n_frames, float* ms);
/* debugging aid: globaltimer stamps (ns) of the phases of one linearise+update launch */
tsdf_status tsdf_debug_phase_times(tsdf_handle h, const float* depth, int32_t mem, int64_t out[5]);
/* debugging aid: exhaustive device check of the tracker's short fp32 reciprocal against IEEE
 * 1.0f/x for every float in [x_lo, x_hi]; *n_bad = mismatches (must be 0 on [2^-17, 4]) */
tsdf_status tsdf_debug_check_rcp(tsdf_handle h, float x_lo, float x_hi, int64_t* n_bad);
/* Debugging aid: the fusion weight of sdf.cpp:278, w = (float)exp(-0.5 e^2), evaluated on the device for EVERY
 * float e = d_new - epsilon in [e_lo, e_hi] (0 <= e_lo).  *n_ambiguous counts the operands whose device weight is
 * not provably the correctly rounded float of the true exponential; the first `cap` of them are returned as
 * (e_list[i], w_list[i]) so a test can compare them with the host libm the reference calls. */
tsdf_status tsdf_debug_check_weight_exp(tsdf_handle h, float e_lo, float e_hi, int64_t* n_ambiguous, float* e_list, float* w_list, int32_t cap);
/* debugging aid: fusion self-check at the current pose, no voxel written: out[0] = voxels decided
 * by the certified fp32 fast path, out[1] = of those, verdicts that disagree with the exact fp64
 * path (must be 0), out[2] = work items */
tsdf_status tsdf_debug_fuse_check(tsdf_handle h, const float* depth, int32_t mem, int64_t out[3]);
/* roofline ceiling probe: mean milliseconds of a plain read-modify-write stream over the whole
 * store (the free-space update on every voxel, no geometry).  MODIFIES the grid. */
tsdf_status tsdf_debug_stream_rmw(tsdf_handle h, int32_t reps, float* ms);
/* running total of voxels updated by fusion since the last reset (for GB/s accounting) */
tsdf_status tsdf_total_updates(tsdf_handle h, int32_t reset, int64_t* total);
tsdf_status tsdf_flush_l2(tsdf_handle h);                       /* overwrite a >L2-sized scratch buffer */

/* Pure host function (no device needed): the z-slab plan tsdf_create would use for `cfg`:
 * out = {k_own_begin, k_own_end, k_stored_begin, k_stored_end, halo}. */
tsdf_status tsdf_slab_plan(const tsdf_config* cfg, int32_t out[5]);

/* Sharded tracking across processes (one process per GPU): each rank exports a handle to
 * its 30-double mailbox, the caller exchanges them (e.g. torch.distributed all_gather) and
 * attaches all of them; the per-iteration all-reduce of the normal equations then runs
 * inside the tracking kernels over NVLink peer stores, summed in rank order. */
#define TSDF_IPC_HANDLE_BYTES 64
tsdf_status tsdf_shard_ipc_export(tsdf_handle h, uint8_t out[TSDF_IPC_HANDLE_BYTES]);
tsdf_status tsdf_shard_ipc_attach(tsdf_handle h, int32_t world, const uint8_t* handles /* world x 64 */);
/* Same, for shards living in one process: `handles` ordered by shard_rank.  Shards on
 * different devices exchange in-kernel over peer memory; shards that share one device
 * (testing the slab logic on a single GPU) run on one stream with a deferred rank-order sum. */
tsdf_status tsdf_shard_attach_local(tsdf_handle* handles, int32_t world);
/* Group entry points for in-process shards: the same operations as tsdf_set_pose /
 * tsdf_linearize / tsdf_track_and_fuse, issued to every shard in lock step.  depth is a HOST
 * buffer, or a device buffer when all shards share one device.  Results are shard 0's (all
 * shards hold identical poses); n_updated sums the voxels each shard owns. */
tsdf_status tsdf_group_set_intrinsics(tsdf_handle* handles, int32_t world, const double K[9]);
tsdf_status tsdf_group_set_pose(tsdf_handle* handles, int32_t world, const double R[9], const double t[3]);
tsdf_status tsdf_group_linearize(tsdf_handle* handles, int32_t world, const float* depth, int32_t mem,
                                 double A[36], double b[6], tsdf_track_stats* stats);
tsdf_status tsdf_group_frame(tsdf_handle* handles, int32_t world, const float* depth, int32_t mem,
                             int32_t track, int32_t fuse, double R_out[9], double t_out[3],
                             tsdf_track_stats* stats, int64_t* n_updated);

#ifdef __cplusplus
}
#endif
#endif /* TSDF_B200_H_ */
