/*
 * tsdf_abi.cu — the C ABI of include/tsdf_b200.h: handle management, stream orchestration
 * of the K1/K2/K3 kernels, accessors.  Host C++ only; no PyTorch, no CPU compute path.
 */
#include "tsdf_b200.h"
#include "tsdf_internal.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

using namespace tsdf;

namespace {

thread_local std::string g_err;

constexpr int POSE_RING = 4096;
/* ticket words: [0] final tracker ticket, [1 .. MAX_LIN_GROUPS] group tickets, [PYRAMID_TICKET] k_pyramid (it runs
 * concurrently on the preprocessing stream, so it must never alias a group ticket) */
constexpr int MAX_LIN_GROUPS = 999;
constexpr int PYRAMID_TICKET = 1000;
constexpr int N_TICKETS = 1024;
constexpr int64_t FLUSH_BYTES = 256ll << 20;   /* > 126 MB L2 */

struct Impl {
    tsdf_config cfg;
    GridParams g;
    int device = 0;
    cudaStream_t stream = nullptr;
    int* stream_refs = nullptr;             /* handles sharing `stream` (same-device shard groups); the last one destroys it */
    float2* grid = nullptr;
    int64_t n_stored = 0;
    /* per-frame records, double-buffered: frame n+1 is preprocessed on prep_stream while frame n
     * is still being tracked and fused.  pix/pts/cert point at the current frame's buffers. */
    PixRec* pix = nullptr;
    float4* pts = nullptr;
    PixRec* pix_buf[2] = {nullptr, nullptr};
    float4* pts_buf[2] = {nullptr, nullptr};
    float2* cert_buf[2] = {nullptr, nullptr};
    cudaStream_t prep_stream = nullptr;
    cudaEvent_t prepped[2] = {nullptr, nullptr};      /* records of buffer b are complete          */
    cudaEvent_t rec_free[2] = {nullptr, nullptr};     /* every consumer of buffer b has finished   */
    cudaEvent_t depth_copied = nullptr;               /* synchronous-API H2D of the depth image    */
    cudaEvent_t depth_ready = nullptr, depth_done = nullptr;   /* optional events for the next enqueue_prep */
    unsigned long long prep_seq = 0;
    int lin_first = 0, lin_iter = 0;
    /* colour (tsdf_enable_color): {Color_W, R, G, B} per voxel, the frame's packed colour image
     * (double-buffered like the other records) and a staging buffer for host images */
    float4* color = nullptr;
    uchar4* rgb4_buf[2] = {nullptr, nullptr};
    uchar4* rgb4 = nullptr;
    double* cosn_buf[2] = {nullptr, nullptr};
    double* cosn = nullptr;
    uint8_t* rgb_stage = nullptr;
    const uint8_t* frame_rgb = nullptr;           /* device pointer consumed by the next enqueue_prep (or NULL) */
    bool frame_has_color = false;                 /* the current records carry a colour image */
    int fuse_color_blocks = 0;
    /* mesher: per-row vertex counts / offsets (m*m + 1), scan scratch, the last mesh */
    unsigned int* mc_count = nullptr;
    unsigned int* mc_off = nullptr;
    void* mc_tmp = nullptr;
    size_t mc_tmp_bytes = 0;
    float* mesh_xyz = nullptr;
    int64_t mesh_n = 0, mesh_cap = 0;
    unsigned long long* mc_cells = nullptr;   /* surface-cell list of the one-sweep path */
    unsigned int* mc_cell_counter = nullptr;
    unsigned int mc_cell_cap = 0;
    float* depth_stage = nullptr;
    /* K0 (tsdf_config.preprocess): bilateral grid, filtered depth, discontinuity map, gradients, normals */
    bool k0 = false;
    K0Buffers k0b = {};
    PoseState* pose_dev = nullptr;
    PoseState* pose_pin = nullptr;          /* pinned: D2H landing zone for track results */
    PoseState* ring_pin = nullptr;          /* pinned pose ring for the async path */
    double* partials = nullptr;
    double* group_partials = nullptr;
    unsigned int* ticket = nullptr;         /* [0] final ticket, [1..] group tickets */
    double* fuse_tables = nullptr;
    unsigned long long* fuse_items = nullptr;
    float4* fuse_item_c = nullptr;
    unsigned int mc_last_cells = 0, mc_emit_threads = 0;   /* mesher: list length of the last extraction / threads of the speculative emit */
    unsigned int* fuse_item_count = nullptr;      /* [0] items, [1] units */
    unsigned long long* fuse_units = nullptr;
    float2* cert = nullptr;
    CertPyramid pyr;
    int fuse_cert_blocks = 0;
    int fuse_check = 0;
    unsigned long long* n_upd_dev = nullptr;
    unsigned long long* n_upd_pin = nullptr;
    float* dbgJ = nullptr; float* dbgPsi = nullptr; uint8_t* dbgFlag = nullptr;
    unsigned long long* dbg_times = nullptr;   /* set by tsdf_debug_phase_times */
    Mailbox* mailbox = nullptr;
    ShardLinks links;
    int exchange_mode = 0;
    unsigned long long seqno = 0;
    std::vector<void*> ipc_opened;
    bool have_K = false;
    int force_idx64 = 0;                    /* TSDF_B200_IDX64=1 in the environment at create (tests) */
    bool ev_recorded = false;               /* p->ev[] hold a frame (not the stage-timing ring) */
    cudaEvent_t ev[4];
    cudaEvent_t tmr[2];
    bool stage_valid = false;
    std::vector<cudaEvent_t> ring_ev;       /* optional per-frame stage events (4 per frame) */
    int ring_frames = 0, ring_pos = 0;
    int lin_blocks = 0, px_per_block = 0, fuse_blocks = 0;
    int64_t launches = 0;
    float* flush_buf = nullptr;
    /* host-buffer streaming (tsdf_submit_frame): a ring of device frames fed by a copy stream */
    static constexpr int NSTAGE = 4;
    cudaStream_t copy_stream = nullptr;
    float* stage[NSTAGE] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t copied[NSTAGE] = {nullptr, nullptr, nullptr, nullptr};    /* H2D of stage b finished   */
    cudaEvent_t consumed[NSTAGE] = {nullptr, nullptr, nullptr, nullptr};  /* compute done with stage b */
    unsigned long long submit_seq = 0;
    double* scratch_d = nullptr;            /* 32 doubles */
};

#define CK(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            g_err = std::string(#call) + ": " + cudaGetErrorString(e_);                           \
            return TSDF_ERR_CUDA;                                                                 \
        }                                                                                         \
    } while (0)

tsdf_status bad(const char* msg) { g_err = msg; return TSDF_ERR_BAD_ARG; }

/* every kernel launch and event / stream-wait call of the enqueue helpers reports into note_cuda();
 * this turns the first failure since the last check into a status */
tsdf_status launch_status() {
    cudaError_t e = take_launch_error();
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { g_err = std::string("enqueue failed: ") + cudaGetErrorString(e); return TSDF_ERR_CUDA; }
    return TSDF_OK;
}

/* device scratch that is released on every exit path of an accessor */
struct DevTmp {
    void* p = nullptr;
    ~DevTmp() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16); }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

Impl* I(tsdf_handle h) { return reinterpret_cast<Impl*>(h); }

/* the fp32/fp64 constants exactly as the reference's constructors build them */
void build_params(const tsdf_config& c, GridParams& g) {
    memset(&g, 0, sizeof g);
    g.m = c.m;
    g.metric = c.metric;
    g.img_w = c.image_width; g.img_h = c.image_height;
    g.stride = c.pixel_stride;
    g.ni = (c.image_width + c.pixel_stride - 1) / c.pixel_stride;
    g.nj = (c.image_height + c.pixel_stride - 1) / c.pixel_stride;
    g.m_div_height = c.m / c.height;                       /* sdf.cpp:19-21 */
    g.m_div_width = c.m / c.width;
    g.m_div_depth = c.m / c.depth;
    g.m_div_d[0] = (double)g.m_div_width; g.m_div_d[1] = (double)g.m_div_height; g.m_div_d[2] = (double)g.m_div_depth;
    g.vs_x = c.width / ((float)c.m);                       /* sdf.h:154-156 */
    g.vs_y = c.height / ((float)c.m);
    g.vs_z = c.depth / ((float)c.m);
    g.delta = c.distance_delta; g.eps = c.distance_epsilon;
    g.v_h = c.v_h; g.w_h = c.w_h;                          /* camera_tracking.cpp:11-17 */
    const float v_h2 = 2 * c.v_h;
    g.v_h2_width = v_h2 / g.m_div_width;
    g.v_h2_height = v_h2 / g.m_div_height;
    g.v_h2_depth = v_h2 / g.m_div_depth;
    g.two_w_h = 2 * (c.w_h);
    g.max_twist_diff = c.maximum_twist_diff;
    g.max_iter = c.gauss_newton_max_iteration;
    for (int q = 0; q < 3; q++) g.origin[q] = c.origin[q];
}

/* z-slab partition: rank r owns [r*m/G, (r+1)*m/G); it also stores and
 * fuses `halo` layers on each side so the tracker's stencil never leaves local memory. */
void slab_range(const tsdf_config& c, int& ko0, int& ko1, int& ks0, int& ks1, int& halo) {
    const int G = c.n_shards < 1 ? 1 : c.n_shards;
    ko0 = (int)(((int64_t)c.m * c.shard_rank) / G);
    ko1 = (int)(((int64_t)c.m * (c.shard_rank + 1)) / G);
    if (G > 1 && c.slab_k_end > c.slab_k_begin) { ko0 = c.slab_k_begin; ko1 = c.slab_k_end; }   /* explicit (balanced) partition */
    halo = c.halo;
    if (G == 1) halo = 0;
    else if (halo < 0) {
        /* centre cell (+1), +-v_h voxels, and the rotational perturbation: the sample moves by
         * w_h * e_k x v, v = world-frame vector from the camera centre to the point; its z component is
         * w_h * v_y (k = x), -w_h * v_x (k = y), 0 (k = z), so with the camera centre and the point both
         * over the volume's footprint |dz| <= w_h * max(width, height).  (A camera far outside the footprint
         * can exceed this: the tracker then reports TSDF_ERR_HALO and tsdf_config.halo must be set by hand.) */
        const double reach = (double)(c.width > c.height ? c.width : c.height);
        const double vz = (double)c.depth / c.m;
        halo = 2 + (int)ceil(c.v_h) + (int)ceil(c.w_h * reach / vz);
    }
    ks0 = ko0 - halo < 0 ? 0 : ko0 - halo;
    ks1 = ko1 + halo > c.m ? c.m : ko1 + halo;
}

tsdf_status upload_pose(Impl* p, const PoseState& ps) {
    memcpy(p->pose_pin, &ps, sizeof ps);
    CK(cudaMemcpyAsync(p->pose_dev, p->pose_pin, sizeof ps, cudaMemcpyHostToDevice, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return TSDF_OK;
}
tsdf_status fetch_pose(Impl* p) {
    CK(cudaMemcpyAsync(p->pose_pin, p->pose_dev, sizeof(PoseState), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return TSDF_OK;
}

tsdf_status stage_depth(Impl* p, const float* depth, int mem, const float** dptr) {
    if (!depth) return bad("depth is NULL");
    if (mem == TSDF_DEVICE) { *dptr = depth; return TSDF_OK; }
    const size_t bytes = (size_t)p->g.img_w * p->g.img_h * sizeof(float);
    CK(cudaMemcpyAsync(p->depth_stage, depth, bytes, cudaMemcpyHostToDevice, p->stream));
    CK(cudaEventRecord(p->depth_copied, p->stream));
    p->depth_ready = p->depth_copied;
    *dptr = p->depth_stage;
    return TSDF_OK;
}

/* colour image for the next frame: host images go through rgb_stage on the main stream (the
 * same event as the depth copy orders K1 behind both) */
tsdf_status stage_rgb(Impl* p, const uint8_t* rgb, int mem) {
    if (!rgb) return bad("rgb is NULL");
    if (p->g.metric != 0) return bad("colour fusion needs the point-to-plane metric (sdf.cpp:294 uses the normal)");
    if (!p->color) return bad("colour store not allocated (tsdf_enable_color)");
    if (mem == TSDF_DEVICE) { p->frame_rgb = rgb; return TSDF_OK; }
    const size_t bytes = (size_t)p->g.img_w * p->g.img_h * 3;
    CK(cudaMemcpyAsync(p->rgb_stage, rgb, bytes, cudaMemcpyHostToDevice, p->stream));
    CK(cudaEventRecord(p->depth_copied, p->stream));
    p->depth_ready = p->depth_copied;
    p->frame_rgb = p->rgb_stage;
    return TSDF_OK;
}

K0Params k0_params(const Impl* p) {
    K0Params P;
    P.img_w = p->g.img_w; P.img_h = p->g.img_h;
    P.sigma_s = 15.0f; P.sigma_r = 0.05f;              /* pcl::FastBilateralFilter defaults; the node sets none (:38-41) */
    P.max_depth_change = 0.02f; P.smoothing = 10.0f;   /* sdf_reconstruction.cpp:46-47 */
    const K1Params kp = k1_params(p->g.K);
    P.cx = kp.cx; P.cy = kp.cy; P.inv_fx = kp.inv_fx; P.inv_fy = kp.inv_fy;
    return P;
}

LinearizeArgs lin_args(Impl* p, int do_update, bool debug) {
    LinearizeArgs a;
    a.g = p->g;
    a.grid = p->grid; a.pix = p->pix; a.pts = p->pts; a.pose = p->pose_dev;
    a.partials = p->partials; a.ticket = p->ticket;
    a.group_ticket = p->ticket + 1; a.group_partials = p->group_partials;
    a.dbgJ = debug ? p->dbgJ : nullptr; a.dbgPsi = debug ? p->dbgPsi : nullptr; a.dbgFlag = debug ? p->dbgFlag : nullptr;
    a.dbg_times = p->dbg_times;
    a.do_update = do_update;
    a.first = 0;
    a.iter = 0;
    a.px_per_block = p->px_per_block;
    a.force_idx64 = p->force_idx64;
    a.links = p->links;
    return a;
}

/* K1 on prep_stream into the other record buffer.  It depends only on the depth image, so it
 * overlaps the previous frame's tracking tail and fusion; the main stream joins on `prepped`.
 * Buffer b is reused two frames later: rec_free[b] (recorded on the main stream one call later)
 * covers every consumer of it. */
void enqueue_prep(Impl* p, const float* dptr, int reset_track) {
    const int b = (int)(p->prep_seq & 1ull);
    note_cuda(cudaEventRecord(p->rec_free[b ^ 1], p->stream));   /* all work on the previous frame's records is enqueued */
    if (p->prep_seq >= 2) {
        /* HOST wait: the preprocessing of the frame enqueued two calls ago has read its depth image.  This is
         * what makes the documented lifetime true ("depth_dev must not change until two later frames were
         * enqueued"): the host can no longer run further ahead of K1 than that.  K1 runs early on its own
         * stream, so in steady state the event is long complete and the wait costs nothing. */
        note_cuda(cudaEventSynchronize(p->prepped[b]));
        note_cuda(cudaStreamWaitEvent(p->prep_stream, p->rec_free[b], 0));
    }
    if (p->depth_ready) note_cuda(cudaStreamWaitEvent(p->prep_stream, p->depth_ready, 0));
    p->pix = p->pix_buf[b]; p->pts = p->pts_buf[b]; p->cert = p->cert_buf[b]; p->rgb4 = p->rgb4_buf[b]; p->cosn = p->cosn_buf[b];
    p->frame_has_color = p->frame_rgb != nullptr;
    const float4* nrm_in = nullptr;
    if (p->k0) {
        /* sdf_reconstruction.cpp:37-49: filter the frame, estimate normals on the filtered cloud; the filtered
         * image is what tracking and fusion see (cloud_filtered at :70 and :74) */
        p->launches += launch_k0(k0_params(p), p->k0b, dptr, p->prep_stream);
        dptr = p->k0b.zf;
        nrm_in = p->k0b.normals;
    }
    launch_prep(p->g, dptr, nrm_in, p->pix, p->cert + p->pyr.off[0], p->pts, p->frame_rgb, p->rgb4, p->cosn, p->prep_stream);
    p->frame_rgb = nullptr;
    launch_pyramid(p->pyr, p->cert, p->ticket + PYRAMID_TICKET, p->prep_stream);
    if (p->depth_done) note_cuda(cudaEventRecord(p->depth_done, p->prep_stream));
    note_cuda(cudaEventRecord(p->prepped[b], p->prep_stream));
    note_cuda(cudaStreamWaitEvent(p->stream, p->prepped[b], 0));
    p->depth_ready = nullptr; p->depth_done = nullptr;
    p->prep_seq++;
    p->lin_first = reset_track;
    if (reset_track) p->lin_iter = 0;
    p->launches += 2;
}
void enqueue_linearize(Impl* p, int do_update, bool debug) {
    p->seqno++;
    LinearizeArgs a = lin_args(p, do_update, debug);
    a.first = p->lin_first;
    a.iter = p->lin_iter;
    p->lin_first = 0;
    if (do_update) p->lin_iter++;
    launch_linearize(a, p->lin_blocks, p->exchange_mode, p->seqno, p->stream);
    p->launches++;
}
void enqueue_combine(Impl* p, int do_update) {
    LinearizeArgs a = lin_args(p, do_update, false);
    a.iter = p->lin_iter > 0 ? p->lin_iter - 1 : 0;        /* the iteration enqueue_linearize just counted */
    launch_gn_combine(a, p->seqno, p->stream);
    p->launches++;
}
void enqueue_fuse(Impl* p) {
    FuseArgs f;
    f.g = p->g; f.grid = p->grid; f.pix = p->pix; f.pose = p->pose_dev;
    f.tables = p->fuse_tables; f.items = p->fuse_items; f.item_c = p->fuse_item_c; f.item_count = p->fuse_item_count;
    f.n_updated = p->n_upd_dev; f.nblk = p->fuse_blocks; f.nblk_cert = p->fuse_cert_blocks; f.check = p->fuse_check;
    f.pyr = p->pyr; f.cert = p->cert; f.units = p->fuse_units; f.unit_count = p->fuse_item_count + 1;
    f.color = p->frame_has_color ? p->color : nullptr; f.rgb4 = p->rgb4; f.cosn = p->cosn; f.nblk_color = p->fuse_color_blocks;
    p->launches += launch_fuse(f, p->stream);
}

/* the per-frame sequence of sdf_reconstruction.cpp:69-74 on the stream, no host round trip */
/* stage_events: record the four stage-boundary events (tsdf_last_stage_ms / tsdf_stage_timing_*).  An event record
 * between two kernels ends the programmatic-dependent-launch chain there (the successor can no longer start
 * early), so the streaming entry points record them only inside a tsdf_stage_timing_begin/end window. */
tsdf_status enqueue_frame(Impl* p, const float* dptr, bool do_track, bool do_fuse, bool stage_events = true) {
    if (!p->have_K) { g_err = "camera matrix not set (tsdf_set_intrinsics)"; return TSDF_ERR_NO_INTRINSICS; }
    if (p->exchange_mode == 2) return bad("same-device shard group: use tsdf_group_* entry points");
    cudaEvent_t* ev = p->ev;
    const bool in_window = p->ring_pos < p->ring_frames;
    if (in_window) { ev = &p->ring_ev[(size_t)4 * p->ring_pos]; p->ring_pos++; }
    const bool rec = stage_events || in_window;
    if (rec && !in_window) p->ev_recorded = true;
    if (rec) note_cuda(cudaEventRecord(ev[0], p->stream));
    enqueue_prep(p, dptr, do_track ? 1 : 0);
    if (rec) note_cuda(cudaEventRecord(ev[1], p->stream));
    if (do_track)
        for (int it = 0; it < p->g.max_iter; it++) enqueue_linearize(p, 1, false);
    if (rec) note_cuda(cudaEventRecord(ev[2], p->stream));
    if (do_fuse) enqueue_fuse(p);
    if (rec) note_cuda(cudaEventRecord(ev[3], p->stream));
    if (rec) p->stage_valid = true;
    return launch_status();
}

void fill_stats(const PoseState& ps, tsdf_track_stats* st) {
    if (!st) return;
    memset(st, 0, sizeof *st);
    st->iterations = ps.iterations; st->stopped = ps.stopped; st->singular = ps.singular; st->halo_miss = ps.halo_miss;
    st->n_valid = (int32_t)ps.sums[SLOT_NVALID]; st->n_oob = (int32_t)ps.sums[SLOT_NOOB];
    st->residual = ps.sums[SLOT_RES];
    int q = 0;
    for (int r = 0; r < 6; r++)
        for (int c = r; c < 6; c++) { st->A[6 * r + c] = ps.sums[SLOT_A + q]; st->A[6 * c + r] = ps.sums[SLOT_A + q]; q++; }
    for (int r = 0; r < 6; r++) { st->b[r] = ps.sums[SLOT_B + r]; st->twist[r] = ps.twist[r]; }
}

/* status of one frame's tracking record — the same mapping for the synchronous calls and the pose ring */
tsdf_status pose_status(const PoseState& ps) {
    if (ps.halo_miss & 0x40000000) { g_err = "sharded tracking: a peer rank did not deliver its normal equations within 2 s; the pose was not updated"; return TSDF_ERR_PEER; }
    if (ps.halo_miss) { g_err = "a tracking sample needed a voxel outside this shard's slab+halo (increase tsdf_config.halo)"; return TSDF_ERR_HALO; }
    if (ps.singular) { g_err = "tracking lost: singular normal equations or non-finite twist"; return TSDF_ERR_TRACKING_LOST; }
    return TSDF_OK;
}
tsdf_status track_result(Impl* p, double R_out[9], double t_out[3], tsdf_track_stats* stats) {
    const PoseState& ps = *p->pose_pin;
    if (R_out) memcpy(R_out, ps.R, sizeof ps.R);
    if (t_out) memcpy(t_out, ps.t, sizeof ps.t);
    fill_stats(ps, stats);
    return pose_status(ps);
}

}  // namespace

extern "C" {

int32_t tsdf_abi_version(void) { return TSDF_ABI_VERSION; }
const char* tsdf_last_error(void) { return g_err.c_str(); }
int32_t tsdf_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

void tsdf_default_config(tsdf_config* c) {
    memset(c, 0, sizeof *c);
    c->m = 256; c->width = 6.0f; c->height = 6.0f; c->depth = 3.5f;      /* sdf_reconstruction.cpp:83-85 */
    c->origin[0] = -3.0; c->origin[1] = -3.0; c->origin[2] = -0.5;
    c->distance_delta = 0.3f; c->distance_epsilon = 0.025f;
    c->gauss_newton_max_iteration = 20; c->maximum_twist_diff = 0.001f;  /* sdf_reconstruction.cpp:88 */
    c->v_h = 1.0f; c->w_h = 0.01f;
    c->pixel_stride = 3;                                                  /* camera_tracking.cpp:162-163 */
    c->metric = TSDF_POINT_TO_PLANE;                                      /* sdf.cpp:272 */
    c->image_width = 640; c->image_height = 480;
    c->device = 0; c->n_shards = 1; c->shard_rank = 0; c->halo = -1;
}

tsdf_status tsdf_create(const tsdf_config* cfg, tsdf_handle* out) {
    if (!cfg || !out) return bad("null argument");
    *out = nullptr;
    if (cfg->m < 8 || cfg->m > 4092 || (cfg->m % 4) != 0) return bad("m must be a multiple of 4 in [8, 4092]");
    if (!(cfg->width > 0 && cfg->height > 0 && cfg->depth > 0)) return bad("extents must be positive");
    if (cfg->image_width < 3 || cfg->image_height < 3 || cfg->image_width > 8192 || cfg->image_height > 8192) return bad("bad image size");
    if (cfg->pixel_stride < 1) return bad("pixel_stride must be >= 1");
    if (cfg->gauss_newton_max_iteration < 0 || cfg->gauss_newton_max_iteration > 1000) return bad("bad iteration count");
    if (cfg->metric != TSDF_POINT_TO_PLANE && cfg->metric != TSDF_POINT_TO_POINT) return bad("bad metric");
    if (cfg->preprocess != 0 && cfg->preprocess != 1) return bad("preprocess must be 0 or 1");
    if (cfg->n_shards < 1 || cfg->n_shards > MAX_WORLD || cfg->shard_rank < 0 || cfg->shard_rank >= cfg->n_shards) return bad("bad shard configuration");
    if (cfg->n_shards > cfg->m / 4) return bad("too many shards for this m");
    if (cfg->slab_k_end > cfg->slab_k_begin) {
        if (cfg->slab_k_begin < 0 || cfg->slab_k_end > cfg->m) return bad("explicit slab outside the grid");
        if ((cfg->shard_rank == 0) != (cfg->slab_k_begin == 0) || (cfg->shard_rank == cfg->n_shards - 1) != (cfg->slab_k_end == cfg->m))
            return bad("explicit slabs must tile [0, m) in rank order");
    } else if (cfg->slab_k_end != 0 || cfg->slab_k_begin != 0) return bad("empty explicit slab");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        g_err = "no CUDA device: libtsdf_b200 has no CPU fallback";
        return TSDF_ERR_CUDA;
    }
    if (cfg->device < 0 || cfg->device >= ndev) return bad("bad device ordinal");
    CK(cudaSetDevice(cfg->device));

    Impl* p = new Impl();
    p->cfg = *cfg;
    p->device = cfg->device;
    build_params(*cfg, p->g);
    int halo;
    slab_range(*cfg, p->g.ko0, p->g.ko1, p->g.ks0, p->g.ks1, halo);
    p->cfg.halo = halo;
    p->n_stored = (int64_t)(p->g.ks1 - p->g.ks0) * cfg->m * cfg->m;
    p->links.world = 1; p->links.rank = 0;
    for (int r = 0; r < MAX_WORLD; r++) p->links.box[r] = nullptr;

    const size_t npx = (size_t)cfg->image_width * cfg->image_height;
    const int P = p->g.ni * p->g.nj;
    cudaError_t e = cudaSuccess;
    auto A = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    A(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    p->stream_refs = new int(1);
    A(cudaMalloc(&p->grid, (size_t)p->n_stored * sizeof(float2)));
    A(cudaStreamCreateWithFlags(&p->prep_stream, cudaStreamNonBlocking));
    for (int q = 0; q < 2; q++) {
        A(cudaMalloc(&p->pix_buf[q], npx * sizeof(PixRec)));
        A(cudaMalloc(&p->pts_buf[q], (size_t)P * sizeof(float4)));
        A(cudaEventCreateWithFlags(&p->prepped[q], cudaEventDisableTiming));
        A(cudaEventCreateWithFlags(&p->rec_free[q], cudaEventDisableTiming));
    }
    A(cudaEventCreateWithFlags(&p->depth_copied, cudaEventDisableTiming));
    p->pix = p->pix_buf[0]; p->pts = p->pts_buf[0];
    A(cudaMalloc(&p->depth_stage, npx * sizeof(float)));
    p->k0 = cfg->preprocess != 0;
    if (p->k0) {
        const size_t sw = (size_t)((float)(cfg->image_width - 1) / 15.0f) + 1 + 2 * K0_PAD, sh = (size_t)((float)(cfg->image_height - 1) / 15.0f) + 1 + 2 * K0_PAD;
        const size_t ncell = sw * sh * K0_SD_MAX;
        A(cudaMalloc(&p->k0b.minmax, 2 * sizeof(unsigned int)));
        A(cudaMalloc(&p->k0b.grid_a, ncell * sizeof(float2)));
        A(cudaMalloc(&p->k0b.grid_b, ncell * sizeof(float2)));
        A(cudaMalloc(&p->k0b.bins, npx * sizeof(short)));
        A(cudaMalloc(&p->k0b.zf, npx * sizeof(float)));
        A(cudaMalloc(&p->k0b.edge, npx));
        A(cudaMalloc(&p->k0b.DX, npx * sizeof(float4)));
        A(cudaMalloc(&p->k0b.DY, npx * sizeof(float4)));
        A(cudaMalloc(&p->k0b.normals, npx * sizeof(float4)));
    }
    A(cudaMalloc(&p->pose_dev, sizeof(PoseState)));
    A(cudaMallocHost(&p->pose_pin, sizeof(PoseState)));
    A(cudaMallocHost(&p->ring_pin, sizeof(PoseState) * POSE_RING));
    A(cudaMalloc(&p->ticket, N_TICKETS * sizeof(unsigned int)));
    A(cudaMalloc(&p->fuse_tables, ((size_t)9 * cfg->m + 8) * sizeof(double)));
    A(cudaMalloc(&p->fuse_items, (size_t)(p->g.ks1 - p->g.ks0) * cfg->m * ((cfg->m + 127) / 128 + 1) * sizeof(unsigned long long)));
    A(cudaMalloc(&p->fuse_item_c, (size_t)(p->g.ks1 - p->g.ks0) * cfg->m * ((cfg->m + 127) / 128 + 1) * sizeof(float4)));
    A(cudaMalloc(&p->fuse_item_count, 4 * sizeof(unsigned int)));      /* [0] items, [1] units, [2] items of row-certified free space */
    A(cudaMalloc(&p->fuse_units, (size_t)(p->n_stored / 4 + 64) * sizeof(unsigned long long)));
    {
        int64_t off = 0;
        for (int l = 0; l < CERT_LEVELS; l++) {
            p->pyr.w[l] = (cfg->image_width + (1 << l) - 1) >> l;
            p->pyr.h[l] = (cfg->image_height + (1 << l) - 1) >> l;
            p->pyr.off[l] = off;
            off += (int64_t)p->pyr.w[l] * p->pyr.h[l];
        }
        for (int q = 0; q < 2; q++) A(cudaMalloc(&p->cert_buf[q], (size_t)off * sizeof(float2)));
        p->cert = p->cert_buf[0];
    }
    A(cudaMalloc(&p->n_upd_dev, 4 * sizeof(unsigned long long)));
    A(cudaMallocHost(&p->n_upd_pin, sizeof(unsigned long long)));
    A(cudaMalloc(&p->dbgJ, (size_t)P * 6 * sizeof(float)));
    A(cudaMalloc(&p->dbgPsi, (size_t)P * sizeof(float)));
    A(cudaMalloc(&p->dbgFlag, (size_t)P));
    A(cudaMalloc(&p->mailbox, sizeof(Mailbox)));
    A(cudaMalloc(&p->scratch_d, 64 * sizeof(double)));
    A(cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking));
    for (int q = 0; q < Impl::NSTAGE; q++) {
        A(cudaMalloc(&p->stage[q], npx * sizeof(float)));
        A(cudaEventCreateWithFlags(&p->copied[q], cudaEventDisableTiming));
        A(cudaEventCreateWithFlags(&p->consumed[q], cudaEventDisableTiming));
    }
    for (int q = 0; q < 4; q++) A(cudaEventCreate(&p->ev[q]));
    for (int q = 0; q < 2; q++) A(cudaEventCreate(&p->tmr[q]));
    if (e != cudaSuccess) {
        g_err = std::string("allocation failed: ") + cudaGetErrorString(e);
        tsdf_destroy(reinterpret_cast<tsdf_handle>(p));
        return e == cudaErrorMemoryAllocation ? TSDF_ERR_NOMEM : TSDF_ERR_CUDA;
    }
    cudaMemset(p->ticket, 0, N_TICKETS * sizeof(unsigned int));
    {
        const char* force64 = getenv("TSDF_B200_IDX64");       /* tests: exercise the 64-bit voxel index path on small stores */
        p->force_idx64 = (force64 && force64[0] == '1') ? 1 : 0;
    }
    cudaMemset(p->n_upd_dev, 0, 4 * sizeof(unsigned long long));
    cudaMemset(p->mailbox, 0, sizeof(Mailbox));
    p->links.box[0] = p->mailbox;

    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, cfg->device));
    const int sms = prop.multiProcessorCount;
    /* K2: one block per LIN_TW x LIN_TH tile of the strided pixel grid; LIN_MICRO: one block per SM slot, each with a
     * contiguous run of 4 x 4 micro-tiles (tsdf_kernels.cu: k_linearize, "Work distribution") */
#ifdef LIN_MICRO
    const int n_micro = ((p->g.ni + 3) / 4) * ((p->g.nj + 3) / 4);
    int lb = linearize_blocks_per_sm();
    if (lb < 1) lb = 1;
    const int nb = n_micro < sms * lb ? n_micro : sms * lb;
    const int ppb = ((n_micro + nb - 1) / nb) * 16;
#else
    const int ppb = LIN_TW * LIN_TH;
    const int nb = ((p->g.ni + LIN_TW - 1) / LIN_TW) * ((p->g.nj + LIN_TH - 1) / LIN_TH);
#endif
    if ((nb + LIN_GROUP - 1) / LIN_GROUP > MAX_LIN_GROUPS) { g_err = "image too large for the reduction tree"; tsdf_destroy(reinterpret_cast<tsdf_handle>(p)); return TSDF_ERR_BAD_ARG; }
    p->px_per_block = ppb; p->lin_blocks = nb;
    A(cudaMalloc(&p->partials, (size_t)nb * LIN_PARTIAL_STRIDE * sizeof(double)));
    A(cudaMalloc(&p->group_partials, (size_t)((nb + LIN_GROUP - 1) / LIN_GROUP) * LIN_PARTIAL_STRIDE * sizeof(double)));
    /* K3: persistent grid, a whole number of resident blocks per SM */
    int fb = fuse_blocks_per_sm();
    if (fb < 1) fb = 1;
    p->fuse_blocks = sms * fb;
    int cb = fuse_cert_blocks_per_sm();
    if (cb < 1) cb = 1;
    p->fuse_cert_blocks = sms * cb;
    if (e != cudaSuccess) { g_err = "allocation failed"; tsdf_destroy(reinterpret_cast<tsdf_handle>(p)); return TSDF_ERR_NOMEM; }

    *out = reinterpret_cast<tsdf_handle>(p);
    tsdf_status st = tsdf_reset(*out);
    if (st != TSDF_OK) { tsdf_destroy(*out); *out = nullptr; return st; }
    return TSDF_OK;
}

tsdf_status tsdf_destroy(tsdf_handle h) {
    if (!h) return TSDF_OK;
    Impl* p = I(h);
    cudaSetDevice(p->device);
    if (p->prep_stream) cudaStreamSynchronize(p->prep_stream);
    if (p->stream) cudaStreamSynchronize(p->stream);
    for (void* q : p->ipc_opened) cudaIpcCloseMemHandle(q);
    for (int q = 0; q < 2; q++) {
        cudaFree(p->pix_buf[q]); cudaFree(p->pts_buf[q]); cudaFree(p->cert_buf[q]);
        if (p->prepped[q]) cudaEventDestroy(p->prepped[q]);
        if (p->rec_free[q]) cudaEventDestroy(p->rec_free[q]);
    }
    if (p->depth_copied) cudaEventDestroy(p->depth_copied);
    if (p->prep_stream) cudaStreamDestroy(p->prep_stream);
    cudaFree(p->mc_count); cudaFree(p->mc_off); cudaFree(p->mc_tmp); cudaFree(p->mesh_xyz); cudaFree(p->mc_cells); cudaFree(p->mc_cell_counter);
    cudaFree(p->color); cudaFree(p->rgb4_buf[0]); cudaFree(p->rgb4_buf[1]); cudaFree(p->rgb_stage);
    cudaFree(p->cosn_buf[0]); cudaFree(p->cosn_buf[1]);
    cudaFree(p->k0b.minmax); cudaFree(p->k0b.grid_a); cudaFree(p->k0b.grid_b); cudaFree(p->k0b.bins); cudaFree(p->k0b.zf); cudaFree(p->k0b.edge);
    cudaFree(p->k0b.DX); cudaFree(p->k0b.DY); cudaFree(p->k0b.normals);
    cudaFree(p->grid); cudaFree(p->depth_stage); cudaFree(p->pose_dev);
    cudaFreeHost(p->pose_pin); cudaFreeHost(p->ring_pin); cudaFree(p->partials); cudaFree(p->ticket);
    cudaFree(p->group_partials); cudaFree(p->fuse_tables); cudaFree(p->fuse_items); cudaFree(p->fuse_item_c); cudaFree(p->fuse_item_count);
    cudaFree(p->fuse_units);
    cudaFree(p->n_upd_dev); cudaFreeHost(p->n_upd_pin);
    cudaFree(p->dbgJ); cudaFree(p->dbgPsi); cudaFree(p->dbgFlag); cudaFree(p->mailbox);
    cudaFree(p->flush_buf); cudaFree(p->scratch_d);
    for (cudaEvent_t e : p->ring_ev) cudaEventDestroy(e);
    if (p->copy_stream) { cudaStreamSynchronize(p->copy_stream); cudaStreamDestroy(p->copy_stream); }
    for (int q = 0; q < Impl::NSTAGE; q++) {
        cudaFree(p->stage[q]);
        if (p->copied[q]) cudaEventDestroy(p->copied[q]);
        if (p->consumed[q]) cudaEventDestroy(p->consumed[q]);
    }
    for (int q = 0; q < 4; q++) if (p->ev[q]) cudaEventDestroy(p->ev[q]);
    for (int q = 0; q < 2; q++) if (p->tmr[q]) cudaEventDestroy(p->tmr[q]);
    if (p->stream_refs && --*p->stream_refs == 0) {           /* last handle on this stream */
        if (p->stream) cudaStreamDestroy(p->stream);
        delete p->stream_refs;
    }
    cudaGetLastError();
    delete p;
    return TSDF_OK;
}

tsdf_status tsdf_reset(tsdf_handle h) {
    if (!h) return bad("null handle");
    Impl* p = I(h);
    CK(cudaSetDevice(p->device));
    const float d0 = p->cfg.width + p->cfg.height + p->cfg.depth;        /* sdf.cpp:29 */
    launch_fill(p->grid, p->n_stored, d0, p->stream);
    p->launches++;
    if (p->color) { launch_fill_color(p->color, p->n_stored, p->stream); p->launches++; }   /* sdf.cpp:30-34 */
    CK(cudaGetLastError());
    /* camera_tracking.cpp:5-8 initial pose */
    const double R0[9] = {1, 0, 0, 0, 0, -1, 0, -1, 0};
    const double t0[3] = {0, 0, 1};
    PoseState ps;
    memset(&ps, 0, sizeof ps);
    pose_set(ps, R0, t0);
    return upload_pose(p, ps);
}

tsdf_status tsdf_get_config(tsdf_handle h, tsdf_config* cfg) {
    if (!h || !cfg) return bad("null argument");
    *cfg = I(h)->cfg;
    return TSDF_OK;
}

tsdf_status tsdf_set_intrinsics(tsdf_handle h, const double K[9]) {
    if (!h || !K) return bad("null argument");
    Impl* p = I(h);
    for (int q = 0; q < 9; q++) p->g.K[q] = K[q];                        /* camera_tracking.cpp:24-32 */
    p->g.k_simple = (K[1] == 0.0 && K[3] == 0.0 && K[6] == 0.0 && K[7] == 0.0 && K[8] == 1.0) ? 1 : 0;
    p->have_K = true;
    return TSDF_OK;
}

tsdf_status tsdf_set_pose(tsdf_handle h, const double R[9], const double t[3]) {
    if (!h || !R || !t) return bad("null argument");
    Impl* p = I(h);
    CK(cudaSetDevice(p->device));
    tsdf_status st = fetch_pose(p);
    if (st != TSDF_OK) return st;
    PoseState ps = *p->pose_pin;
    pose_set(ps, R, t);                                                   /* camera_tracking.cpp:59-65 */
    return upload_pose(p, ps);
}
tsdf_status tsdf_get_pose(tsdf_handle h, double R[9], double t[3]) {
    if (!h) return bad("null handle");
    Impl* p = I(h);
    CK(cudaSetDevice(p->device));
    tsdf_status st = fetch_pose(p);
    if (st != TSDF_OK) return st;
    if (R) memcpy(R, p->pose_pin->R, sizeof p->pose_pin->R);
    if (t) memcpy(t, p->pose_pin->t, sizeof p->pose_pin->t);
    return TSDF_OK;
}
tsdf_status tsdf_get_pose_inv(tsdf_handle h, double Rinv[9], double tinv[3]) {
    if (!h) return bad("null handle");
    Impl* p = I(h);
    CK(cudaSetDevice(p->device));
    tsdf_status st = fetch_pose(p);
    if (st != TSDF_OK) return st;
    if (Rinv) memcpy(Rinv, p->pose_pin->Rinv, sizeof p->pose_pin->Rinv);
    if (tinv) memcpy(tinv, p->pose_pin->tinv, sizeof p->pose_pin->tinv);
    return TSDF_OK;
}

tsdf_status tsdf_track(tsdf_handle h, const float* depth, int32_t mem, double R_out[9], double t_out[3], tsdf_track_stats* stats) {
    if (!h) return bad("null handle");
    Impl* p = I(h);
    CK(cudaSetDevice(p->device));
    const float* dptr;
    tsdf_status st = stage_depth(p, depth, mem, &dptr);
    if (st != TSDF_OK) return st;
    st = enqueue_frame(p, dptr, true, false);
    if (st != TSDF_OK) return st;
    st = fetch_pose(p);
    if (st != TSDF_OK) return st;
    return track_result(p, R_out, t_out, stats);
}

tsdf_status tsdf_fuse(tsdf_handle h, const float* depth, int32_t mem, const double R[9], const double t[3], int64_t* n_updated) {
    if (!h) return bad("null handle");
    if ((R == nullptr) != (t == nullptr)) return bad("R and t must both be given or both be NULL");
    Impl* p = I(h);
    CK(cudaSetDevice(p->device));
    if (!p->have_K) { g_err = "camera matrix not set (tsdf_set_intrinsics)"; return TSDF_ERR_NO_INTRINSICS; }
    tsdf_status st;
    if (R) { st = tsdf_set_pose(h, R, t); if (st != TSDF_OK) return st; }
    const float* dptr;
    st = stage_depth(p, depth, mem, &dptr);
    if (st != TSDF_OK) return st;
    st = enqueue_frame(p, dptr, false, true);
    if (st != TSDF_OK) return st;
    CK(cudaMemcpyAsync(p->n_upd_pin, p->n_upd_dev, sizeof(unsigned long long), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    if (n_updated) *n_updated = (int64_t)*p->n_upd_pin;
    return TSDF_OK;
}

tsdf_status tsdf_track_and_fuse(tsdf_handle h, const float* depth, int32_t mem, double R_out[9], double t_out[3],
                                tsdf_track_stats* stats, int64_t* n_updated) {
    if (!h) return bad("null handle");
    Impl* p = I(h);
    CK(cudaSetDevice(p->device));
    const float* dptr;
    tsdf_status st = stage_depth(p, depth, mem, &dptr);
    if (st != TSDF_OK) return st;
    st = enqueue_frame(p, dptr, true, true);
    if (st != TSDF_OK) return st;
    CK(cudaMemcpyAsync(p->n_upd_pin, p->n_upd_dev, sizeof(unsigned long long), cudaMemcpyDeviceToHost, p->stream));
    st = fetch_pose(p);
    if (st != TSDF_OK) return st;
    if (n_updated) *n_updated = (int64_t)*p->n_upd_pin;
    return track_result(p, R_out, t_out, stats);
}

tsdf_status tsdf_enable_color(tsdf_handle h) {
    if (!h) return bad("null handle");
    Impl* p = I(h);
    CK(cudaSetDevice(p->device));
    if (p->color) return TSDF_OK;
    if (p->g.metric != 0) return bad("colour fusion needs the point-to-plane metric (sdf.cpp:294 uses the normal)");
    const size_t npx = (size_t)p->g.img_w * p->g.img_h;
    cudaError_t e = cudaMalloc(&p->color, (size_t)p->n_stored * sizeof(float4));
    for (int q = 0; q < 2 && e == cudaSuccess; q++) e = cudaMalloc(&p->rgb4_buf[q], npx * sizeof(uchar4));
    for (int q = 0; q < 2 && e == cudaSuccess; q++) e = cudaMalloc(&p->cosn_buf[q], npx * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&p->rgb_stage, npx * 3);
    if (e != cudaSuccess) {
        cudaFree(p->color); p->color = nullptr;
        for (int q = 0; q < 2; q++) { cudaFree(p->rgb4_buf[q]); p->rgb4_buf[q] = nullptr; cudaFree(p->cosn_buf[q]); p->cosn_buf[q] = nullptr; }
        cudaFree(p->rgb_stage); p->rgb_stage = nullptr;
        cudaGetLastError();
        g_err = std::string("colour store allocation failed: ") + cudaGetErrorString(e);
        return TSDF_ERR_NOMEM;
    }
    int cb = fuse_color_blocks_per_sm();
    if (cb < 1) cb = 1;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, p->device));
    p->fuse_color_blocks = prop.multiProcessorCount * cb;
    launch_fill_color(p->color, p->n_stored, p->stream);
    p->launches++;
    CK(cudaStreamSynchronize(p->stream));
    return TSDF_OK;
}

tsdf_status tsdf_fuse_rgb(tsdf_handle h, const float* depth, const uint8_t* rgb, int32_t mem,
                          const double R[9], const double t[3], int64_t* n_updated) {
    if (!h) return bad("null handle");
    if ((R == nullptr) != (t == nullptr)) return bad("R and t must both be given or both be NULL");
    Impl* p = I(h);
    CK(cudaSetDevice(p->device));
    if (!p->have_K) { g_err = "camera matrix not set (tsdf_set_intrinsics)"; return TSDF_ERR_NO_INTRINSICS; }
    tsdf_status st = tsdf_enable_color(h);
    if (st != TSDF_OK) return st;
    if (R) { st = tsdf_set_pose(h, R, t); if (st != TSDF_OK) return st; }
    const float* dptr;
    st = stage_depth(p, depth, mem, &dptr);
    if (st != TSDF_OK) return st;
    st = stage_rgb(p, rgb, mem);
    if (st != TSDF_OK) return st;
    st = enqueue_frame(p, dptr, false, true);
    if (st != TSDF_OK) return st;
    CK(cudaMemcpyAsync(p->n_upd_pin, p->n_upd_dev, sizeof(unsigned long long), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    if (n_updated) *n_updated = (int64_t)*p->n_upd_pin;
    return TSDF_OK;
}

tsdf_status tsdf_track_and_fuse_rgb(tsdf_handle h, const float* depth, const uint8_t* rgb, int32_t mem,
                                    double R_out[9], double t_out[3], tsdf_track_stats* stats, int64_t* n_updated) {
    if (!h) return bad("null handle");
    Impl* p = I(h);
    CK(cudaSetDevice(p->device));
    tsdf_status st = tsdf_enable_color(h);
    if (st != TSDF_OK) return st;
    const float* dptr;
    st = stage_depth(p, depth, mem, &dptr);
    if (st != TSDF_OK) return st;
    st = stage_rgb(p, rgb, mem);
    if (st != TSDF_OK) return st;
    st = enqueue_frame(p, dptr, true, true);
    if (st != TSDF_OK) return st;
    CK(cudaMemcpyAsync(p->n_upd_pin, p->n_upd_dev, sizeof(unsigned long long), cudaMemcpyDeviceToHost, p->stream));
    st = fetch_pose(p);
    if (st != TSDF_OK) return st;
    if (n_updated) *n_updated = (int64_t)*p->n_upd_pin;
    return track_result(p, R_out, t_out, stats);
}

tsdf_status tsdf_interpolate_color(tsdf_handle h, int64_t n, const double* global_pts, float* rgba) {
    if (!h || !global_pts || !rgba || n < 0) return bad("bad argument");
    Impl* p = I(h);
    if (!p->color) return bad("colour store not allocated (tsdf_enable_color)");
    if (n == 0) return TSDF_OK;
    CK(cudaSetDevice(p->device));
    DevTmp dp, dout;
    CK(dp.alloc((size_t)n * 3 * sizeof(double)));
    CK(dout.alloc((size_t)n * 4 * sizeof(float)));
    CK(cudaMemcpyAsync(dp.p, global_pts, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    launch_sample_color(p->g, p->color, n, dp.as<double>(), dout.as<float>(), p->stream);
    p->launches++;
    CK(cudaMemcpyAsync(rgba, dout.p, (size_t)n * 4 * sizeof(float), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return TSDF_OK;
}

tsdf_status tsdf_download_color(tsdf_handle h, float* cw, float* r, float* g, float* b, int32_t layout) {
    if (!h || !cw || !r || !g || !b) return bad("null argument");
    if (layout != TSDF_LAYOUT_REFERENCE && layout != TSDF_LAYOUT_XFASTEST) return bad("bad layout");
    Impl* p = I(h);
    if (!p->color) return bad("colour store not allocated (tsdf_enable_color)");
    CK(cudaSetDevice(p->device));
    DevTmp d[4];
    float* hst[4] = {cw, r, g, b};
    for (int q = 0; q < 4; q++) CK(d[q].alloc((size_t)p->n_stored * sizeof(float)));
    launch_export_color(p->g, p->color, d[0].as<float>(), d[1].as<float>(), d[2].as<float>(), d[3].as<float>(), layout == TSDF_LAYOUT_REFERENCE ? 1 : 0, p->stream);
    p->launches++;
    for (int q = 0; q < 4; q++) CK(cudaMemcpyAsync(hst[q], d[q].p, (size_t)p->n_stored * sizeof(float), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    CK(cudaGetLastError());
    return TSDF_OK;
}

/* cost of slab [a, b) = weight of every layer it FUSES (its own layers plus `halo` on each side: weights) plus the
 * weight of what it alone does for its OWN layers (weights_own, may be NULL: e.g. the tracked pixels it owns) */
tsdf_status tsdf_balanced_slabs2(int32_t m, int32_t n_shards, const double* weights, const double* weights_own,
                                 int32_t min_layers, int32_t halo, int32_t* bounds) {
    if (!weights || !bounds || m < 1 || n_shards < 1) return bad("bad argument");
    if (min_layers < 1) min_layers = 1;
    if (halo < 0) halo = 0;
    if ((int64_t)min_layers * n_shards > m) return bad("min_layers * n_shards exceeds m");
    std::vector<double> P((size_t)m + 1, 0.0), Q((size_t)m + 1, 0.0);
    for (int k = 0; k < m; k++) {
        if (!(weights[k] >= 0.0) || !(weights[k] <= 1e300)) return bad("weights must be finite and non-negative");
        if (weights_own && (!(weights_own[k] >= 0.0) || !(weights_own[k] <= 1e300))) return bad("weights must be finite and non-negative");
        P[k + 1] = P[k] + weights[k];
        Q[k + 1] = Q[k] + (weights_own ? weights_own[k] : 0.0);
    }
    auto cost = [&](int a, int b) { return (P[b + halo > m ? m : b + halo] - P[a - halo < 0 ? 0 : a - halo]) + (Q[b] - Q[a]); };
    /* exact: dynamic programme over the cut positions, f[r][b] = min over a <= b - min_layers of max(f[r-1][a], cost(a, b)).
     * (With a minimum thickness the classic greedy "as thick as the cap allows" is not optimal: the cost of the thinnest
     * admissible slab is a sliding window, not monotone in its start.)  O(n_shards * m^2) on the host, once per run. */
    const double INF = 1e308;
    std::vector<double> prev((size_t)m + 1, INF), cur((size_t)m + 1, INF);
    std::vector<std::vector<int32_t>> arg((size_t)n_shards, std::vector<int32_t>((size_t)m + 1, -1));
    for (int bq = min_layers; bq <= m; bq++) { prev[bq] = cost(0, bq); arg[0][bq] = 0; }
    for (int r = 1; r < n_shards; r++) {
        std::fill(cur.begin(), cur.end(), INF);
        for (int bq = (r + 1) * min_layers; bq <= m; bq++) {
            double best = INF; int besta = -1;
            for (int a = r * min_layers; a <= bq - min_layers; a++) {
                if (prev[a] >= best) continue;
                const double c = cost(a, bq);
                const double v = prev[a] > c ? prev[a] : c;
                if (v < best) { best = v; besta = a; }
            }
            cur[bq] = best; arg[r][bq] = besta;
        }
        prev.swap(cur);
    }
    if (!(prev[m] < INF)) return bad("no partition satisfies min_layers");
    int bq = m;
    bounds[n_shards] = m;
    for (int r = n_shards - 1; r >= 0; r--) { bq = arg[r][bq]; bounds[r] = bq; }
    return TSDF_OK;
}
tsdf_status tsdf_balanced_slabs(int32_t m, int32_t n_shards, const double* weights, int32_t min_layers, int32_t halo, int32_t* bounds) {
    return tsdf_balanced_slabs2(m, n_shards, weights, nullptr, min_layers, halo, bounds);
}

tsdf_status tsdf_mesh_extract(tsdf_handle h, float iso_level, int64_t* n_vertices) {
    if (!h) return bad("null handle");
    Impl* p = I(h);
    CK(cudaSetDevice(p->device));
    p->mesh_n = 0;
    if (n_vertices) *n_vertices = 0;
    if (!(iso_level >= 0.0f && iso_level < 1.0f)) return TSDF_OK;         /* marching_cubes_sdf.cpp:248-254: empty cloud */
    const int64_t n_rows = (int64_t)p->g.m * p->g.m * mesh_zsplit() + 1;
    if (!p->mc_count) {
        p->mc_tmp_bytes = mesh_scan_bytes(n_rows);
        CK(cudaMalloc(&p->mc_count, (size_t)n_rows * sizeof(unsigned int)));
        CK(cudaMalloc(&p->mc_off, (size_t)n_rows * sizeof(unsigned int)));
        CK(cudaMalloc(&p->mc_tmp, p->mc_tmp_bytes ? p->mc_tmp_bytes : 16));
        CK(cudaMalloc(&p->mc_cell_counter, sizeof(unsigned int)));
    }
    if (!p->mc_cells) {
        p->mc_cell_cap = 1u << 20;                                    /* grows to fit the surface, see below */
        const char* cap_env = getenv("TSDF_B200_MC_CELLS");           /* tests: a tiny list forces the overflow fallback */
        if (cap_env && atoi(cap_env) > 0) p->mc_cell_cap = (unsigned int)atoi(cap_env);
        CK(cudaMalloc(&p->mc_cells, (size_t)p->mc_cell_cap * sizeof(unsigned long long)));
    }
    McParams P;
    P.width = p->cfg.width; P.height = p->cfg.height; P.depth = p->cfg.depth; P.iso = iso_level;
    launch_mesh_count(p->g, P, p->grid, p->mc_count, p->mc_off, p->mc_tmp, p->mc_tmp_bytes, p->mc_cells, p->mc_cell_counter, p->mc_cell_cap, p->stream);
    p->launches += 2;
    /* The emit is enqueued right behind the sweep, before the host knows the counts: it takes the list length from the
     * device and stays inside the capacities of this moment.  One synchronisation per extraction; the emit is repeated
     * (with larger buffers) only when the surface outgrew them. */
    const unsigned int cell_cap0 = p->mc_cell_cap;
    const int64_t vtx_cap0 = p->mesh_cap;
    const bool speculated = p->mesh_xyz != nullptr && vtx_cap0 > 0 && cell_cap0 > 0;
    if (speculated) {
        /* threads: the last list length with head room (the surface grows slowly), at most the list capacity */
        unsigned int nt = p->mc_last_cells ? p->mc_last_cells + p->mc_last_cells / 8 + 4096u : cell_cap0;
        if (nt > cell_cap0) nt = cell_cap0;
        p->mc_emit_threads = nt;
        launch_mesh_emit_list(p->g, P, p->grid, p->mc_cells, p->mc_cell_counter, nt, p->mc_off, p->mesh_xyz, cell_cap0,
                              (unsigned int)(vtx_cap0 > 0xffffffffll ? 0xffffffffll : vtx_cap0), p->stream);
        p->launches++;
    }
    unsigned int total = 0, n_cells = 0;
    CK(cudaMemcpyAsync(&total, p->mc_off + (n_rows - 1), sizeof total, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaMemcpyAsync(&n_cells, p->mc_cell_counter, sizeof n_cells, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    p->mc_last_cells = n_cells;
    const bool emitted = speculated && (int64_t)total <= vtx_cap0 && n_cells <= cell_cap0 && n_cells <= p->mc_emit_threads;
    if (total && !emitted) {
        if ((int64_t)total > p->mesh_cap) {                          /* grow with head room: the next mesh is usually a bit larger */
            cudaFree(p->mesh_xyz); p->mesh_xyz = nullptr; p->mesh_cap = 0;
            const int64_t cap = (int64_t)total + (int64_t)total / 4 + 1024;
            cudaError_t e = cudaMalloc(&p->mesh_xyz, (size_t)cap * 3 * sizeof(float));
            if (e != cudaSuccess) { cudaGetLastError(); p->mesh_xyz = nullptr; g_err = "mesh buffer allocation failed"; return TSDF_ERR_NOMEM; }
            p->mesh_cap = cap;
        }
        if (n_cells <= p->mc_cell_cap) {
            /* one sweep: the triangles come from the listed surface cells */
            launch_mesh_emit_list(p->g, P, p->grid, p->mc_cells, p->mc_cell_counter, n_cells, p->mc_off, p->mesh_xyz, p->mc_cell_cap,
                                  (unsigned int)(p->mesh_cap > 0xffffffffll ? 0xffffffffll : p->mesh_cap), p->stream);
        } else {
            /* the list overflowed: second sweep over the store this time, a larger list for the next extraction */
            launch_mesh_emit(p->g, P, p->grid, p->mc_off, p->mesh_xyz, p->stream);
            cudaFree(p->mc_cells); p->mc_cells = nullptr;
            const unsigned int cap = n_cells + n_cells / 2 + 1024u;
            if (cudaMalloc(&p->mc_cells, (size_t)cap * sizeof(unsigned long long)) == cudaSuccess) p->mc_cell_cap = cap;
            else { cudaGetLastError(); p->mc_cells = nullptr; p->mc_cell_cap = 0; }
        }
        p->launches++;
        CK(cudaStreamSynchronize(p->stream));
    }
    CK(cudaGetLastError());
    p->mesh_n = (int64_t)total;
    if (n_vertices) *n_vertices = p->mesh_n;
    return TSDF_OK;
}

tsdf_status tsdf_mesh_download(tsdf_handle h, float* xyz, double* world, float* rgba) {
    if (!h) return bad("null handle");
    Impl* p = I(h);
    CK(cudaSetDevice(p->device));
    const int64_t n = p->mesh_n;
    if (n == 0) return TSDF_OK;
    if (rgba && !p->color) return bad("vertex colours need the colour store (tsdf_enable_color)");
    if (xyz) CK(cudaMemcpyAsync(xyz, p->mesh_xyz, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, p->stream));
    if (world || rgba) {
        DevTmp dw, dc;
        CK(dw.alloc((size_t)n * 3 * sizeof(double)));
        launch_mesh_world(p->g, p->mesh_xyz, n, dw.as<double>(), p->stream);
        p->launches++;
        if (world) CK(cudaMemcpyAsync(world, dw.p, (size_t)n * 3 * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
        if (rgba) {
            CK(dc.alloc((size_t)n * 4 * sizeof(float)));
            launch_sample_color(p->g, p->color, n, dw.as<double>(), dc.as<float>(), p->stream);
            p->launches++;
            CK(cudaMemcpyAsync(rgba, dc.p, (size_t)n * 4 * sizeof(float), cudaMemcpyDeviceToHost, p->stream));
        }
        CK(cudaStreamSynchronize(p->stream));
    }
    CK(cudaStreamSynchronize(p->stream));
    CK(cudaGetLastError());
    return TSDF_OK;
}

int32_t tsdf_pose_ring_capacity(void) { return POSE_RING; }

tsdf_status tsdf_enqueue_frame(tsdf_handle h, const float* depth_dev, int32_t track, int32_t slot) {
    if (!h || !depth_dev) return bad("null argument");
    if (slot < 0 || slot >= POSE_RING) return bad("slot out of range");
    Impl* p = I(h);
    CK(cudaSetDevice(p->device));
    tsdf_status st = enqueue_frame(p, depth_dev, track != 0, true, false);
    if (st != TSDF_OK) return st;
    CK(cudaMemcpyAsync(&p->ring_pin[slot], p->pose_dev, sizeof(PoseState), cudaMemcpyDeviceToHost, p->stream));
    return TSDF_OK;
}
/* Streaming with HOST buffers: the H2D copy of frame n+1 runs on a copy stream while frame n is
 * being tracked and fused; poses come back through the pinned ring like tsdf_enqueue_frame. */
tsdf_status tsdf_submit_frame(tsdf_handle h, const float* depth_host, int32_t track, int32_t slot) {
    if (!h || !depth_host) return bad("null argument");
    if (slot < 0 || slot >= POSE_RING) return bad("slot out of range");
    Impl* p = I(h);
    CK(cudaSetDevice(p->device));
    if (!p->have_K) { g_err = "camera matrix not set (tsdf_set_intrinsics)"; return TSDF_ERR_NO_INTRINSICS; }
    const int b = (int)(p->submit_seq % Impl::NSTAGE);
    const size_t bytes = (size_t)p->g.img_w * p->g.img_h * sizeof(float);
    if (p->submit_seq >= (unsigned long long)Impl::NSTAGE) {
        /* HOST wait for the H2D copy issued NSTAGE submissions ago (same stage buffer): when this call returns,
         * that call's host buffer has been read — the documented lifetime ("until 4 later submissions returned") */
        CK(cudaEventSynchronize(p->copied[b]));
        CK(cudaStreamWaitEvent(p->copy_stream, p->consumed[b], 0));
    }
    CK(cudaMemcpyAsync(p->stage[b], depth_host, bytes, cudaMemcpyHostToDevice, p->copy_stream));
    CK(cudaEventRecord(p->copied[b], p->copy_stream));
    p->depth_ready = p->copied[b];               /* K1 waits for the copy and releases the stage buffer itself */
    p->depth_done = p->consumed[b];
    tsdf_status st = enqueue_frame(p, p->stage[b], track != 0, true, false);
    if (st != TSDF_OK) return st;
    CK(cudaMemcpyAsync(&p->ring_pin[slot], p->pose_dev, sizeof(PoseState), cudaMemcpyDeviceToHost, p->stream));
    p->submit_seq++;
    return TSDF_OK;
}

tsdf_status tsdf_sync(tsdf_handle h) {
    if (!h) return bad("null handle");
    Impl* p = I(h);
    CK(cudaSetDevice(p->device));
    CK(cudaStreamSynchronize(p->stream));
    return TSDF_OK;
}
tsdf_status tsdf_read_pose_ring(tsdf_handle h, int32_t slot, double R[9], double t[3], tsdf_track_stats* stats) {
    if (!h) return bad("null handle");
    if (slot < 0 || slot >= POSE_RING) return bad("slot out of range");
    const PoseState& ps = I(h)->ring_pin[slot];
    if (R) memcpy(R, ps.R, sizeof ps.R);
    if (t) memcpy(t, ps.t, sizeof ps.t);
    fill_stats(ps, stats);
    return pose_status(ps);                 /* same mapping as tsdf_track: lost / halo / peer timeout are reported here too */
}

tsdf_status tsdf_linearize(tsdf_handle h, const float* depth, int32_t mem, double A[36], double b[6], tsdf_track_stats* stats) {
    if (!h) return bad("null handle");
    Impl* p = I(h);
    CK(cudaSetDevice(p->device));
    if (!p->have_K) { g_err = "camera matrix not set"; return TSDF_ERR_NO_INTRINSICS; }
    if (p->exchange_mode == 2) return bad("same-device shard group: use tsdf_group_* entry points");
    const float* dptr;
    tsdf_status st = stage_depth(p, depth, mem, &dptr);
    if (st != TSDF_OK) return st;
    enqueue_prep(p, dptr, 1);
    enqueue_linearize(p, 0, false);
    st = launch_status();
    if (st != TSDF_OK) return st;
    st = fetch_pose(p);
    if (st != TSDF_OK) return st;
    tsdf_track_stats s;
    fill_stats(*p->pose_pin, &s);
    if (A) memcpy(A, s.A, sizeof s.A);
    if (b) memcpy(b, s.b, sizeof s.b);
    if (stats) *stats = s;
    if (p->pose_pin->halo_miss) { g_err = "halo too small for the tracking stencil"; return TSDF_ERR_HALO; }
    return TSDF_OK;
}

int32_t tsdf_num_strided_pixels(tsdf_handle h) { return h ? I(h)->g.ni * I(h)->g.nj : 0; }

tsdf_status tsdf_linearize_pixels(tsdf_handle h, const float* depth, int32_t mem, float* J, float* psi, uint8_t* flag) {
    if (!h) return bad("null handle");
    Impl* p = I(h);
    CK(cudaSetDevice(p->device));
    if (!p->have_K) { g_err = "camera matrix not set"; return TSDF_ERR_NO_INTRINSICS; }
    if (p->links.world > 1) return bad("per-pixel records are a single-shard debugging aid");
    const float* dptr;
    tsdf_status st = stage_depth(p, depth, mem, &dptr);
    if (st != TSDF_OK) return st;
    const int P = p->g.ni * p->g.nj;
    CK(cudaMemsetAsync(p->dbgFlag, 0, P, p->stream));
    CK(cudaMemsetAsync(p->dbgJ, 0, (size_t)P * 6 * sizeof(float), p->stream));
    CK(cudaMemsetAsync(p->dbgPsi, 0, (size_t)P * sizeof(float), p->stream));
    enqueue_prep(p, dptr, 1);
    enqueue_linearize(p, 0, true);
    st = launch_status();
    if (st != TSDF_OK) return st;
    if (J) CK(cudaMemcpyAsync(J, p->dbgJ, (size_t)P * 6 * sizeof(float), cudaMemcpyDeviceToHost, p->stream));
    if (psi) CK(cudaMemcpyAsync(psi, p->dbgPsi, (size_t)P * sizeof(float), cudaMemcpyDeviceToHost, p->stream));
    if (flag) CK(cudaMemcpyAsync(flag, p->dbgFlag, (size_t)P, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return TSDF_OK;
}

tsdf_status tsdf_backproject(tsdf_handle h, const float* depth, int32_t mem, float* cloud, float* normals) {
    if (!h || !cloud) return bad("null argument");
    Impl* p = I(h);
    CK(cudaSetDevice(p->device));
    if (!p->have_K) { g_err = "camera matrix not set"; return TSDF_ERR_NO_INTRINSICS; }
    const float* dptr;
    tsdf_status st = stage_depth(p, depth, mem, &dptr);
    if (st != TSDF_OK) return st;
    const size_t n = (size_t)p->g.img_w * p->g.img_h * 3;
    DevTmp dc, dn;
    CK(dc.alloc(n * sizeof(float)));
    if (normals) CK(dn.alloc(n * sizeof(float)));
    enqueue_prep(p, dptr, 0);
    launch_cloud(p->g, p->pix, dc.as<float>(), normals ? dn.as<float>() : nullptr, p->stream);
    p->launches++;
    CK(cudaMemcpyAsync(cloud, dc.p, n * sizeof(float), cudaMemcpyDeviceToHost, p->stream));
    if (normals) CK(cudaMemcpyAsync(normals, dn.p, n * sizeof(float), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return TSDF_OK;
}

/* K0 alone, for tests: filtered depth [h*w] and normals [h*w*3] (NaN = none) on the host */
tsdf_status tsdf_preprocess(tsdf_handle h, const float* depth, int32_t mem, float* depth_filtered, float* normals) {
    if (!h || !depth_filtered) return bad("null argument");
    Impl* p = I(h);
    if (!p->k0) return bad("handle created without tsdf_config.preprocess");
    CK(cudaSetDevice(p->device));
    if (!p->have_K) { g_err = "camera matrix not set"; return TSDF_ERR_NO_INTRINSICS; }
    const float* dptr;
    tsdf_status st = stage_depth(p, depth, mem, &dptr);
    if (st != TSDF_OK) return st;
    enqueue_prep(p, dptr, 0);                     /* K0 + K1 on the preprocessing stream; the main stream joins */
    st = launch_status();
    if (st != TSDF_OK) return st;
    const size_t npx = (size_t)p->g.img_w * p->g.img_h;
    CK(cudaMemcpyAsync(depth_filtered, p->k0b.zf, npx * sizeof(float), cudaMemcpyDeviceToHost, p->stream));
    if (normals) {
        DevTmp dc, dn;
        CK(dc.alloc(npx * 3 * sizeof(float)));
        CK(dn.alloc(npx * 3 * sizeof(float)));
        launch_cloud(p->g, p->pix, dc.as<float>(), dn.as<float>(), p->stream);     /* the records K1 built from K0's output */
        p->launches++;
        CK(cudaMemcpyAsync(normals, dn.p, npx * 3 * sizeof(float), cudaMemcpyDeviceToHost, p->stream));
        CK(cudaStreamSynchronize(p->stream));
    }
    CK(cudaStreamSynchronize(p->stream));
    return TSDF_OK;
}

tsdf_status tsdf_interpolate_distance(tsdf_handle h, int64_t n, const double* pts, float* out, uint8_t* ok) {
    if (!h || !pts || !out || !ok || n < 0) return bad("bad argument");
    if (n == 0) return TSDF_OK;
    Impl* p = I(h);
    CK(cudaSetDevice(p->device));
    DevTmp dp, dout, dok;
    CK(dp.alloc((size_t)n * 3 * sizeof(double)));
    CK(dout.alloc((size_t)n * sizeof(float)));
    CK(dok.alloc((size_t)n));
    CK(cudaMemcpyAsync(dp.p, pts, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    launch_sample(p->g, p->grid, n, dp.as<double>(), dout.as<float>(), dok.as<uint8_t>(), p->stream);
    p->launches++;
    CK(cudaMemcpyAsync(out, dout.p, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaMemcpyAsync(ok, dok.p, (size_t)n, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return TSDF_OK;
}

int64_t tsdf_number_of_voxels(tsdf_handle h) {
    if (!h) return 0;
    const int64_t m = I(h)->cfg.m;
    return m * m * m;                                                     /* sdf.h:107 (64-bit) */
}
tsdf_status tsdf_stored_range(tsdf_handle h, int32_t* k_begin, int32_t* k_end, int32_t* k_own_begin, int32_t* k_own_end) {
    if (!h) return bad("null handle");
    const GridParams& g = I(h)->g;
    if (k_begin) *k_begin = g.ks0;
    if (k_end) *k_end = g.ks1;
    if (k_own_begin) *k_own_begin = g.ko0;
    if (k_own_end) *k_own_end = g.ko1;
    return TSDF_OK;
}

static tsdf_status xfer_grid(Impl* p, float* D, float* W, int layout, bool down) {
    if (!D || !W) return bad("null argument");
    if (layout != TSDF_LAYOUT_REFERENCE && layout != TSDF_LAYOUT_XFASTEST) return bad("bad layout");
    CK(cudaSetDevice(p->device));
    /* stage through device buffers in slabs of z so the extra footprint stays small */
    const int m = p->g.m, nk = p->g.ks1 - p->g.ks0;
    if (layout == TSDF_LAYOUT_XFASTEST) {
        const int kstep = 16;
        DevTmp tD, tW;
        const size_t plane = (size_t)m * m;
        CK(tD.alloc(plane * kstep * sizeof(float)));
        CK(tW.alloc(plane * kstep * sizeof(float)));
        float *dD = tD.as<float>(), *dW = tW.as<float>();
        for (int k0 = 0; k0 < nk; k0 += kstep) {
            const int kn = nk - k0 < kstep ? nk - k0 : kstep;
            GridParams gs = p->g;
            gs.ks0 = 0; gs.ks1 = kn;
            float2* sub = p->grid + plane * k0;
            if (down) {
                launch_export(gs, sub, dD, dW, 1, p->stream);
                CK(cudaMemcpyAsync(D + plane * k0, dD, plane * kn * sizeof(float), cudaMemcpyDeviceToHost, p->stream));
                CK(cudaMemcpyAsync(W + plane * k0, dW, plane * kn * sizeof(float), cudaMemcpyDeviceToHost, p->stream));
            } else {
                CK(cudaMemcpyAsync(dD, D + plane * k0, plane * kn * sizeof(float), cudaMemcpyHostToDevice, p->stream));
                CK(cudaMemcpyAsync(dW, W + plane * k0, plane * kn * sizeof(float), cudaMemcpyHostToDevice, p->stream));
                launch_import(gs, sub, dD, dW, 1, p->stream);
            }
            p->launches++;
            CK(cudaStreamSynchronize(p->stream));
        }
    } else {
        DevTmp tD, tW;
        CK(tD.alloc((size_t)p->n_stored * sizeof(float)));
        CK(tW.alloc((size_t)p->n_stored * sizeof(float)));
        float *dD = tD.as<float>(), *dW = tW.as<float>();
        if (down) {
            launch_export(p->g, p->grid, dD, dW, 0, p->stream);
            CK(cudaMemcpyAsync(D, dD, (size_t)p->n_stored * sizeof(float), cudaMemcpyDeviceToHost, p->stream));
            CK(cudaMemcpyAsync(W, dW, (size_t)p->n_stored * sizeof(float), cudaMemcpyDeviceToHost, p->stream));
        } else {
            CK(cudaMemcpyAsync(dD, D, (size_t)p->n_stored * sizeof(float), cudaMemcpyHostToDevice, p->stream));
            CK(cudaMemcpyAsync(dW, W, (size_t)p->n_stored * sizeof(float), cudaMemcpyHostToDevice, p->stream));
            launch_import(p->g, p->grid, dD, dW, 0, p->stream);
        }
        p->launches++;
        CK(cudaStreamSynchronize(p->stream));
    }
    CK(cudaGetLastError());
    return TSDF_OK;
}
tsdf_status tsdf_download(tsdf_handle h, float* D, float* W, int32_t layout) {
    if (!h) return bad("null handle");
    return xfer_grid(I(h), D, W, layout, true);
}
tsdf_status tsdf_upload(tsdf_handle h, const float* D, const float* W, int32_t layout) {
    if (!h) return bad("null handle");
    return xfer_grid(I(h), const_cast<float*>(D), const_cast<float*>(W), layout, false);
}
tsdf_status tsdf_device_grid(tsdf_handle h, void** dw, int64_t* n) {
    if (!h) return bad("null handle");
    if (dw) *dw = I(h)->grid;
    if (n) *n = I(h)->n_stored;
    return TSDF_OK;
}

/* ---- pure index / coordinate maps (host), sdf.h:113-157 ---- */
int64_t tsdf_get_array_index(tsdf_handle h, int32_t i, int32_t j, int32_t k) {
    if (!h) return -1;
    const int64_t m = I(h)->cfg.m;
    if (i < 0 || j < 0 || k < 0) return -1;
    if (i >= m || j >= m || k >= m) return -1;
    return m * m * i + m * j + k;
}
void tsdf_get_voxel_coordinates_idx(tsdf_handle h, int64_t idx, int32_t ijk[3]) {
    const int64_t m = I(h)->cfg.m, m2 = m * m;
    ijk[1] = (int32_t)((idx % m2) / m);
    ijk[0] = (int32_t)(idx / m2);
    ijk[2] = (int32_t)(idx % m);
}
void tsdf_get_voxel_coordinates(tsdf_handle h, const double gl[3], double v[3]) {
    world_to_voxel(I(h)->g, gl[0], gl[1], gl[2], v[0], v[1], v[2]);
}
void tsdf_get_global_coordinates(tsdf_handle h, const int32_t ijk[3], double gl[3]) {
    const GridParams& g = I(h)->g;
    gl[0] = voxel_centre(g.vs_x, ijk[0], g.origin[0]);
    gl[1] = voxel_centre(g.vs_y, ijk[1], g.origin[1]);
    gl[2] = voxel_centre(g.vs_z, ijk[2], g.origin[2]);
}

tsdf_status tsdf_exp_map(tsdf_handle h, const double twist[6], double R[9], double t[3]) {
    if (!h || !twist) return bad("null argument");
    Impl* p = I(h);
    CK(cudaSetDevice(p->device));
    double out[12];
    CK(cudaMemcpyAsync(p->scratch_d, twist, 6 * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    launch_exp_map(p->scratch_d, p->scratch_d + 8, p->stream);
    p->launches++;
    CK(cudaMemcpyAsync(out, p->scratch_d + 8, 12 * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    if (R) memcpy(R, out, 9 * sizeof(double));
    if (t) memcpy(t, out + 9, 3 * sizeof(double));
    return TSDF_OK;
}

tsdf_status tsdf_dev_alloc(tsdf_handle h, int64_t bytes, void** dev_ptr) {
    if (!h || !dev_ptr || bytes <= 0) return bad("bad argument");
    CK(cudaSetDevice(I(h)->device));
    CK(cudaMalloc(dev_ptr, (size_t)bytes));
    return TSDF_OK;
}
tsdf_status tsdf_dev_free(tsdf_handle h, void* dev_ptr) {
    if (!h) return bad("null handle");
    CK(cudaSetDevice(I(h)->device));
    CK(cudaFree(dev_ptr));
    return TSDF_OK;
}
tsdf_status tsdf_dev_upload(tsdf_handle h, void* dev_dst, const void* host_src, int64_t bytes) {
    if (!h || !dev_dst || !host_src || bytes < 0) return bad("bad argument");
    Impl* p = I(h);
    CK(cudaSetDevice(p->device));
    CK(cudaMemcpyAsync(dev_dst, host_src, (size_t)bytes, cudaMemcpyHostToDevice, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return TSDF_OK;
}
tsdf_status tsdf_host_alloc_pinned(int64_t bytes, void** host_ptr) {
    if (!host_ptr || bytes <= 0) return bad("bad argument");
    CK(cudaMallocHost(host_ptr, (size_t)bytes));
    return TSDF_OK;
}
tsdf_status tsdf_host_free_pinned(void* host_ptr) {
    CK(cudaFreeHost(host_ptr));
    return TSDF_OK;
}

tsdf_status tsdf_last_stage_ms(tsdf_handle h, float out[3]) {
    if (!h || !out) return bad("null argument");
    Impl* p = I(h);
    if (!p->stage_valid || !p->ev_recorded) return bad("no frame enqueued outside a tsdf_stage_timing_begin/end window yet");
    CK(cudaSetDevice(p->device));
    CK(cudaEventSynchronize(p->ev[3]));
    for (int q = 0; q < 3; q++) CK(cudaEventElapsedTime(&out[q], p->ev[q], p->ev[q + 1]));
    return TSDF_OK;
}
tsdf_status tsdf_event_timer_begin(tsdf_handle h) {
    if (!h) return bad("null handle");
    Impl* p = I(h);
    CK(cudaSetDevice(p->device));
    CK(cudaEventRecord(p->tmr[0], p->stream));
    return TSDF_OK;
}
tsdf_status tsdf_event_timer_end(tsdf_handle h, float* ms) {
    if (!h || !ms) return bad("null argument");
    Impl* p = I(h);
    CK(cudaSetDevice(p->device));
    CK(cudaEventRecord(p->tmr[1], p->stream));
    CK(cudaEventSynchronize(p->tmr[1]));
    CK(cudaEventElapsedTime(ms, p->tmr[0], p->tmr[1]));
    return TSDF_OK;
}
int64_t tsdf_kernel_launch_count(tsdf_handle h) { return h ? I(h)->launches : 0; }

tsdf_status tsdf_stage_timing_begin(tsdf_handle h, int32_t n_frames) {
    if (!h || n_frames < 1 || n_frames > (1 << 20)) return bad("bad argument");
    Impl* p = I(h);
    CK(cudaSetDevice(p->device));
    for (cudaEvent_t e : p->ring_ev) cudaEventDestroy(e);
    p->ring_ev.assign((size_t)4 * n_frames, nullptr);
    for (auto& e : p->ring_ev) CK(cudaEventCreate(&e));
    p->ring_frames = n_frames; p->ring_pos = 0;
    return TSDF_OK;
}
tsdf_status tsdf_stage_timing_end(tsdf_handle h, int32_t* n_frames, float* ms /* n x 3 */) {
    if (!h || !n_frames || !ms) return bad("null argument");
    Impl* p = I(h);
    CK(cudaSetDevice(p->device));
    CK(cudaStreamSynchronize(p->stream));
    const int n = p->ring_pos;
    for (int f = 0; f < n; f++)
        for (int q = 0; q < 3; q++) CK(cudaEventElapsedTime(&ms[3 * f + q], p->ring_ev[(size_t)4 * f + q], p->ring_ev[(size_t)4 * f + q + 1]));
    *n_frames = n;
    for (cudaEvent_t e : p->ring_ev) cudaEventDestroy(e);
    p->ring_ev.clear(); p->ring_frames = 0; p->ring_pos = 0;
    return TSDF_OK;
}
/* debugging aid: globaltimer stamps (ns) of one linearise+update launch at the current pose:
 * out[0] first block start, [1] last block's pixel loop end, [2] final-block start, [3] sums
 * ready, [4] pose updated.  The pose IS updated (one GN step). */
tsdf_status tsdf_debug_phase_times(tsdf_handle h, const float* depth, int32_t mem, int64_t out[5]) {
    if (!h || !out) return bad("null argument");
    Impl* p = I(h);
    CK(cudaSetDevice(p->device));
    if (!p->have_K) { g_err = "camera matrix not set"; return TSDF_ERR_NO_INTRINSICS; }
    const float* dptr;
    tsdf_status st = stage_depth(p, depth, mem, &dptr);
    if (st != TSDF_OK) return st;
    unsigned long long* d = nullptr;
    CK(cudaMalloc(&d, 8 * sizeof(unsigned long long)));
    unsigned long long init[8] = {~0ull, 0, 0, 0, 0, 0, 0, 0};
    CK(cudaMemcpyAsync(d, init, sizeof init, cudaMemcpyHostToDevice, p->stream));
    enqueue_prep(p, dptr, 1);
    enqueue_linearize(p, 1, false);      /* warm */
    enqueue_prep(p, dptr, 1);
    p->dbg_times = d;
    enqueue_linearize(p, 1, false);
    p->dbg_times = nullptr;
    unsigned long long res[8];
    CK(cudaMemcpyAsync(res, d, sizeof res, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    for (int q = 0; q < 5; q++) out[q] = (int64_t)res[q];
    cudaFree(d);
    return TSDF_OK;
}

/* debugging aid: exhaustively compare the kernels' short reciprocal with IEEE 1.0f/x for every
 * float in [x_lo, x_hi] (positive normals); *n_bad = number of mismatching operands */
tsdf_status tsdf_debug_check_rcp(tsdf_handle h, float x_lo, float x_hi, int64_t* n_bad) {
    if (!h || !n_bad || !(x_lo > 0.0f) || !(x_hi >= x_lo)) return bad("bad argument");
    Impl* p = I(h);
    CK(cudaSetDevice(p->device));
    unsigned int lo, hi;
    memcpy(&lo, &x_lo, 4); memcpy(&hi, &x_hi, 4);
    unsigned long long* d = reinterpret_cast<unsigned long long*>(p->scratch_d);
    CK(cudaMemsetAsync(d, 0, sizeof(unsigned long long), p->stream));
    launch_check_rcp(lo, hi + 1u, d, p->stream);
    p->launches++;
    unsigned long long r = 0;
    CK(cudaMemcpyAsync(&r, d, sizeof r, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    *n_bad = (int64_t)r;
    return TSDF_OK;
}

/* debugging aid: the fusion weight w = (float)exp(-0.5 e^2) (sdf.cpp:278) for EVERY float e in [e_lo, e_hi]:
 * *n_ambiguous = operands whose weight is not provably the correctly rounded float (see k_check_wexp); the first
 * `cap` of them come back as (e, w_device) pairs for a comparison with the host libm */
tsdf_status tsdf_debug_check_weight_exp(tsdf_handle h, float e_lo, float e_hi, int64_t* n_ambiguous, float* e_list, float* w_list, int32_t cap) {
    if (!h || !n_ambiguous || !e_list || !w_list || cap < 1 || !(e_lo >= 0.0f) || !(e_hi >= e_lo)) return bad("bad argument");
    Impl* p = I(h);
    CK(cudaSetDevice(p->device));
    unsigned int lo, hi;
    memcpy(&lo, &e_lo, 4); memcpy(&hi, &e_hi, 4);
    DevTmp de, dw;
    CK(de.alloc((size_t)cap * sizeof(float)));
    CK(dw.alloc((size_t)cap * sizeof(float)));
    unsigned long long* d = reinterpret_cast<unsigned long long*>(p->scratch_d);
    CK(cudaMemsetAsync(d, 0, sizeof(unsigned long long), p->stream));
    launch_check_wexp(lo, hi + 1u, d, de.as<float>(), dw.as<float>(), cap, p->stream);
    p->launches++;
    unsigned long long r = 0;
    CK(cudaMemcpyAsync(&r, d, sizeof r, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    const size_t n = r < (unsigned long long)cap ? (size_t)r : (size_t)cap;
    if (n) {
        CK(cudaMemcpy(e_list, de.p, n * sizeof(float), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(w_list, dw.p, n * sizeof(float), cudaMemcpyDeviceToHost));
    }
    *n_ambiguous = (int64_t)r;
    return TSDF_OK;
}

/* debugging aid: run fusion's self-check build for `depth` at the current pose (no voxel is
 * written): every voxel of every unit certified by the pyramid is compared with the exact fp64 path.
 * out[0] = voxels in certified units, out[1] = of those, wrong (must be 0), out[2] = work items */
tsdf_status tsdf_debug_fuse_check(tsdf_handle h, const float* depth, int32_t mem, int64_t out[3]) {
    if (!h || !out) return bad("null argument");
    Impl* p = I(h);
    CK(cudaSetDevice(p->device));
    if (!p->have_K) { g_err = "camera matrix not set"; return TSDF_ERR_NO_INTRINSICS; }
    const float* dptr;
    tsdf_status st = stage_depth(p, depth, mem, &dptr);
    if (st != TSDF_OK) return st;
    CK(cudaMemsetAsync(p->n_upd_dev + 2, 0, 2 * sizeof(unsigned long long), p->stream));
    enqueue_prep(p, dptr, 0);
    p->fuse_check = 1;
    enqueue_fuse(p);
    p->fuse_check = 0;
    unsigned long long v[4];
    unsigned int items = 0;
    CK(cudaMemcpyAsync(v, p->n_upd_dev, sizeof v, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaMemcpyAsync(&items, p->fuse_item_count, sizeof items, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    out[0] = (int64_t)v[2]; out[1] = (int64_t)v[3]; out[2] = (int64_t)items;
    return TSDF_OK;
}

/* debugging aid / roofline ceiling: time `reps` plain read-modify-write streams over the whole store
 * (the free-space update on every voxel, no geometry).  MODIFIES the grid.  ms = mean per pass. */
tsdf_status tsdf_debug_stream_rmw(tsdf_handle h, int32_t reps, float* ms) {
    if (!h || !ms || reps < 1) return bad("bad argument");
    Impl* p = I(h);
    CK(cudaSetDevice(p->device));
    launch_stream_rmw(p->grid, p->n_stored, -p->g.delta, p->stream);
    CK(cudaEventRecord(p->tmr[0], p->stream));
    for (int r = 0; r < reps; r++) launch_stream_rmw(p->grid, p->n_stored, -p->g.delta, p->stream);
    CK(cudaEventRecord(p->tmr[1], p->stream));
    CK(cudaEventSynchronize(p->tmr[1]));
    p->launches += reps + 1;
    float t = 0;
    CK(cudaEventElapsedTime(&t, p->tmr[0], p->tmr[1]));
    *ms = t / reps;
    return TSDF_OK;
}

#ifdef TSDF_SWZ_EXPERIMENT
/* layout experiment (variant builds only): time `reps` linearisations (no pose update) at the current pose on the
 * store as it is and on a swizzled copy; out_ms[0], out_ms[1] = mean ms per launch; sums_equal = the 30 sums match bitwise */
tsdf_status tsdf_debug_swizzle_probe(tsdf_handle h, const float* depth, int32_t mem, int32_t reps, float out_ms[2], int32_t* sums_equal) {
    Impl* p = I(h);
    CK(cudaSetDevice(p->device));
    const float* dptr;
    tsdf_status st = stage_depth(p, depth, mem, &dptr);
    if (st != TSDF_OK) return st;
    float2* swz = nullptr;
    CK(cudaMalloc(&swz, (size_t)p->n_stored * sizeof(float2)));
    launch_swizzle(p->g, p->grid, swz, p->stream);
    enqueue_prep(p, dptr, 1);
    double sums[2][30];
    for (int v = 0; v < 2; v++) {
        LinearizeArgs a = lin_args(p, 0, false);
        if (v == 1) a.grid = swz;
        auto go = [&]() { if (v == 0) launch_linearize(a, p->lin_blocks, 0, 0ull, p->stream); else launch_linearize_swz(a, p->lin_blocks, p->stream); };
        go(); go();
        CK(cudaEventRecord(p->tmr[0], p->stream));
        for (int r = 0; r < reps; r++) go();
        CK(cudaEventRecord(p->tmr[1], p->stream));
        CK(cudaEventSynchronize(p->tmr[1]));
        float t = 0; CK(cudaEventElapsedTime(&t, p->tmr[0], p->tmr[1]));
        out_ms[v] = t / reps;
        st = fetch_pose(p); if (st != TSDF_OK) return st;
        memcpy(sums[v], p->pose_pin->sums, sizeof sums[v]);
    }
    *sums_equal = memcmp(sums[0], sums[1], sizeof sums[0]) == 0 ? 1 : 0;
    cudaFree(swz);
    return launch_status();
}
#endif

tsdf_status tsdf_total_updates(tsdf_handle h, int32_t reset, int64_t* total) {
    if (!h) return bad("null handle");
    Impl* p = I(h);
    CK(cudaSetDevice(p->device));
    unsigned long long v[2] = {0, 0};
    CK(cudaMemcpyAsync(v, p->n_upd_dev, sizeof v, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    if (total) *total = (int64_t)v[1];
    if (reset) { CK(cudaMemsetAsync(p->n_upd_dev + 1, 0, sizeof(unsigned long long), p->stream)); }
    return TSDF_OK;
}

tsdf_status tsdf_flush_l2(tsdf_handle h) {
    if (!h) return bad("null handle");
    Impl* p = I(h);
    CK(cudaSetDevice(p->device));
    if (!p->flush_buf) CK(cudaMalloc(&p->flush_buf, FLUSH_BYTES));
    launch_flush(p->flush_buf, FLUSH_BYTES / sizeof(float), p->stream);
    p->launches++;
    CK(cudaGetLastError());
    return TSDF_OK;
}

/* ---- sharding ------------------------------------------------------------------------------ */
tsdf_status tsdf_slab_plan(const tsdf_config* cfg, int32_t out[5]) {
    if (!cfg || !out) return bad("null argument");
    if (cfg->m < 8 || (cfg->m % 4) != 0) return bad("m must be a multiple of 4, >= 8");
    if (cfg->n_shards < 1 || cfg->n_shards > MAX_WORLD || cfg->shard_rank < 0 || cfg->shard_rank >= cfg->n_shards) return bad("bad shard configuration");
    int ko0, ko1, ks0, ks1, halo;
    slab_range(*cfg, ko0, ko1, ks0, ks1, halo);
    out[0] = ko0; out[1] = ko1; out[2] = ks0; out[3] = ks1; out[4] = halo;
    return TSDF_OK;
}
tsdf_status tsdf_shard_ipc_export(tsdf_handle h, uint8_t out[TSDF_IPC_HANDLE_BYTES]) {
    if (!h || !out) return bad("null argument");
    Impl* p = I(h);
    CK(cudaSetDevice(p->device));
    static_assert(sizeof(cudaIpcMemHandle_t) <= TSDF_IPC_HANDLE_BYTES, "ipc handle size");
    cudaIpcMemHandle_t mh;
    CK(cudaIpcGetMemHandle(&mh, p->mailbox));
    memset(out, 0, TSDF_IPC_HANDLE_BYTES);
    memcpy(out, &mh, sizeof mh);
    return TSDF_OK;
}
tsdf_status tsdf_shard_ipc_attach(tsdf_handle h, int32_t world, const uint8_t* handles) {
    if (!h || !handles) return bad("null argument");
    Impl* p = I(h);
    if (world != p->cfg.n_shards) return bad("world must equal tsdf_config.n_shards");
    CK(cudaSetDevice(p->device));
    p->links.world = world; p->links.rank = p->cfg.shard_rank;
    for (int r = 0; r < world; r++) {
        if (r == p->cfg.shard_rank) { p->links.box[r] = p->mailbox; continue; }
        cudaIpcMemHandle_t mh;
        memcpy(&mh, handles + (size_t)r * TSDF_IPC_HANDLE_BYTES, sizeof mh);
        void* ptr = nullptr;
        CK(cudaIpcOpenMemHandle(&ptr, mh, cudaIpcMemLazyEnablePeerAccess));
        p->ipc_opened.push_back(ptr);
        p->links.box[r] = reinterpret_cast<Mailbox*>(ptr);
    }
    p->exchange_mode = world > 1 ? 1 : 0;
    p->seqno = 0;
    return TSDF_OK;
}
tsdf_status tsdf_shard_attach_local(tsdf_handle* handles, int32_t world) {
    if (!handles || world < 1 || world > MAX_WORLD) return bad("bad argument");
    bool same_dev = true;
    for (int r = 0; r < world; r++) {
        if (!handles[r]) return bad("null handle in group");
        Impl* p = I(handles[r]);
        if (p->cfg.n_shards != world || p->cfg.shard_rank != r) return bad("handles must be ordered by shard_rank with n_shards = world");
        if (p->device != I(handles[0])->device) same_dev = false;
    }
    for (int r = 0; r < world; r++) {
        Impl* p = I(handles[r]);
        CK(cudaSetDevice(p->device));
        p->links.world = world; p->links.rank = r;
        for (int q = 0; q < world; q++) {
            Impl* o = I(handles[q]);
            if (!same_dev && o->device != p->device) {
                cudaError_t e = cudaDeviceEnablePeerAccess(o->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { g_err = "peer access unavailable"; return TSDF_ERR_CUDA; }
                cudaGetLastError();
            }
            p->links.box[q] = o->mailbox;
        }
        p->exchange_mode = world > 1 ? (same_dev ? 2 : 1) : 0;
        p->seqno = 0;
        if (same_dev && r > 0 && p->stream != I(handles[0])->stream) {
            /* one stream for the whole group: program order is the cross-shard dependency.  The stream is
             * reference-counted, so the handles may be destroyed in any order (and attaching twice is a no-op) */
            CK(cudaStreamSynchronize(p->stream));
            if (--*p->stream_refs == 0) { cudaStreamDestroy(p->stream); delete p->stream_refs; }
            p->stream = I(handles[0])->stream;
            p->stream_refs = I(handles[0])->stream_refs;
            ++*p->stream_refs;
        }
    }
    return TSDF_OK;
}


static tsdf_status group_check(tsdf_handle* handles, int32_t world) {
    if (!handles || world < 1 || world > MAX_WORLD) return bad("bad group");
    for (int r = 0; r < world; r++) {
        if (!handles[r]) return bad("null handle in group");
        if (I(handles[r])->links.world != world || I(handles[r])->links.rank != r) return bad("group not attached (tsdf_shard_attach_local)");
    }
    return TSDF_OK;
}
tsdf_status tsdf_group_set_intrinsics(tsdf_handle* handles, int32_t world, const double K[9]) {
    tsdf_status st = group_check(handles, world);
    for (int r = 0; r < world && st == TSDF_OK; r++) st = tsdf_set_intrinsics(handles[r], K);
    return st;
}
tsdf_status tsdf_group_set_pose(tsdf_handle* handles, int32_t world, const double R[9], const double t[3]) {
    tsdf_status st = group_check(handles, world);
    for (int r = 0; r < world && st == TSDF_OK; r++) st = tsdf_set_pose(handles[r], R, t);
    return st;
}
static tsdf_status group_stage(tsdf_handle* handles, int32_t world, const float* depth, int32_t mem, std::vector<const float*>& dptr) {
    dptr.resize(world);
    for (int r = 0; r < world; r++) {
        Impl* p = I(handles[r]);
        if (!p->have_K) { g_err = "camera matrix not set"; return TSDF_ERR_NO_INTRINSICS; }
        if (mem == TSDF_DEVICE && p->device != I(handles[0])->device) return bad("device depth needs all shards on one device");
        CK(cudaSetDevice(p->device));
        tsdf_status st = stage_depth(p, depth, mem, &dptr[r]);
        if (st != TSDF_OK) return st;
    }
    return TSDF_OK;
}
static void group_gn_iteration(tsdf_handle* handles, int32_t world, int do_update) {
    for (int r = 0; r < world; r++) { cudaSetDevice(I(handles[r])->device); enqueue_linearize(I(handles[r]), do_update, false); }
    if (I(handles[0])->exchange_mode == 2)
        for (int r = 0; r < world; r++) enqueue_combine(I(handles[r]), do_update);
}
tsdf_status tsdf_group_linearize(tsdf_handle* handles, int32_t world, const float* depth, int32_t mem,
                                 double A[36], double b[6], tsdf_track_stats* stats) {
    tsdf_status st = group_check(handles, world);
    if (st != TSDF_OK) return st;
    std::vector<const float*> dptr;
    st = group_stage(handles, world, depth, mem, dptr);
    if (st != TSDF_OK) return st;
    for (int r = 0; r < world; r++) { cudaSetDevice(I(handles[r])->device); enqueue_prep(I(handles[r]), dptr[r], 1); }
    group_gn_iteration(handles, world, 0);
    st = launch_status();
    if (st != TSDF_OK) return st;
    int miss = 0;
    for (int r = world - 1; r >= 0; r--) {
        Impl* p = I(handles[r]);
        CK(cudaSetDevice(p->device));
        st = fetch_pose(p);
        if (st != TSDF_OK) return st;
        miss |= p->pose_pin->halo_miss;
    }
    tsdf_track_stats s;
    fill_stats(*I(handles[0])->pose_pin, &s);
    s.halo_miss = miss;
    if (A) memcpy(A, s.A, sizeof s.A);
    if (b) memcpy(b, s.b, sizeof s.b);
    if (stats) *stats = s;
    if (miss) { g_err = "halo too small for the tracking stencil"; return TSDF_ERR_HALO; }
    return TSDF_OK;
}
tsdf_status tsdf_group_frame(tsdf_handle* handles, int32_t world, const float* depth, int32_t mem,
                             int32_t track, int32_t fuse, double R_out[9], double t_out[3],
                             tsdf_track_stats* stats, int64_t* n_updated) {
    tsdf_status st = group_check(handles, world);
    if (st != TSDF_OK) return st;
    std::vector<const float*> dptr;
    st = group_stage(handles, world, depth, mem, dptr);
    if (st != TSDF_OK) return st;
    for (int r = 0; r < world; r++) { cudaSetDevice(I(handles[r])->device); enqueue_prep(I(handles[r]), dptr[r], track ? 1 : 0); }
    if (track)
        for (int it = 0; it < I(handles[0])->g.max_iter; it++) group_gn_iteration(handles, world, 1);
    if (fuse)
        for (int r = 0; r < world; r++) {
            Impl* p = I(handles[r]);
            cudaSetDevice(p->device);
            enqueue_fuse(p);
            note_cuda(cudaMemcpyAsync(p->n_upd_pin, p->n_upd_dev, sizeof(unsigned long long), cudaMemcpyDeviceToHost, p->stream));
        }
    st = launch_status();
    if (st != TSDF_OK) return st;
    int64_t nupd = 0;
    int miss = 0;
    for (int r = world - 1; r >= 0; r--) {
        Impl* p = I(handles[r]);
        CK(cudaSetDevice(p->device));
        st = fetch_pose(p);
        if (st != TSDF_OK) return st;
        if (fuse) nupd += (int64_t)*p->n_upd_pin;
        miss |= p->pose_pin->halo_miss;
    }
    if (n_updated) *n_updated = nupd;
    I(handles[0])->pose_pin->halo_miss = miss;
    if (!track) {
        const PoseState& ps = *I(handles[0])->pose_pin;
        if (R_out) memcpy(R_out, ps.R, sizeof ps.R);
        if (t_out) memcpy(t_out, ps.t, sizeof ps.t);
        if (stats) memset(stats, 0, sizeof *stats);
        return TSDF_OK;
    }
    return track_result(I(handles[0]), R_out, t_out, stats);
}

}  // extern "C"
