/* tsdf_internal.h — launcher declarations shared by tsdf_kernels.cu and tsdf_abi.cu */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "tsdf_core.cuh"
#include "mc_params.h"

namespace tsdf {

/* K2 launch shape.  Default (LIN_MICRO): ONE 768-thread block per SM (24 warps at 72 registers fill the register file
 * about as far as three 256-thread blocks did), each with a contiguous run of 4 x 4 micro-tiles of the strided pixel
 * grid.  148 blocks means 148 partial sums: ONE ticket level instead of two (three dependent global round trips
 * instead of six between the slowest block's last pixel and the solve) — 22.6 -> 21.4 us per iteration.
 * -DLIN_TILES selects the previous shape: one 256-thread block per LIN_TW x LIN_TH pixel tile, groups of LIN_GROUP. */
#ifndef LIN_TILES
#define LIN_MICRO 1
#endif
#ifndef LIN_THREADS_DEF
#ifdef LIN_MICRO
#define LIN_THREADS_DEF 768
#else
#define LIN_THREADS_DEF 256
#endif
#endif
constexpr int LIN_THREADS = LIN_THREADS_DEF;   /* a multiple of 256: two pixels per warp, 16 pixels per micro-tile */
static_assert(LIN_THREADS % 256 == 0 && LIN_THREADS <= 1024, "k_linearize: whole micro-tiles per sweep");
#ifndef LIN_TW_DEF
#define LIN_TW_DEF 8
#endif
#ifndef LIN_TH_DEF
#define LIN_TH_DEF 10
#endif
constexpr int LIN_TW = LIN_TW_DEF, LIN_TH = LIN_TH_DEF;      /* LIN_TILES: strided-pixel tile per block (columns x rows) = 80 pixels = 5 sweeps of 16 */
constexpr int LIN_PARTIAL_STRIDE = 32;    /* doubles per block partial (30 used) */
constexpr int MAX_WORLD = 16;
#ifndef LIN_GROUP_DEF
#ifdef LIN_MICRO
#define LIN_GROUP_DEF 256               /* >= the block count of any current GPU: a single reduction level */
#else
#define LIN_GROUP_DEF 32
#endif
#endif
constexpr int LIN_GROUP = LIN_GROUP_DEF;  /* blocks per first-level reduction group */
#ifndef FUSE_THREADS_DEF
#define FUSE_THREADS_DEF 128
#endif
constexpr int FUSE_THREADS = FUSE_THREADS_DEF;
#ifndef CERT_THREADS_DEF
#define CERT_THREADS_DEF 512
#endif
constexpr int CERT_THREADS = CERT_THREADS_DEF;   /* k_fuse_cert's block size: two 512-thread blocks per SM (the same 32 warps as eight
                                                  * 128-thread blocks; measured: trajectory fusion 125.0 -> 121.8 us, 256: 122.2, 1024: 122.7) */
#ifndef FUSE_MIN_BLOCKS
#define FUSE_MIN_BLOCKS 5               /* resident blocks per SM the fusion kernel is compiled for */
#endif
#ifndef FUSE_COLOR_MIN_BLOCKS
#define FUSE_COLOR_MIN_BLOCKS 5         /* colour variants of the exact pass (more live state per voxel): 102 registers, 20 warps/SM
                                         * (dense colour: 1.42 ms at 4 blocks, 1.36 ms at 5, measured) */
#endif
#ifndef CERT_MIN_BLOCKS
#define CERT_MIN_BLOCKS 2
#endif
#ifndef LIN_MIN_BLOCKS
#ifdef LIN_MICRO
#define LIN_MIN_BLOCKS 1
#else
#define LIN_MIN_BLOCKS 3
#endif
#endif

/* cross-shard exchange of the reduced normal equations (one slot per rank, double-buffered
 * by sequence parity so a fast rank cannot overwrite a slot a slow rank still reads) */
struct Mailbox {
    double sums[2][MAX_WORLD][32];          /* deferred same-device exchange (mode 2): plain values, ordered by the shared stream */
    unsigned long long seq[2][MAX_WORLD];   /* (unused by mode 1 since the self-validating words below) */
    /* cross-device exchange (mode 1): every 64-bit word carries 32 bits of payload and the 32-bit sequence tag of
     * the iteration it belongs to, so a word is either entirely old or entirely new (naturally aligned 8-byte
     * peer stores are single transactions): no system-scope fence and no separate flag round trip.
     * [parity][source rank][slot][low / high half of the double] */
    unsigned long long w[2][MAX_WORLD][32][2];
};

struct ShardLinks {
    int32_t world, rank;
    Mailbox* box[MAX_WORLD];               /* box[r] = rank r's mailbox (peer-mapped), box[rank] = ours */
};

struct LinearizeArgs {
    GridParams g;
    const float2* grid;
    const PixRec* pix;
    const float4* pts;                     /* strided pixels back-projected by k_prep: x, y, z (NaN = invalid) */
    PoseState* pose;
    double* partials;                      /* nblk * LIN_PARTIAL_STRIDE */
    unsigned int* ticket;
    unsigned int* group_ticket;            /* one per LIN_GROUP blocks */
    double* group_partials;                /* ngroups * LIN_PARTIAL_STRIDE */
    float* dbgJ; float* dbgPsi; uint8_t* dbgFlag;   /* optional per-pixel records */
    unsigned long long* dbg_times;         /* optional: globaltimer stamps of the kernel's phases */
    int32_t do_update;                     /* 1: solve + pose update in the last block */
    int32_t first;                         /* 1: first launch of a frame (resets the per-frame tracking state) */
    int32_t iter;                          /* 0-based index of this launch within the frame's GN loop */
    int32_t px_per_block;
    int32_t force_idx64;                   /* tests: take the 64-bit voxel index path on a small store (TSDF_B200_IDX64=1 at create) */
    ShardLinks links;                      /* world = 1: no exchange */
};

/* result of every launch / event call since the last take: checked by the ABI after enqueueing a frame */
void note_cuda(cudaError_t e);
cudaError_t take_launch_error();

/* K0 (tsdf_k0.cu): the node's pre-processing, sdf_reconstruction.cpp:37-49 */
constexpr int K0_PAD = 2;                  /* bilateral grid padding (PCL padding_xy = padding_z = 2) */
constexpr int K0_SD_MAX = 256;             /* depth bins: ranges beyond (256 - 5) * sigma_r share the last bins */
constexpr int K0_RADIUS = 10;              /* largest normal smoothing size supported (the node uses 10) */
struct K0Params {
    int32_t img_w, img_h;
    float sigma_s, sigma_r;                /* pcl::FastBilateralFilter defaults 15, 0.05 (the node sets none, :38-41) */
    float max_depth_change, smoothing;     /* 0.02, 10 (:46-47) */
    float cx, cy, inv_fx, inv_fy;          /* K1's back-projection constants */
};
struct K0Buffers {
    unsigned int* minmax;                  /* [0] ~bits(zmin), [1] bits(zmax) */
    float2 *grid_a, *grid_b;               /* bilateral grid, (sum z, count) per cell */
    short* bins;                           /* depth bin of every pixel (-1 = none) */
    float* zf;                             /* filtered depth */
    uint8_t* edge;                         /* depth-discontinuity map */
    float4 *DX, *DY;                       /* central-difference 3-D gradients (+ validity) */
    float4* normals;                       /* nx, ny, nz, valid */
};
int launch_k0(const K0Params& P, const K0Buffers& B, const float* depth, cudaStream_t s);   /* returns #launches */

/* nrm_in: per-pixel normals from K0 (NULL: K1 computes its own 4-neighbour normals) */
void launch_prep(const GridParams& g, const float* depth, const float4* nrm_in, PixRec* pix, float2* cert0, float4* pts, const uint8_t* rgb3, uchar4* rgb4, double* cosn, cudaStream_t s);
void launch_pyramid(const CertPyramid& P, float2* cert, unsigned int* ticket, cudaStream_t s);
/* exchange_mode: 0 none, 1 in-kernel mailbox all-reduce over peer memory (one kernel per
 * device, all running concurrently), 2 deferred (same-device shards: publish, then
 * launch_gn_combine sums in rank order).  seqno labels the exchange. */
void launch_linearize(const LinearizeArgs& a, int nblk, int exchange_mode, unsigned long long seqno, cudaStream_t s);
void launch_gn_combine(const LinearizeArgs& a, unsigned long long seqno, cudaStream_t s);
#ifdef TSDF_SWZ_EXPERIMENT
void launch_swizzle(const GridParams& g, const float2* src, float2* dst, cudaStream_t s);
void launch_linearize_swz(const LinearizeArgs& a, int nblk, cudaStream_t s);
#endif
struct FuseArgs {
    GridParams g;
    float2* grid;
    const PixRec* pix;
    const PoseState* pose;
    double* tables;                        /* 9*m + 8 doubles: nine tables, tinv[3], affine flag, affine step[3] */
    unsigned long long* items;             /* capacity rows * (m/128 + 1) */
    float4* item_c;                        /* per item: the row's camera-space centre at i = 0 in fp32 (affine certificates) */
    unsigned int* item_count;
    unsigned long long* n_updated;         /* [0] this launch, [1] running total, [2..3] self-check counters */
    CertPyramid pyr;                       /* certificate pyramid (CERT_LEVELS levels, up to the whole image) */
    const float2* cert;
    unsigned long long* units;             /* queue of uncertified lane units (capacity: stored voxels / 4) */
    unsigned int* unit_count;
    int nblk, nblk_cert, nblk_color;
    int check;                             /* 1: run the self-check build (no stores) */
    float4* color;                         /* {Color_W, R, G, B} per voxel, or NULL: no colour update this frame */
    const uchar4* rgb4;                    /* the frame's colour image, packed by k_prep */
    const double* cosn;                    /* per-pixel |n_z| / ||n|| (sdf.cpp:294), tabulated by k_prep */
};
int launch_fuse(const FuseArgs& f, cudaStream_t s);    /* returns the number of kernels launched */
int fuse_cert_blocks_per_sm();
int fuse_color_blocks_per_sm();
void launch_fill_color(float4* color, int64_t n, cudaStream_t s);
void launch_export_color(const GridParams& g, const float4* color, float* cw, float* r, float* gg, float* b, int layout_ref, cudaStream_t s);
void launch_sample_color(const GridParams& g, const float4* color, int64_t n, const double* gpts, float* rgba, cudaStream_t s);
void launch_fill(float2* grid, int64_t n, float d0, cudaStream_t s);
void launch_sample(const GridParams& g, const float2* grid, int64_t n, const double* pts, float* out, uint8_t* ok, cudaStream_t s);
void launch_export(const GridParams& g, const float2* grid, float* D, float* W, int layout, cudaStream_t s);
void launch_import(const GridParams& g, float2* grid, const float* D, const float* W, int layout, cudaStream_t s);
void launch_cloud(const GridParams& g, const PixRec* pix, float* cloud, float* normals, cudaStream_t s);
void launch_exp_map(const double* twist, double* out12, cudaStream_t s);
void launch_flush(float* buf, int64_t n, cudaStream_t s);
void launch_stream_rmw(float2* grid, int64_t n, float neg_delta, cudaStream_t s);
void launch_check_rcp(unsigned int lo, unsigned int hi, unsigned long long* n_bad, cudaStream_t s);
void launch_check_wexp(unsigned int lo, unsigned int hi, unsigned long long* n_amb, float* e_list, float* w_list, int cap, cudaStream_t s);
/* mesher (tsdf_mesh.cu) */
size_t mesh_scan_bytes(int64_t n_rows);
int mesh_zsplit();
void launch_mesh_count(const GridParams& g, const McParams& P, const float2* grid, unsigned int* row_count, unsigned int* row_off,
                       void* scan_tmp, size_t scan_bytes, unsigned long long* cells, unsigned int* cell_counter, unsigned int cell_cap, cudaStream_t s);
void launch_mesh_emit_list(const GridParams& g, const McParams& P, const float2* grid, const unsigned long long* cells, const unsigned int* cell_counter,
                           unsigned int n_threads, const unsigned int* row_off, float* xyz, unsigned int cell_cap, unsigned int vtx_cap, cudaStream_t s);
void launch_mesh_emit(const GridParams& g, const McParams& P, const float2* grid, const unsigned int* row_off, float* xyz, cudaStream_t s);
void launch_mesh_world(const GridParams& g, const float* xyz, int64_t n, double* world, cudaStream_t s);
int  fuse_blocks_per_sm();
int  linearize_blocks_per_sm();

}  // namespace tsdf
