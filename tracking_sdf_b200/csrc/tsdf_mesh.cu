/*
 * tsdf_mesh.cu — the reference's mesher on the device:
 * pcl::MarchingCubesSDF::performReconstruction (marching_cubes_sdf.cpp:243-287) over the interior
 * cells, triangle soup in the reference's order (cells in (i,j,k) lexicographic order = its idx
 * order, triangles in table order), plus the marker post-processing of SDF::visualize
 * (sdf.cpp:352-385: + sdf_origin in double, one interpolate_color per vertex).
 *
 * The store is x fastest, the output order is k fastest, so a warp takes 32 consecutive i at a fixed
 * j and sweeps k: every load is a coalesced 256-byte row segment, each lane walks one (i,j) row of
 * cells in output order.
 *
 * ONE sweep over the store: it counts the vertices of every (row, k-chunk) segment AND appends one 8-byte record
 * per surface cell (i, j, k, configuration, vertex offset inside its segment — the lane's running count) to a
 * compact list (warp-staged appends, one atomic per >= 33 records).  After the exclusive scan of the segment
 * counts, a second kernel walks the LIST (about 1 % of the cells), re-gathers each cell's eight corners and
 * writes its triangles at row_off[segment] + offset: the soup comes out in the reference's order although the
 * list is unordered.  The store is read once (plus the listed cells' corners); the previous two-sweep emit
 * (k_mc_sweep<true>) remains as the fallback when the list overflows its capacity.
 */
#include <cub/device/device_scan.cuh>

#include "mc_core.cuh"
#include "tsdf_internal.h"

namespace tsdf {

constexpr int MC_WARPS = 8;      /* warps per block = consecutive j rows */
#ifndef MC_ZSPLIT_DEF
#define MC_ZSPLIT_DEF 4
#endif
constexpr int MC_ZSPLIT = MC_ZSPLIT_DEF;     /* chunks of the k sweep (more loads in flight) */

constexpr int MC_TILE = 31;      /* cells per warp along i: 32 lanes load 32 voxels, lanes 0..30 own a cell (i+1 comes from lane+1) */

/* surface-cell record: i (12 bits) | j (12) | k (12) | configuration (8) | vertex offset inside the segment (16) */
__device__ __forceinline__ unsigned long long mc_pack_cell(int i, int j, int k, int ci, unsigned int off) {
    return (unsigned long long)i | ((unsigned long long)j << 12) | ((unsigned long long)k << 24) | ((unsigned long long)ci << 36) |
           ((unsigned long long)off << 44);
}

template <bool EMIT>
__global__ void __launch_bounds__(MC_WARPS * 32) k_mc_sweep(GridParams g, McParams P, const float2* __restrict__ grid, int k_lo_all, int k_hi_all,
                                                            unsigned int* __restrict__ row_count, const unsigned int* __restrict__ row_off,
                                                            float* __restrict__ xyz,
                                                            unsigned long long* __restrict__ cells, unsigned int* cell_counter, unsigned int cell_cap) {
    __shared__ unsigned long long s_stage[EMIT ? 1 : MC_WARPS][EMIT ? 1 : 96];
    int staged = 0;                                                   /* warp-uniform */
    const int m = g.m;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = 1 + blockIdx.x * MC_TILE + lane;                    /* this lane's voxel column; its cell if lane < 31 */
    const int j = blockIdx.y * MC_WARPS + warp;
    if (j < 1 || j > m - 2) return;                                   /* warp-uniform */
    unsigned long long* stage = s_stage[EMIT ? 0 : warp];
    auto flush = [&]() {                                              /* append the warp's staged records: one atomic */
        if (staged == 0) return;
        unsigned int basec = 0;
        if (lane == 0) basec = atomicAdd(cell_counter, (unsigned int)staged);
        basec = __shfl_sync(0xffffffffu, basec, 0);
        for (int q = lane; q < staged; q += 32)
            if (basec + (unsigned int)q < cell_cap) cells[basec + q] = stage[q];     /* past the capacity: counted, not stored */
        __syncwarp();
        staged = 0;
    };
    /* the k range is cut into gridDim.z chunks; a row's chunks are consecutive in the output */
    const int nz = (int)gridDim.z, zc = (int)blockIdx.z;
    const int per = (k_hi_all - k_lo_all + nz) / nz;
    const int k_lo = k_lo_all + zc * per;
    const int k_hi = (k_lo + per - 1 < k_hi_all) ? (k_lo + per - 1) : k_hi_all;
    if (k_lo > k_hi) return;
    const bool cell_ok = (lane < MC_TILE) && (i <= m - 2);
    const float fm = (float)m;
    const size_t plane = (size_t)m * m;
    const float iso = P.iso;
    /* lanes past the row end load the last voxel again (never used: their cells are not cell_ok) */
    const float2* p = grid + (size_t)(k_lo - g.ks0) * plane + (size_t)j * m + (i < m ? i : m - 1);   /* (i, j, k_lo); row j+1 is + m */
    /* Per layer each lane keeps two voxels, a = (i,j) and b = (i,j+1), and a 4-bit summary of them
     * (below-iso and W>0 flags); one shuffle brings the neighbour column's summary, so a cell is
     * classified from four small words.  The eight distances are only gathered (four more shuffles)
     * when some cell of the warp actually holds surface.  Loads run two layers ahead. */
    auto summary = [&](const float2& a, const float2& b) {
        return (unsigned)(a.x < iso) | ((unsigned)(b.x < iso) << 1) | ((unsigned)(a.y > 0.0f) << 2) | ((unsigned)(b.y > 0.0f) << 3);
    };
    float2 a0 = __ldg(p), b0 = __ldg(p + m);                          /* layer k   */
    p += plane;
    float2 a1 = __ldg(p), b1 = __ldg(p + m);                          /* layer k+1 */
    p += plane;
    unsigned s0 = summary(a0, b0);
    unsigned t0 = __shfl_down_sync(0xffffffffu, s0, 1);               /* column i+1 */
    unsigned int n_row = 0;
    const unsigned int base = EMIT ? (cell_ok ? row_off[((size_t)i * m + j) * nz + zc] : 0u) : 0u;
    for (int k = k_lo; k <= k_hi; k++) {
        float2 a2 = a1, b2 = b1;
        if (k + 2 <= k_hi + 1) { a2 = __ldg(p); b2 = __ldg(p + m); p += plane; }     /* warp-uniform */
        const unsigned s1 = summary(a1, b1);
        const unsigned t1 = __shfl_down_sync(0xffffffffu, s1, 1);
        /* corners: 0 (i,j,k) 1 (i+1,j,k) 2 (i+1,j,k+1) 3 (i,j,k+1) 4 (i,j+1,k) 5 (i+1,j+1,k) 6 (i+1,j+1,k+1) 7 (i,j+1,k+1) */
        const unsigned all_w = (s0 & t0 & s1 & t1) >> 2;              /* == 3: all eight corners observed (marching_cubes_sdf.cpp:221) */
        const unsigned lo = (s0 & 3u) | ((t0 & 3u) << 2), hi = (s1 & 3u) | ((t1 & 3u) << 2);   /* below-iso flags of the two layers */
        const bool surf = cell_ok && (all_w == 3u) && ((lo | hi) != 0u) && ((lo & hi) != 15u);
        const unsigned any = __ballot_sync(0xffffffffu, surf);
        if (any) {                                                    /* warp-uniform, rare */
            const float a0n = __shfl_down_sync(0xffffffffu, a0.x, 1), b0n = __shfl_down_sync(0xffffffffu, b0.x, 1);
            const float a1n = __shfl_down_sync(0xffffffffu, a1.x, 1), b1n = __shfl_down_sync(0xffffffffu, b1.x, 1);
            if (surf) {
                const int ci = (int)((lo & 1u) | ((lo >> 1) & 2u) | ((hi >> 0) & 4u) | ((hi & 1u) << 3) |
                                     ((lo & 2u) << 3) | ((lo & 8u) << 2) | ((hi & 8u) << 3) | ((hi & 2u) << 6));
                const unsigned long long row = c_mc_tri[ci];
                const int nv = mc_vertex_count(row);
                if (!EMIT) stage[staged + __popc(any & ((1u << lane) - 1u))] = mc_pack_cell(i, j, k, ci, n_row);
                if (EMIT) {
                    const float d[8] = {a0.x, a0n, a1n, a1.x, b0.x, b0n, b1n, b1.x};
                    float* o = xyz + 3 * (size_t)(base + n_row);
                    for (int q = 0; q < nv; q++) {
                        float v[3];
                        mc_edge_vertex(P, fm, i, j, k, (int)((row >> (4 * q)) & 0xFull), d, v);
                        o[3 * q] = v[0]; o[3 * q + 1] = v[1]; o[3 * q + 2] = v[2];
                    }
                }
                n_row += (unsigned int)nv;
            }
            if (!EMIT) {
                __syncwarp();
                staged += __popc(any);
                if (staged > 64) flush();
            }
        }
        a0 = a1; b0 = b1; s0 = s1; t0 = t1;
        a1 = a2; b1 = b2;
    }
    if (!EMIT) flush();
    if (!EMIT && cell_ok) row_count[((size_t)i * m + j) * nz + zc] = n_row;
}

/* pass 2 of the one-sweep path: one thread per listed surface cell */
__global__ void __launch_bounds__(128) k_mc_emit_list(GridParams g, McParams P, const float2* __restrict__ grid, int k_lo_all, int k_hi_all, int nz,
                                                      const unsigned long long* __restrict__ cells, const unsigned int* __restrict__ cell_counter,
                                                      const unsigned int* __restrict__ row_off, float* __restrict__ xyz,
                                                      unsigned int cell_cap, unsigned int vtx_cap) {
    /* launched BEFORE the host knows the counts (no sync between the sweep and the emit): the list length comes from the
     * device counter, and nothing is read or written beyond the capacities the buffers had at launch — the host checks
     * the counts afterwards and repeats the emit with larger buffers if either was exceeded */
    const unsigned int n_all = *cell_counter;
    const unsigned int n = n_all < cell_cap ? n_all : cell_cap;
    const unsigned int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const unsigned long long c = cells[q];
    const int i = (int)(c & 0xfff), j = (int)((c >> 12) & 0xfff), k = (int)((c >> 24) & 0xfff), ci = (int)((c >> 36) & 0xff);
    const unsigned int off = (unsigned int)(c >> 44);
    const int m = g.m;
    const int per = (k_hi_all - k_lo_all + nz) / nz;
    const int zc = (k - k_lo_all) / per;
    const size_t plane = (size_t)m * m;
    const float2* p = grid + (size_t)(k - g.ks0) * plane + (size_t)j * m + i;
    /* corners: 0 (i,j,k) 1 (i+1,j,k) 2 (i+1,j,k+1) 3 (i,j,k+1) 4 (i,j+1,k) 5 (i+1,j+1,k) 6 (i+1,j+1,k+1) 7 (i,j+1,k+1) */
    const float d[8] = {__ldg(p).x, __ldg(p + 1).x, __ldg(p + plane + 1).x, __ldg(p + plane).x,
                        __ldg(p + m).x, __ldg(p + m + 1).x, __ldg(p + plane + m + 1).x, __ldg(p + plane + m).x};
    const unsigned long long row = c_mc_tri[ci];
    const int nv = mc_vertex_count(row);
    const float fm = (float)m;
    const unsigned int v0 = row_off[((size_t)i * m + j) * nz + zc] + off;
    if (v0 + (unsigned int)nv > vtx_cap) return;
    float* o = xyz + 3 * (size_t)v0;
    for (int t = 0; t < nv; t++) {
        float v[3];
        mc_edge_vertex(P, fm, i, j, k, (int)((row >> (4 * t)) & 0xFull), d, v);
        o[3 * t] = v[0]; o[3 * t + 1] = v[1]; o[3 * t + 2] = v[2];
    }
}

/* marker points of SDF::visualize: (double)vertex + sdf_origin (sdf.cpp:354-356) */
__global__ void k_mesh_world(GridParams g, const float* __restrict__ xyz, int64_t n, double* __restrict__ world) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    world[3 * q] = (double)xyz[3 * q] + g.origin[0];
    world[3 * q + 1] = (double)xyz[3 * q + 1] + g.origin[1];
    world[3 * q + 2] = (double)xyz[3 * q + 2] + g.origin[2];
}
void launch_mesh_world(const GridParams& g, const float* xyz, int64_t n, double* world, cudaStream_t s) {
    if (n > 0) k_mesh_world<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(g, xyz, n, world);
}

/* cells of this handle: interior cells whose k lies in the owned range and whose k+1 layer is stored */
void mesh_k_range(const GridParams& g, int& k_lo, int& k_hi) {
    k_lo = g.ko0 < 1 ? 1 : g.ko0;
    k_hi = g.ko1 - 1;
    if (k_hi > g.m - 2) k_hi = g.m - 2;
    if (k_hi + 1 > g.ks1 - 1) k_hi = g.ks1 - 2;
}

size_t mesh_scan_bytes(int64_t n_rows) {
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, (unsigned int*)nullptr, (unsigned int*)nullptr, (int)n_rows);
    return bytes;
}

/* pass 1 + scan: row_count and row_off hold m*m*MC_ZSPLIT + 1 entries (the last row_count is 0, so
 * the last row_off is the total) */
int mesh_zsplit() { return MC_ZSPLIT; }
void launch_mesh_count(const GridParams& g, const McParams& P, const float2* grid, unsigned int* row_count, unsigned int* row_off,
                       void* scan_tmp, size_t scan_bytes, unsigned long long* cells, unsigned int* cell_counter, unsigned int cell_cap, cudaStream_t s) {
    const int64_t n_rows = (int64_t)g.m * g.m * MC_ZSPLIT + 1;
    cudaMemsetAsync(row_count, 0, (size_t)n_rows * sizeof(unsigned int), s);
    cudaMemsetAsync(cell_counter, 0, sizeof(unsigned int), s);
    int k_lo, k_hi;
    mesh_k_range(g, k_lo, k_hi);
    if (k_hi >= k_lo) {
        dim3 grid3((g.m - 2 + MC_TILE - 1) / MC_TILE, (g.m + MC_WARPS - 1) / MC_WARPS, MC_ZSPLIT);
        k_mc_sweep<false><<<grid3, MC_WARPS * 32, 0, s>>>(g, P, grid, k_lo, k_hi, row_count, nullptr, nullptr, cells, cell_counter, cell_cap);
    }
    cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, row_count, row_off, (int)n_rows, s);
}
void launch_mesh_emit(const GridParams& g, const McParams& P, const float2* grid, const unsigned int* row_off, float* xyz, cudaStream_t s) {
    int k_lo, k_hi;
    mesh_k_range(g, k_lo, k_hi);
    if (k_hi < k_lo) return;
    dim3 grid3((g.m - 2 + MC_TILE - 1) / MC_TILE, (g.m + MC_WARPS - 1) / MC_WARPS, MC_ZSPLIT);
    k_mc_sweep<true><<<grid3, MC_WARPS * 32, 0, s>>>(g, P, grid, k_lo, k_hi, nullptr, row_off, xyz, nullptr, nullptr, 0u);
}
/* pass 2 of the one-sweep path (n_cells <= the list's capacity) */
void launch_mesh_emit_list(const GridParams& g, const McParams& P, const float2* grid, const unsigned long long* cells, const unsigned int* cell_counter,
                           unsigned int n_threads, const unsigned int* row_off, float* xyz, unsigned int cell_cap, unsigned int vtx_cap, cudaStream_t s) {
    int k_lo, k_hi;
    mesh_k_range(g, k_lo, k_hi);
    if (k_hi < k_lo || n_threads == 0) return;
    k_mc_emit_list<<<(n_threads + 127) / 128, 128, 0, s>>>(g, P, grid, k_lo, k_hi, MC_ZSPLIT, cells, cell_counter, row_off, xyz, cell_cap, vtx_cap);
}

}  // namespace tsdf
