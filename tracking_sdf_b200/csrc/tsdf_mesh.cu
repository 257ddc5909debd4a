/*
 * tsdf_mesh.cu — the reference's mesher on the device:
 * pcl::MarchingCubesSDF::performReconstruction (marching_cubes_sdf.cpp:243-287) over the interior
 * cells, triangle soup in the reference's order (cells in (i,j,k) lexicographic order = its idx
 * order, triangles in table order), plus the marker post-processing of SDF::visualize
 * (sdf.cpp:352-385: + sdf_origin in double, one interpolate_color per vertex).
 *
 * The store is x fastest, the output order is k fastest, so a warp takes 32 consecutive i at a fixed
 * j and sweeps k: every load is a coalesced 256-byte row segment, each lane walks one (i,j) row of
 * cells in output order.  Two sweeps: count vertices per row -> exclusive scan over the m*m rows ->
 * emit.  HBM-bound: the store is read about once per sweep (the j+1 row is an L1/L2 hit).
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

template <bool EMIT>
__global__ void __launch_bounds__(MC_WARPS * 32) k_mc_sweep(GridParams g, McParams P, const float2* __restrict__ grid, int k_lo_all, int k_hi_all,
                                                            unsigned int* __restrict__ row_count, const unsigned int* __restrict__ row_off,
                                                            float* __restrict__ xyz) {
    const int m = g.m;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + lane;
    const int j = blockIdx.y * MC_WARPS + warp;
    if (j < 1 || j > m - 2) return;                                   /* warp-uniform */
    /* the k range is cut into gridDim.z chunks; a row's chunks are consecutive in the output */
    const int nz = (int)gridDim.z, zc = (int)blockIdx.z;
    const int per = (k_hi_all - k_lo_all + nz) / nz;
    const int k_lo = k_lo_all + zc * per;
    const int k_hi = (k_lo + per - 1 < k_hi_all) ? (k_lo + per - 1) : k_hi_all;
    if (k_lo > k_hi) return;
    const bool cell_ok = (i >= 1) && (i <= m - 2);
    const bool have = i < m, have1 = (i + 1) < m;
    const float fm = (float)m;
    const float2 zero = make_float2(0.0f, 0.0f);
    const size_t plane = (size_t)m * m;
    const float2* pj = grid + (size_t)j * m + i;                      /* (i, j, .) ; row j+1 is + m */
    /* software pipeline: the raw loads of layer k+2 are in flight while layer k+1 is shuffled and
     * cell k is evaluated.  raw = this lane's (i,j) and (i,j+1) voxels plus, on lane 31, the i+1 pair. */
    struct Raw { float2 a, b, xa, xb; };
    auto load_raw = [&](int k) {
        Raw r;
        const float2* p = pj + (size_t)(k - g.ks0) * plane;
        r.a = have ? __ldg(p) : zero;
        r.b = have ? __ldg(p + m) : zero;
        r.xa = zero; r.xb = zero;
        if (lane == 31 && have1) { r.xa = __ldg(p + 1); r.xb = __ldg(p + m + 1); }
        return r;
    };
    auto neighbours = [&](const Raw& r, float2& a1, float2& b1) {     /* (i+1,j), (i+1,j+1) */
        a1.x = __shfl_down_sync(0xffffffffu, r.a.x, 1); a1.y = __shfl_down_sync(0xffffffffu, r.a.y, 1);
        b1.x = __shfl_down_sync(0xffffffffu, r.b.x, 1); b1.y = __shfl_down_sync(0xffffffffu, r.b.y, 1);
        if (lane == 31) { a1 = r.xa; b1 = r.xb; }
    };
    Raw r0 = load_raw(k_lo), r1 = load_raw(k_lo + 1);
    float2 a = r0.a, b = r0.b, a1, b1;                                /* layer k:   (i,j) (i+1,j) (i,j+1) (i+1,j+1) */
    neighbours(r0, a1, b1);
    unsigned int n_row = 0;
    const unsigned int base = EMIT ? (cell_ok ? row_off[((size_t)i * m + j) * nz + zc] : 0u) : 0u;
    for (int k = k_lo; k <= k_hi; k++) {
        Raw r2;
        r2.a = r2.b = r2.xa = r2.xb = zero;
        if (k + 2 <= k_hi + 1) r2 = load_raw(k + 2);                  /* warp-uniform */
        const float2 c = r1.a, e = r1.b;                              /* layer k+1 */
        float2 c1, e1;
        neighbours(r1, c1, e1);
        if (cell_ok) {
            const float d[8] = {a.x, a1.x, c1.x, c.x, b.x, b1.x, e1.x, e.x};
            const float w[8] = {a.y, a1.y, c1.y, c.y, b.y, b1.y, e1.y, e.y};
            const int ci = mc_cube_index(d, w, P.iso);
            if (ci != 0 && ci != 255) {
                const unsigned long long row = c_mc_tri[ci];
                const int nv = mc_vertex_count(row);
                if (EMIT) {
                    float* o = xyz + 3 * (size_t)(base + n_row);
                    for (int q = 0; q < nv; q++) {
                        float v[3];
                        mc_edge_vertex(P, fm, i, j, k, (int)((row >> (4 * q)) & 0xFull), d, v);
                        o[3 * q] = v[0]; o[3 * q + 1] = v[1]; o[3 * q + 2] = v[2];
                    }
                }
                n_row += (unsigned int)nv;
            }
        }
        a = c; a1 = c1; b = e; b1 = e1;
        r1 = r2;
    }
    if (!EMIT && cell_ok) row_count[((size_t)i * m + j) * nz + zc] = n_row;
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
                       void* scan_tmp, size_t scan_bytes, cudaStream_t s) {
    const int64_t n_rows = (int64_t)g.m * g.m * MC_ZSPLIT + 1;
    cudaMemsetAsync(row_count, 0, (size_t)n_rows * sizeof(unsigned int), s);
    int k_lo, k_hi;
    mesh_k_range(g, k_lo, k_hi);
    if (k_hi >= k_lo) {
        dim3 grid3((g.m + 31) / 32, (g.m + MC_WARPS - 1) / MC_WARPS, MC_ZSPLIT);
        k_mc_sweep<false><<<grid3, MC_WARPS * 32, 0, s>>>(g, P, grid, k_lo, k_hi, row_count, nullptr, nullptr);
    }
    cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, row_count, row_off, (int)n_rows, s);
}
void launch_mesh_emit(const GridParams& g, const McParams& P, const float2* grid, const unsigned int* row_off, float* xyz, cudaStream_t s) {
    int k_lo, k_hi;
    mesh_k_range(g, k_lo, k_hi);
    if (k_hi < k_lo) return;
    dim3 grid3((g.m + 31) / 32, (g.m + MC_WARPS - 1) / MC_WARPS, MC_ZSPLIT);
    k_mc_sweep<true><<<grid3, MC_WARPS * 32, 0, s>>>(g, P, grid, k_lo, k_hi, nullptr, row_off, xyz);
}

}  // namespace tsdf
