/*
 * mc_core.cuh — per-cell arithmetic of the reference's mesher (pcl::MarchingCubesSDF,
 * marching_cubes_sdf.cpp:101-238), shared by the CUDA kernels (tsdf_mesh.cu).  Plain fp32, same
 * operations in the same order; build with -fmad=false.
 *
 * Cell (i,j,k), 1 <= i,j,k <= m-2 (interior voxels only, sdf.cpp:36-39); its corners
 * (marching_cubes_sdf.cpp:206-220):
 *   n :  0        1          2            3          4          5            6              7
 *        (i,j,k)  (i+1,j,k)  (i+1,j,k+1)  (i,j,k+1)  (i,j+1,k)  (i+1,j+1,k)  (i+1,j+1,k+1)  (i,j+1,k+1)
 * edges e = 0..11 join corners (0,1)(1,2)(2,3)(3,0)(4,5)(5,6)(6,7)(7,4)(0,4)(1,5)(2,6)(3,7), interpolated
 * from the first to the second (marching_cubes_sdf.cpp:146-169).
 */
#pragma once
#include "tsdf_core.cuh"
#include "mc_params.h"

namespace tsdf {

#if defined(__CUDACC__)
#define MC_TABLE_SPACE __constant__
#else
#define MC_TABLE_SPACE static const
#endif
MC_TABLE_SPACE unsigned long long c_mc_tri[256] = {
#include "mc_tables.inc"
};


/* configuration index of a cell: bit n set when corner n is below the iso level
 * (marching_cubes_sdf.cpp:107-115).  When any corner was never observed (W <= 0) the reference
 * fills all eight with corner 0's value (:221-241), i.e. configuration 0 or 255: no surface. */
TSDF_HD int mc_cube_index(const float* d, const float* w, float iso) {
    bool all = true;
#pragma unroll
    for (int n = 0; n < 8; n++) all = all && (w[n] > 0.0f);
    if (!all) return 0;
    int c = 0;
#pragma unroll
    for (int n = 0; n < 8; n++) c |= (d[n] < iso) ? (1 << n) : 0;
    return c;
}
TSDF_HD int mc_vertex_count(unsigned long long row) {
    int n = 0;
    while (n < 15 && ((row >> (4 * n)) & 0xFull) != 0xFull) n++;
    return n;
}

/* the n-th vertex of the triangle list of one cell: edge e of the cell at (i,j,k), interpolated
 * between its end points (marching_cubes_sdf.cpp:92-99, 119-140).  fm = (float)m. */
TSDF_HD void mc_edge_vertex(const McParams& P, float fm, int i, int j, int k, int e, const float* d, float out[3]) {
    const int ea = (e < 8) ? e : (e - 8), eb = (e < 8) ? (((e & 3) == 3) ? (e - 3) : (e + 1)) : (e - 4);
    /* corner position: centre + one cell step on the axes the corner is offset on (:127-140) */
    const float cx = 0.0f + (P.width - 0.0f) * (float)i / fm;
    const float cy = 0.0f + (P.height - 0.0f) * (float)j / fm;
    const float cz = 0.0f + (P.depth - 0.0f) * (float)k / fm;
    const float sx = (float)(cx + (P.width - 0.0f) / fm), sy = (float)(cy + (P.height - 0.0f) / fm), sz = (float)(cz + (P.depth - 0.0f) / fm);
    float p1[3], p2[3];
    {
        const int n = ea;
        p1[0] = (((n & 1) ^ ((n >> 1) & 1)) != 0) ? sx : cx; p1[1] = (n & 4) ? sy : cy; p1[2] = (n & 2) ? sz : cz;
    }
    {
        const int n = eb;
        p2[0] = (((n & 1) ^ ((n >> 1) & 1)) != 0) ? sx : cx; p2[1] = (n & 4) ? sy : cy; p2[2] = (n & 2) ? sz : cz;
    }
    const float mu = (P.iso - d[ea]) / (d[eb] - d[ea]);              /* :97 */
#pragma unroll
    for (int c = 0; c < 3; c++) out[c] = p1[c] + mu * (p2[c] - p1[c]);   /* :98 */
}

}  // namespace tsdf
