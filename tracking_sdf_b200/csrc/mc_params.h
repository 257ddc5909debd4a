/* parameters of the mesher (pcl::MarchingCubesSDF setters), shared by host code, kernels and the CPU emulation */
#pragma once
namespace tsdf {
struct McParams {
    float width, height, depth;       /* setBBox (marching_cubes_sdf.cpp:52-63): min_p = 0, max_p = extents */
    float iso;                        /* setIsoLevel; must be in [0, 1) (marching_cubes_sdf.cpp:248) */
};
}  // namespace tsdf
