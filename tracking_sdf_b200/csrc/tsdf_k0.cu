/*
 * tsdf_k0.cu — K0, the pre-processing the reference's node applies to every frame BEFORE the boundary
 * (sdf_reconstruction.cpp:37-49):
 *
 *     pcl::FastBilateralFilter<PointXYZRGB>                      (:38-41, PCL defaults sigma_s = 15 px, sigma_r = 0.05 m)
 *     pcl::IntegralImageNormalEstimation, AVERAGE_3D_GRADIENT,   (:43-49)
 *         MaxDepthChangeFactor 0.02, NormalSmoothingSize 10
 *
 * PCL is un-vendored third-party code (not under /root/reference, not in this image): PARITY UNPINNED at this
 * boundary.  What is built here is PCL's published algorithm — bilateral grid (splat, [1 2 1]/4 blur twice per axis,
 * trilinear slice) and the cross product of the summed central-difference 3-D gradients over a window whose size
 * is the chamfer distance to the nearest depth discontinuity, capped at the smoothing size — as ONE definition
 * shared with the oracle (oracle.cpp: k0_bilateral, k0_normals), reproduced here bit for bit: every sum runs in the
 * oracle's order (a grid cell adds its pixels in row-major order; window sums run row-major in double), no float
 * atomics, compiled -fmad=false.  Optional stage (tsdf_config.preprocess): the benchmark inputs are noise free and
 * use K1's own normals; K0 is for real (noisy) sensor data.
 *
 * Twelve small launches on the preprocessing stream, ahead of K1; they overlap the previous frame's tracking and
 * fusion like K1 does.
 */
#include "tsdf_internal.h"

namespace tsdf {

__device__ __forceinline__ bool k0_valid(float z) { return (z > 0.0f) && (z <= 3.402823466e+38f); }

struct K0Dims { int sw, sh, sd; float zmin; };
/* grid dimensions from the frame's depth range: mm[0] = ~bits(zmin) (max-reduced), mm[1] = bits(zmax) */
__device__ __forceinline__ bool k0_dims(const K0Params& P, const unsigned int* mm, K0Dims& d) {
    const unsigned int b0 = ~mm[0], b1 = mm[1];
    if (mm[1] == 0u) return false;                          /* no valid pixel */
    const float zmin = __uint_as_float(b0), zmax = __uint_as_float(b1);
    d.zmin = zmin;
    d.sw = (int)((float)(P.img_w - 1) / P.sigma_s) + 1 + 2 * K0_PAD;
    d.sh = (int)((float)(P.img_h - 1) / P.sigma_s) + 1 + 2 * K0_PAD;
    d.sd = (int)((zmax - zmin) / P.sigma_r) + 1 + 2 * K0_PAD;
    if (d.sd > K0_SD_MAX) d.sd = K0_SD_MAX;
    return true;
}

__global__ void __launch_bounds__(256) k0_minmax(K0Params P, const float* __restrict__ depth, unsigned int* mm) {
    __shared__ unsigned int s_lo[8], s_hi[8];
    const int n = P.img_w * P.img_h;
    unsigned int lo = 0u, hi = 0u;                          /* lo = max over ~bits = ~min bits */
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const float z = depth[p];
        if (k0_valid(z)) { const unsigned int b = __float_as_uint(z); lo = max(lo, ~b); hi = max(hi, b); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { lo = max(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
    if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) { lo = max(lo, s_lo[w]); hi = max(hi, s_hi[w]); }
        if (hi) { atomicMax(&mm[0], lo); atomicMax(&mm[1], hi); }     /* max is order independent: deterministic */
    }
}

/* depth bin of every pixel (-1 = no measurement): (int)((z - zmin) / sigma_r + 0.5f) + pad, clamped to the last bin */
__global__ void __launch_bounds__(256) k0_bins(K0Params P, const float* __restrict__ depth, const unsigned int* mm, short* __restrict__ bins) {
    const int n = P.img_w * P.img_h;
    K0Dims d;
    const bool okd = k0_dims(P, mm, d);
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const float z = depth[p];
        int bz = -1;
        if (okd && k0_valid(z)) {
            bz = (int)((z - d.zmin) / P.sigma_r + 0.5f) + K0_PAD;
            if (bz > d.sd - 1 - K0_PAD) bz = d.sd - 1 - K0_PAD;
        }
        bins[p] = (short)bz;
    }
}

/* splat: one thread per grid cell; the cell adds ITS pixels in row-major order (the oracle's order for that cell).
 * The pixels of a cell's column are the u (v) with (int)(u / sigma_s + 0.5f) == cu (cv): found once per thread as two
 * bit masks over the candidate range, then the depth bin of each (precomputed) decides.  Writes every cell of both
 * buffers (value / zero), so no clearing pass is needed. */
__global__ void __launch_bounds__(128) k0_splat(K0Params P, const float* __restrict__ depth, const short* __restrict__ bins,
                                                const unsigned int* mm, float2* ga, float2* gb) {
    K0Dims d;
    if (!k0_dims(P, mm, d)) return;
    const int sz = blockIdx.x * blockDim.x + threadIdx.x, sy = blockIdx.y, sx = blockIdx.z;
    if (sz >= d.sd || sy >= d.sh || sx >= d.sw) return;
    float sum = 0.0f, cnt = 0.0f;
    const int cu = sx - K0_PAD, cv = sy - K0_PAD;
    if (cu >= 0 && cv >= 0 && sz >= K0_PAD && sz <= d.sd - 1 - K0_PAD) {
        const int span = (int)P.sigma_s + 2;                 /* candidates: centre +- span (<= 31 each side for sigma_s <= 29) */
        const int uc = (int)((float)cu * P.sigma_s), vc = (int)((float)cv * P.sigma_s);
        unsigned long long mu = 0ull, mv = 0ull;
        for (int q = 0; q <= 2 * span; q++) {
            const int u = uc - span + q, v = vc - span + q;
            if (u >= 0 && u < P.img_w && (int)((float)u / P.sigma_s + 0.5f) == cu) mu |= 1ull << q;
            if (v >= 0 && v < P.img_h && (int)((float)v / P.sigma_s + 0.5f) == cv) mv |= 1ull << q;
        }
        for (int qv = 0; qv <= 2 * span; qv++) {
            if (!((mv >> qv) & 1ull)) continue;
            const size_t row = (size_t)(vc - span + qv) * P.img_w;
            for (int qu = 0; qu <= 2 * span; qu++) {
                if (!((mu >> qu) & 1ull)) continue;
                const size_t p = row + (uc - span + qu);
                if ((int)__ldg(&bins[p]) == sz) { sum = sum + __ldg(&depth[p]); cnt = cnt + 1.0f; }
            }
        }
    }
    const size_t c = ((size_t)sx * d.sh + sy) * d.sd + sz;
    ga[c] = make_float2(sum, cnt);
    gb[c] = make_float2(0.0f, 0.0f);
}

/* one [1 2 1]/4 pass along `dim` (0 x, 1 y, 2 z): interior cells only, faces keep their zeros */
__global__ void __launch_bounds__(128) k0_blur(K0Params P, const unsigned int* mm, const float2* src, float2* dst, int dim) {
    K0Dims d;
    if (!k0_dims(P, mm, d)) return;
    const int z = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, x = blockIdx.z;
    if (z < 1 || y < 1 || x < 1 || z >= d.sd - 1 || y >= d.sh - 1 || x >= d.sw - 1) return;
    const size_t c = ((size_t)x * d.sh + y) * d.sd + z;
    const size_t off = dim == 0 ? (size_t)d.sh * d.sd : dim == 1 ? (size_t)d.sd : 1;
    const float2 l = src[c - off], r = src[c + off], m = src[c];
    dst[c] = make_float2(((l.x + r.x) + 2.0f * m.x) / 4.0f, ((l.y + r.y) + 2.0f * m.y) / 4.0f);
}

/* slice: trilinear interpolation of (sum, count) at the pixel's grid position; filtered depth = sum / count */
__global__ void __launch_bounds__(256) k0_slice(K0Params P, const float* __restrict__ depth, const unsigned int* mm, const float2* g, float* zf) {
    const int u = blockIdx.x * 32 + (threadIdx.x & 31), v = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (u >= P.img_w || v >= P.img_h) return;
    const size_t p = (size_t)v * P.img_w + u;
    const float z = depth[p];
    K0Dims d;
    float out = __int_as_float(0x7fc00000);
    if (k0_valid(z) && k0_dims(P, mm, d)) {
        const float gx = (float)u / P.sigma_s + (float)K0_PAD, gy = (float)v / P.sigma_s + (float)K0_PAD;
        float gz = (z - d.zmin) / P.sigma_r + (float)K0_PAD;
        if (gz > (float)(d.sd - 1 - K0_PAD)) gz = (float)(d.sd - 1 - K0_PAD);
        const int x0 = (int)gx, y0 = (int)gy, z0 = (int)gz;
        const int x1 = min(x0 + 1, d.sw - 1), y1 = min(y0 + 1, d.sh - 1), z1 = min(z0 + 1, d.sd - 1);
        const float ax = gx - (float)x0, ay = gy - (float)y0, az = gz - (float)z0;
        auto at = [&](int x, int y, int zz) { return g[((size_t)x * d.sh + y) * d.sd + zz]; };
        const float2 c000 = at(x0, y0, z0), c100 = at(x1, y0, z0), c010 = at(x0, y1, z0), c110 = at(x1, y1, z0);
        const float2 c001 = at(x0, y0, z1), c101 = at(x1, y0, z1), c011 = at(x0, y1, z1), c111 = at(x1, y1, z1);
        const float w000 = ((1.0f - ax) * (1.0f - ay)) * (1.0f - az), w100 = (ax * (1.0f - ay)) * (1.0f - az);
        const float w010 = ((1.0f - ax) * ay) * (1.0f - az), w110 = (ax * ay) * (1.0f - az);
        const float w001 = ((1.0f - ax) * (1.0f - ay)) * az, w101 = (ax * (1.0f - ay)) * az;
        const float w011 = ((1.0f - ax) * ay) * az, w111 = (ax * ay) * az;
        float s = w000 * c000.x; s = s + w100 * c100.x; s = s + w010 * c010.x; s = s + w110 * c110.x;
        s = s + w001 * c001.x; s = s + w101 * c101.x; s = s + w011 * c011.x; s = s + w111 * c111.x;
        float n = w000 * c000.y; n = n + w100 * c100.y; n = n + w010 * c010.y; n = n + w110 * c110.y;
        n = n + w001 * c001.y; n = n + w101 * c101.y; n = n + w011 * c011.y; n = n + w111 * c111.y;
        out = s / n;
    }
    zf[p] = out;
}

/* depth-discontinuity map and central-difference 3-D gradients of the filtered cloud */
__global__ void __launch_bounds__(256) k0_grad(K0Params P, const float* __restrict__ zf, uint8_t* edge, float4* DX, float4* DY) {
    const int u = blockIdx.x * 32 + (threadIdx.x & 31), v = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int W = P.img_w, H = P.img_h;
    if (u >= W || v >= H) return;
    const size_t p = (size_t)v * W + u;
    const float qnan = __int_as_float(0x7fc00000);
    auto Z = [&](int uu, int vv) { return __ldg(&zf[(size_t)vv * W + uu]); };
    const float z = Z(u, v);
    /* pair (a -> b) initiated by a (a.u < W-1, a.v < H-1): marks both when either depth is missing or the step
     * exceeds the threshold computed from a's depth */
    auto pair_bad = [&](float za, float zb) {
        const float thr = P.max_depth_change * (fabsf(za) + 1.0f) * 2.0f;
        return !(za == za) || !(zb == zb) || fabsf(za - zb) > thr;
    };
    bool e = false;
    if (u < W - 1 && v < H - 1) e = pair_bad(z, Z(u + 1, v)) || pair_bad(z, Z(u, v + 1));
    if (u > 0 && v < H - 1) e = e || pair_bad(Z(u - 1, v), z);
    if (v > 0 && u < W - 1) e = e || pair_bad(Z(u, v - 1), z);
    edge[p] = e ? 1 : 0;
    auto point = [&](int uu, int vv, float& x, float& y, float& zz) {
        zz = Z(uu, vv);
        if (zz == zz) { x = ((float)uu - P.cx) * zz * P.inv_fx; y = ((float)vv - P.cy) * zz * P.inv_fy; }
        else { x = qnan; y = qnan; }
    };
    float4 dx = make_float4(qnan, qnan, qnan, 0.0f), dy = dx;
    if (u > 0 && u < W - 1) {
        float xr, yr, zr, xl, yl, zl;
        point(u + 1, v, xr, yr, zr); point(u - 1, v, xl, yl, zl);
        dx.x = xr - xl; dx.y = yr - yl; dx.z = zr - zl;
        dx.w = (dx.x == dx.x && dx.y == dx.y && dx.z == dx.z) ? 1.0f : 0.0f;
    }
    if (v > 0 && v < H - 1) {
        float xd, yd, zd, xu, yu, zu;
        point(u, v + 1, xd, yd, zd); point(u, v - 1, xu, yu, zu);
        dy.x = xd - xu; dy.y = yd - yu; dy.z = zd - zu;
        dy.w = (dy.x == dy.x && dy.y == dy.y && dy.z == dy.z) ? 1.0f : 0.0f;
    }
    DX[p] = dx; DY[p] = dy;
}

/* normals: window size = chamfer distance to the nearest discontinuity (capped), window sums in double, row-major */
__global__ void __launch_bounds__(256) k0_normals(K0Params P, const float* __restrict__ zf, const uint8_t* __restrict__ edge,
                                                  const float4* __restrict__ DX, const float4* __restrict__ DY, float4* nrm) {
    constexpr int R = K0_RADIUS, TW = 32 + 2 * R, TH = 8 + 2 * R;
    __shared__ uint8_t s_edge[TH][TW];
    const int W = P.img_w, H = P.img_h;
    const int u0 = blockIdx.x * 32 - R, v0 = blockIdx.y * 8 - R;
    int n_edge = 0;
    for (int t = threadIdx.x; t < TW * TH; t += 256) {
        const int tx = t % TW, ty = t / TW, uu = u0 + tx, vv = v0 + ty;
        const uint8_t e = (uu >= 0 && vv >= 0 && uu < W && vv < H) ? edge[(size_t)vv * W + uu] : 0;   /* outside the image: no discontinuity */
        s_edge[ty][tx] = e;
        n_edge += e;
    }
    const bool any_edge = __syncthreads_or(n_edge) != 0;     /* smooth regions (most blocks): every distance is the cap */
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    const int u = blockIdx.x * 32 + lx, v = blockIdx.y * 8 + ly;
    if (u >= W || v >= H) return;
    const size_t p = (size_t)v * W + u;
    const float qnan = __int_as_float(0x7fc00000);
    float4 out = make_float4(qnan, qnan, qnan, 0.0f);
    const float z = zf[p];
    if (z == z) {
        float dist = P.smoothing;
        const int Rs = (int)P.smoothing;                     /* <= K0_RADIUS (checked at create) */
        if (any_edge)
        for (int dv = -Rs; dv <= Rs; dv++)
            for (int du = -Rs; du <= Rs; du++) {
                if (!s_edge[ly + R + dv][lx + R + du]) continue;
                const int a = abs(du), b = abs(dv);
                const int mn = min(a, b), mx = max(a, b);
                dist = fminf(dist, 1.4f * (float)mn + 1.0f * (float)(mx - mn));
            }
        if (dist > 2.0f) {
            const int rw = (int)dist, r2 = rw / 2;
            double gx0 = 0, gx1 = 0, gx2 = 0, gy0 = 0, gy1 = 0, gy2 = 0;
            int cx_ = 0, cy_ = 0;
            for (int vv = v - r2; vv < v - r2 + rw; vv++) {
                if (vv < 0 || vv >= H) continue;
                for (int uu = u - r2; uu < u - r2 + rw; uu++) {
                    if (uu < 0 || uu >= W) continue;
                    const size_t q = (size_t)vv * W + uu;
                    const float4 a = __ldg(&DX[q]), b = __ldg(&DY[q]);
                    if (a.w != 0.0f) { gx0 = gx0 + (double)a.x; gx1 = gx1 + (double)a.y; gx2 = gx2 + (double)a.z; cx_++; }
                    if (b.w != 0.0f) { gy0 = gy0 + (double)b.x; gy1 = gy1 + (double)b.y; gy2 = gy2 + (double)b.z; cy_++; }
                }
            }
            if (cx_ != 0 && cy_ != 0) {
                const double nx = gy1 * gx2 - gy2 * gx1;
                const double ny = gy2 * gx0 - gy0 * gx2;
                const double nz = gy0 * gx1 - gy1 * gx0;
                const double len2 = (nx * nx + ny * ny) + nz * nz;
                if (len2 > 0.0) {
                    const double len = sqrt(len2);
                    float fx = (float)(nx / len), fy = (float)(ny / len), fz = (float)(nz / len);
                    const float px = ((float)u - P.cx) * z * P.inv_fx, py = ((float)v - P.cy) * z * P.inv_fy;
                    const float dotp = (fx * px + fy * py) + fz * z;     /* flipNormalTowardsViewpoint, viewpoint = origin */
                    if (dotp > 0.0f) { fx = -fx; fy = -fy; fz = -fz; }
                    out = make_float4(fx, fy, fz, 1.0f);
                }
            }
        }
    }
    nrm[p] = out;
}

int launch_k0(const K0Params& P, const K0Buffers& B, const float* depth, cudaStream_t s) {
    note_cuda(cudaMemsetAsync(B.minmax, 0, 2 * sizeof(unsigned int), s));
    k0_minmax<<<64, 256, 0, s>>>(P, depth, B.minmax);
    const int sw = (int)((float)(P.img_w - 1) / P.sigma_s) + 1 + 2 * K0_PAD, sh = (int)((float)(P.img_h - 1) / P.sigma_s) + 1 + 2 * K0_PAD;
    const dim3 gcells((K0_SD_MAX + 127) / 128, sh, sw);
    k0_bins<<<148 * 2, 256, 0, s>>>(P, depth, B.minmax, B.bins);
    k0_splat<<<gcells, 128, 0, s>>>(P, depth, B.bins, B.minmax, B.grid_a, B.grid_b);
    float2 *src = B.grid_a, *dst = B.grid_b;
    for (int dim = 0; dim < 3; dim++)
        for (int it = 0; it < 2; it++) {
            k0_blur<<<gcells, 128, 0, s>>>(P, B.minmax, src, dst, dim);
            float2* t = src; src = dst; dst = t;
        }
    const dim3 gpx((P.img_w + 31) / 32, (P.img_h + 7) / 8);
    k0_slice<<<gpx, 256, 0, s>>>(P, depth, B.minmax, src, B.zf);        /* six passes: the result is back in grid_a */
    k0_grad<<<gpx, 256, 0, s>>>(P, B.zf, B.edge, B.DX, B.DY);
    k0_normals<<<gpx, 256, 0, s>>>(P, B.zf, B.edge, B.DX, B.DY, B.normals);
    note_cuda(cudaGetLastError());
    return 12;
}

}  // namespace tsdf
