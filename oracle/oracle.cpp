/*
 * oracle.cpp — CPU ORACLE (test infrastructure, never shipped, never a fallback).
 *
 * A restatement, in plain C++ with no third-party dependency, of the arithmetic of
 * the reference's per-frame hot path.  Paths cited below are relative to
 * /root/reference/src/.  Pinned against the reference itself (see oracle.h): its hot-path
 * translation units, compiled unmodified over oracle/shim/, are compared with this file
 * function by function in tests/test_oracle_vs_ref.py.
 *
 * Build discipline: g++ -O2 -fopenmp -ffp-contract=off, no fast-math,
 * no -march (so no FMA contraction).  Every fp32/fp64 boundary of the reference is
 * reproduced: m_div_*, v_h2_*, i/j/k, volume, w, w_sum, sum_d, d_new, w_new and the
 * D/W update are float; all geometry and the normal equations are double.
 *
 * Third-party arithmetic that is NOT under /root/reference and is restated from its
 * published behaviour (Eigen 3.2-era, un-pinned by the reference's package.xml):
 *   - fixed-size matrix*vector / matrix*matrix coefficients accumulate left to right:
 *       ((a0*b0 + a1*b1) + a2*b2)                 [Eigen CoeffBasedProduct]
 *   - Vector3d::dot reduces as  c0 + (c1 + c2)    [Eigen redux_novec_unroller]
 *   - Matrix3d::inverse() = cofactor matrix * (1/det), det expanded along column 0
 *   - Matrix<double,6,6>::inverse() = partial-pivot LU whose column scaling and triangular
 *     solves multiply by ONE reciprocal per pivot (Eigen 3.2: `col /= pivot` is `*= 1/pivot`,
 *     triangular_solve_matrix uses `a = 1/tri(i,i)`); here: LU solve of A x = b with the same
 *     reciprocals (differs from inverse-then-multiply by a few ulp, see test_solve_and_pose_update)
 *   - Affine3d::rotation() (SVD polar of an already orthonormal matrix) = linear part
 */
#include "oracle.h"

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <climits>
#include <vector>
#include <omp.h>

namespace {

struct V3 { double x, y, z; };

/* Eigen fixed-size 3x3 * 3x1 coefficient: ((m0*v0 + m1*v1) + m2*v2) */
static inline V3 matvec(const double M[9], const V3& v) {
    V3 r;
    r.x = (M[0] * v.x + M[1] * v.y) + M[2] * v.z;
    r.y = (M[3] * v.x + M[4] * v.y) + M[5] * v.z;
    r.z = (M[6] * v.x + M[7] * v.y) + M[8] * v.z;
    return r;
}
static inline void matmul3(const double A[9], const double B[9], double C[9]) {
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++)
            C[3 * r + c] = (A[3 * r + 0] * B[0 + c] + A[3 * r + 1] * B[3 + c]) + A[3 * r + 2] * B[6 + c];
}
/* Matrix3d::inverse(): cofactors times reciprocal determinant */
static inline void inverse3(const double M[9], double inv[9]) {
    double c00 = M[4] * M[8] - M[5] * M[7];
    double c01 = M[5] * M[6] - M[3] * M[8];
    double c02 = M[3] * M[7] - M[4] * M[6];
    /* Eigen compute_inverse<.,.,3>: det = (cofactors_col0 .* col(0)).sum(), a 3-term redux c0 + (c1 + c2);
     * pinned against the reference compiled over oracle/shim (tests/test_oracle_vs_ref.py) */
    double c10 = M[7] * M[2] - M[8] * M[1];
    double c20 = M[1] * M[5] - M[2] * M[4];
    double det = c00 * M[0] + (c10 * M[3] + c20 * M[6]);
    double id = 1.0 / det;
    inv[0] = c00 * id;
    inv[1] = c10 * id;
    inv[2] = c20 * id;
    inv[3] = c01 * id;
    inv[4] = (M[0] * M[8] - M[2] * M[6]) * id;
    inv[5] = (M[2] * M[3] - M[0] * M[5]) * id;
    inv[6] = c02 * id;
    inv[7] = (M[1] * M[6] - M[0] * M[7]) * id;
    inv[8] = (M[0] * M[4] - M[1] * M[3]) * id;
}

/* x86 cvttsd2si / cvttss2si semantics of the reference's (int) casts: truncate toward
 * zero; NaN and out-of-range give INT_MIN ("integer indefinite"). */
static inline int cast_int(double v) {
    if (!(v > -2147483649.0 && v < 2147483648.0)) return INT_MIN;
    return (int)v;
}
static inline int cast_int(float v) {
    if (!(v >= -2147483648.0f && v < 2147483648.0f)) return INT_MIN;
    return (int)v;
}

struct Oracle {
    orc_config cfg;
    int m;
    int64_t m_squared, number_of_voxels;
    float m_div_width, m_div_height, m_div_depth;        /* sdf.cpp:19-21 */
    float v_h, w_h, v_h2, w_h2;                          /* camera_tracking.cpp:11-14 */
    float v_h2_width, v_h2_height, v_h2_depth;           /* camera_tracking.cpp:15-17 */
    double origin[3];
    std::vector<float> D, W;                              /* sdf.cpp:10,13 */
    std::vector<float> CW, R, G, B;                       /* sdf.cpp:14-17 Color_W,R,G,B (allocated by orc_enable_color) */
    std::vector<V3> global_coords;                        /* sdf.cpp:11 (optional) */
    double K[9];
    bool isKFilled;
    double rot[9], rot_inv[9], trans[3], rot_inv_trans[3];/* camera_tracking.h:42-49 */
    std::vector<float> mesh;                              /* last orc_mesh result: xyz per vertex */
    /* scratch */
    std::vector<float> cloud, normals;
    std::vector<float> pxJ, pxPsi;
    std::vector<uint8_t> pxFlag;
};

/* sdf.h:113-127 — z-fastest linear index, -1 when out of range */
static inline int64_t get_array_index(const Oracle* o, int i, int j, int k) {
    if (i < 0 || j < 0 || k < 0) return -1;
    if (i >= o->m || j >= o->m || k >= o->m) return -1;
    return o->m_squared * i + (int64_t)o->m * j + k;
}
/* sdf.h:132-136 */
static inline void get_voxel_coordinates_idx(const Oracle* o, int64_t idx, int ijk[3]) {
    ijk[1] = (int)((idx % o->m_squared) / o->m);
    ijk[0] = (int)(idx / o->m_squared);
    ijk[2] = (int)(idx % o->m);
}
/* sdf.h:143-147 — world -> continuous voxel coordinates (fp32 scale promoted to double) */
static inline V3 get_voxel_coordinates(const Oracle* o, const V3& g) {
    V3 v;
    v.x = ((g.x - o->origin[0]) * o->m_div_width - 0.5);
    v.y = ((g.y - o->origin[1]) * o->m_div_height - 0.5);
    v.z = ((g.z - o->origin[2]) * o->m_div_depth - 0.5);
    return v;
}
/* sdf.h:153-157 — voxel centre; the quotient extent/(float)m is fp32 */
static inline V3 get_global_coordinates(const Oracle* o, const int ijk[3]) {
    V3 g;
    g.x = (o->cfg.width / ((float)o->m)) * (ijk[0] + 0.5) + o->origin[0];
    g.y = (o->cfg.height / ((float)o->m)) * (ijk[1] + 0.5) + o->origin[1];
    g.z = (o->cfg.depth / ((float)o->m)) * (ijk[2] + 0.5) + o->origin[2];
    return g;
}

/* sdf.cpp:127-163 — inverse-L1-distance weighted mean of the <=8 neighbours with W>0 */
static float interpolate_distance(const Oracle* o, const V3& vc, bool& is_interpolated) {
    float i = (float)vc.x;
    float j = (float)vc.y;
    float k = (float)vc.z;
    float w_sum = 0.0f;
    float sum_d = 0.0f;
    float w = 0;
    float volume;
    is_interpolated = false;
    const int bi = cast_int(i), bj = cast_int(j), bk = cast_int(k);
    for (int i_offset = 0; i_offset < 2; i_offset++) {
        for (int j_offset = 0; j_offset < 2; j_offset++) {
            for (int k_offset = 0; k_offset < 2; k_offset++) {
                /* INT_MIN + offset stays negative -> get_array_index gives -1 */
                int ci = bi + i_offset, cj = bj + j_offset, ck = bk + k_offset;
                volume = std::fabs((float)ci - i) + std::fabs((float)cj - j) + std::fabs((float)ck - k);
                int64_t a_idx = get_array_index(o, ci, cj, ck);
                if (a_idx != -1) {
                    if (o->W[a_idx] > 0) {
                        is_interpolated = true;
                        if (volume < 0.00001) {          /* double literal, sdf.cpp:151 */
                            return o->D[a_idx];
                        }
                        w = (float)(1.0 / volume);        /* double divide -> float, sdf.cpp:154 */
                        w_sum += w;
                        sum_d += w * o->D[a_idx];
                    }
                }
            }
        }
    }
    return sum_d / w_sum;                                 /* 0/0 = NaN when none qualified */
}

/* camera_tracking.cpp:59-65 */
static void set_camera_transformation(Oracle* o, const double R[9], const double t[3]) {
    double Rc[9], tc[3];
    memcpy(Rc, R, sizeof Rc);
    memcpy(tc, t, sizeof tc);
    memcpy(o->rot, Rc, sizeof Rc);
    inverse3(o->rot, o->rot_inv);
    memcpy(o->trans, tc, sizeof tc);
    V3 tv = {tc[0], tc[1], tc[2]};
    V3 rt = matvec(o->rot_inv, tv);
    o->rot_inv_trans[0] = -1 * rt.x;
    o->rot_inv_trans[1] = -1 * rt.y;
    o->rot_inv_trans[2] = -1 * rt.z;
}

/* ---- K1: back-projection + normals.  NOT in the reference (done upstream by the ROS
 * depth_image_proc nodelet and PCL): one definition shared by oracle
 * and GPU, fp32 like depth_image_proc (x = (u - cx) * depth * (1/fx)). ---- */
static void backproject(const Oracle* o, const float* depth, float* cloud, float* normals) {
    const int Wd = o->cfg.image_width, Hd = o->cfg.image_height;
    const float cxf = (float)o->K[2], cyf = (float)o->K[5];
    const float inv_fx = 1.0f / (float)o->K[0], inv_fy = 1.0f / (float)o->K[4];
    const float nanf_ = std::nanf("");
#pragma omp parallel for
    for (int v = 0; v < Hd; v++) {
        for (int u = 0; u < Wd; u++) {
            float z = depth[(size_t)v * Wd + u];
            float* p = cloud + 3 * ((size_t)v * Wd + u);
            if (!(std::isfinite(z) && z > 0.0f)) {
                p[0] = p[1] = p[2] = nanf_;
            } else {
                p[0] = ((float)u - cxf) * z * inv_fx;
                p[1] = ((float)v - cyf) * z * inv_fy;
                p[2] = z;
            }
        }
    }
    if (!normals) return;
#pragma omp parallel for
    for (int v = 0; v < Hd; v++) {
        for (int u = 0; u < Wd; u++) {
            float* n = normals + 3 * ((size_t)v * Wd + u);
            n[0] = n[1] = n[2] = nanf_;
            if (u == 0 || v == 0 || u == Wd - 1 || v == Hd - 1) continue;
            const float* pc = cloud + 3 * ((size_t)v * Wd + u);
            const float* pl = cloud + 3 * ((size_t)v * Wd + (u - 1));
            const float* pr = cloud + 3 * ((size_t)v * Wd + (u + 1));
            const float* pu = cloud + 3 * ((size_t)(v - 1) * Wd + u);
            const float* pd = cloud + 3 * ((size_t)(v + 1) * Wd + u);
            if (std::isnan(pc[2]) || std::isnan(pl[2]) || std::isnan(pr[2]) || std::isnan(pu[2]) || std::isnan(pd[2]))
                continue;
            /* depth-discontinuity gate, after PCL MaxDepthChangeFactor = 0.02
             * (sdf_reconstruction.cpp:46) */
            float thr = 0.02f * pc[2];
            if (std::fabs(pl[2] - pc[2]) > thr || std::fabs(pr[2] - pc[2]) > thr ||
                std::fabs(pu[2] - pc[2]) > thr || std::fabs(pd[2] - pc[2]) > thr)
                continue;
            float ax = pr[0] - pl[0], ay = pr[1] - pl[1], az = pr[2] - pl[2];
            float bx = pd[0] - pu[0], by = pd[1] - pu[1], bz = pd[2] - pu[2];
            float nx = ay * bz - az * by;
            float ny = az * bx - ax * bz;
            float nz = ax * by - ay * bx;
            float len = std::sqrt((nx * nx + ny * ny) + nz * nz);
            if (!(len > 0.0f)) continue;
            nx = nx / len; ny = ny / len; nz = nz / len;
            /* orient toward the viewpoint (PCL flipNormalTowardsViewpoint, vp = 0) */
            float dotp = (nx * pc[0] + ny * pc[1]) + nz * pc[2];
            if (dotp > 0.0f) { nx = -nx; ny = -ny; nz = -nz; }
            n[0] = nx; n[1] = ny; n[2] = nz;
        }
    }
}

/* ---- K0: the node's pre-processing (sdf_reconstruction.cpp:37-49).  NOT in the reference (PCL); see oracle.h.
 * Parameters = the node's call sites / PCL defaults. */
static const float K0_SIGMA_S = 15.0f, K0_SIGMA_R = 0.05f;        /* pcl::FastBilateralFilter defaults (:38-41 sets none) */
static const int K0_PAD = 2, K0_SD_MAX = 256;
static const float K0_MAX_DEPTH_CHANGE = 0.02f;                    /* ne.setMaxDepthChangeFactor(0.02f), :46 */
static const float K0_SMOOTHING = 10.0f;                           /* ne.setNormalSmoothingSize(10.0f),   :47 */

struct K0Grid { int sw, sh, sd; float zmin; std::vector<float> a, b; };   /* cells hold (sum z, count) interleaved */

/* fast bilateral filter of the depth channel (bilateral grid, Paris & Durand; PCL fast_bilateral.hpp) */
static void k0_bilateral(const Oracle* o, const float* depth, float* out) {
    const int Wd = o->cfg.image_width, Hd = o->cfg.image_height;
    float zmin = 3.402823466e+38f, zmax = -3.402823466e+38f;
    for (size_t p = 0; p < (size_t)Wd * Hd; p++) {
        const float z = depth[p];
        if (std::isfinite(z) && z > 0.0f) { zmin = std::fmin(zmin, z); zmax = std::fmax(zmax, z); }
    }
    for (size_t p = 0; p < (size_t)Wd * Hd; p++) out[p] = std::nanf("");
    if (!(zmax >= zmin)) return;                                     /* no valid pixel */
    K0Grid G;
    G.zmin = zmin;
    G.sw = (int)((float)(Wd - 1) / K0_SIGMA_S) + 1 + 2 * K0_PAD;
    G.sh = (int)((float)(Hd - 1) / K0_SIGMA_S) + 1 + 2 * K0_PAD;
    G.sd = (int)((zmax - zmin) / K0_SIGMA_R) + 1 + 2 * K0_PAD;
    if (G.sd > K0_SD_MAX) G.sd = K0_SD_MAX;                          /* depth ranges beyond 12.5 m share the last bins */
    const size_t nc = (size_t)G.sw * G.sh * G.sd;
    G.a.assign(2 * nc, 0.0f); G.b.assign(2 * nc, 0.0f);
    auto cell = [&](int x, int y, int z) { return 2 * (((size_t)x * G.sh + y) * G.sd + z); };
    /* splat, pixels in row-major order */
    for (int v = 0; v < Hd; v++)
        for (int u = 0; u < Wd; u++) {
            const float z = depth[(size_t)v * Wd + u];
            if (!(std::isfinite(z) && z > 0.0f)) continue;
            const int sx = (int)((float)u / K0_SIGMA_S + 0.5f) + K0_PAD;
            const int sy = (int)((float)v / K0_SIGMA_S + 0.5f) + K0_PAD;
            int sz = (int)((z - zmin) / K0_SIGMA_R + 0.5f) + K0_PAD;
            if (sz > G.sd - 1 - K0_PAD) sz = G.sd - 1 - K0_PAD;
            const size_t c = cell(sx, sy, sz);
            G.a[c] = G.a[c] + z; G.a[c + 1] = G.a[c + 1] + 1.0f;
        }
    /* blur: along x, y, z in turn, two passes each, interior cells only: (l + r + 2 c) / 4 */
    const long off[3] = {(long)G.sh * G.sd, (long)G.sd, 1};
    for (int dim = 0; dim < 3; dim++)
        for (int it = 0; it < 2; it++) {
            std::swap(G.a, G.b);
            for (int x = 1; x < G.sw - 1; x++)
                for (int y = 1; y < G.sh - 1; y++)
                    for (int z = 1; z < G.sd - 1; z++) {
                        const size_t c = cell(x, y, z);
                        for (int q = 0; q < 2; q++)
                            G.a[c + q] = ((G.b[c - 2 * off[dim] + q] + G.b[c + 2 * off[dim] + q]) + 2.0f * G.b[c + q]) / 4.0f;
                    }
            /* cells on the faces are never written; they hold zeros (the padding is never splatted into) */
        }
    /* slice: trilinear interpolation of (sum, count) at the pixel's grid position */
    for (int v = 0; v < Hd; v++)
        for (int u = 0; u < Wd; u++) {
            const float z = depth[(size_t)v * Wd + u];
            if (!(std::isfinite(z) && z > 0.0f)) continue;
            const float gx = (float)u / K0_SIGMA_S + (float)K0_PAD, gy = (float)v / K0_SIGMA_S + (float)K0_PAD;
            float gz = (z - zmin) / K0_SIGMA_R + (float)K0_PAD;
            if (gz > (float)(G.sd - 1 - K0_PAD)) gz = (float)(G.sd - 1 - K0_PAD);
            const int x0 = (int)gx, y0 = (int)gy, z0 = (int)gz;
            const int x1 = x0 + 1 > G.sw - 1 ? G.sw - 1 : x0 + 1, y1 = y0 + 1 > G.sh - 1 ? G.sh - 1 : y0 + 1, z1 = z0 + 1 > G.sd - 1 ? G.sd - 1 : z0 + 1;
            const float ax = gx - (float)x0, ay = gy - (float)y0, az = gz - (float)z0;
            float acc[2];
            for (int q = 0; q < 2; q++) {
                float s = ((1.0f - ax) * (1.0f - ay)) * (1.0f - az) * G.a[cell(x0, y0, z0) + q];
                s = s + (ax * (1.0f - ay)) * (1.0f - az) * G.a[cell(x1, y0, z0) + q];
                s = s + ((1.0f - ax) * ay) * (1.0f - az) * G.a[cell(x0, y1, z0) + q];
                s = s + (ax * ay) * (1.0f - az) * G.a[cell(x1, y1, z0) + q];
                s = s + ((1.0f - ax) * (1.0f - ay)) * az * G.a[cell(x0, y0, z1) + q];
                s = s + (ax * (1.0f - ay)) * az * G.a[cell(x1, y0, z1) + q];
                s = s + ((1.0f - ax) * ay) * az * G.a[cell(x0, y1, z1) + q];
                s = s + (ax * ay) * az * G.a[cell(x1, y1, z1) + q];
                acc[q] = s;
            }
            out[(size_t)v * Wd + u] = acc[0] / acc[1];
        }
}

/* average-3D-gradient normals over an adaptive window (PCL integral_image_normal.hpp, AVERAGE_3D_GRADIENT) */
static void k0_normals(const Oracle* o, const float* zf, float* normals) {
    const int Wd = o->cfg.image_width, Hd = o->cfg.image_height;
    const float cxf = (float)o->K[2], cyf = (float)o->K[5];
    const float inv_fx = 1.0f / (float)o->K[0], inv_fy = 1.0f / (float)o->K[4];
    const float nanf_ = std::nanf("");
    const size_t N = (size_t)Wd * Hd;
    std::vector<float> P(3 * N), DX(3 * N), DY(3 * N);
    std::vector<uint8_t> edge(N, 0);
    for (int v = 0; v < Hd; v++)
        for (int u = 0; u < Wd; u++) {
            const size_t p = (size_t)v * Wd + u;
            const float z = zf[p];
            if (z == z) { P[3 * p] = ((float)u - cxf) * z * inv_fx; P[3 * p + 1] = ((float)v - cyf) * z * inv_fy; P[3 * p + 2] = z; }
            else P[3 * p] = P[3 * p + 1] = P[3 * p + 2] = nanf_;
        }
    /* depth discontinuities (PCL's depthChangeMap): every pixel but the last row / column compares itself with its
     * right and lower neighbour; a missing depth on either side or a step above the threshold (computed from the
     * pixel's own depth) marks BOTH pixels of the pair */
    auto pair_bad = [&](float za, float zb) {
        const float thr = K0_MAX_DEPTH_CHANGE * (std::fabs(za) + 1.0f) * 2.0f;
        return !(za == za) || !(zb == zb) || std::fabs(za - zb) > thr;
    };
    for (int v = 0; v < Hd - 1; v++)
        for (int u = 0; u < Wd - 1; u++) {
            const size_t p = (size_t)v * Wd + u;
            if (pair_bad(zf[p], zf[p + 1])) { edge[p] = 1; edge[p + 1] = 1; }
            if (pair_bad(zf[p], zf[p + Wd])) { edge[p] = 1; edge[p + Wd] = 1; }
        }
    /* central-difference 3-D gradients */
    for (int v = 0; v < Hd; v++)
        for (int u = 0; u < Wd; u++) {
            const size_t p = (size_t)v * Wd + u;
            for (int c = 0; c < 3; c++) {
                DX[3 * p + c] = (u > 0 && u < Wd - 1) ? P[3 * (p + 1) + c] - P[3 * (p - 1) + c] : nanf_;
                DY[3 * p + c] = (v > 0 && v < Hd - 1) ? P[3 * (p + Wd) + c] - P[3 * (p - Wd) + c] : nanf_;
            }
        }
    const int R = (int)K0_SMOOTHING;
#pragma omp parallel for schedule(dynamic, 4)
    for (int v = 0; v < Hd; v++)
        for (int u = 0; u < Wd; u++) {
            const size_t p = (size_t)v * Wd + u;
            float* n = normals + 3 * p;
            n[0] = n[1] = n[2] = nanf_;
            if (!(zf[p] == zf[p])) continue;
            /* chamfer (1 / 1.4) distance to the nearest discontinuity within the smoothing radius, capped */
            float dist = K0_SMOOTHING;
            for (int dv = -R; dv <= R; dv++)
                for (int du = -R; du <= R; du++) {
                    const int uu = u + du, vv = v + dv;
                    if (uu < 0 || vv < 0 || uu >= Wd || vv >= Hd) continue;
                    if (!edge[(size_t)vv * Wd + uu]) continue;
                    const int a = du < 0 ? -du : du, b = dv < 0 ? -dv : dv;
                    const int mn = a < b ? a : b, mx = a < b ? b : a;
                    const float d = 1.4f * (float)mn + 1.0f * (float)(mx - mn);
                    dist = std::fmin(dist, d);
                }
            if (!(dist > 2.0f)) continue;                           /* PCL: smoothing > 2.0f */
            const int rw = (int)dist, r2 = rw / 2;                   /* setRectSize(w, h): width w, starts w/2 left of the pixel */
            double gx[3] = {0, 0, 0}, gy[3] = {0, 0, 0};
            int cx_ = 0, cy_ = 0;
            for (int vv = v - r2; vv < v - r2 + rw; vv++)
                for (int uu = u - r2; uu < u - r2 + rw; uu++) {
                    if (uu < 0 || vv < 0 || uu >= Wd || vv >= Hd) continue;
                    const size_t q = (size_t)vv * Wd + uu;
                    if (DX[3 * q] == DX[3 * q] && DX[3 * q + 1] == DX[3 * q + 1] && DX[3 * q + 2] == DX[3 * q + 2]) {
                        gx[0] = gx[0] + (double)DX[3 * q]; gx[1] = gx[1] + (double)DX[3 * q + 1]; gx[2] = gx[2] + (double)DX[3 * q + 2]; cx_++;
                    }
                    if (DY[3 * q] == DY[3 * q] && DY[3 * q + 1] == DY[3 * q + 1] && DY[3 * q + 2] == DY[3 * q + 2]) {
                        gy[0] = gy[0] + (double)DY[3 * q]; gy[1] = gy[1] + (double)DY[3 * q + 1]; gy[2] = gy[2] + (double)DY[3 * q + 2]; cy_++;
                    }
                }
            if (cx_ == 0 || cy_ == 0) continue;
            /* normal = gradient_y x gradient_x, normalised */
            const double nx = gy[1] * gx[2] - gy[2] * gx[1];
            const double ny = gy[2] * gx[0] - gy[0] * gx[2];
            const double nz = gy[0] * gx[1] - gy[1] * gx[0];
            const double len2 = (nx * nx + ny * ny) + nz * nz;
            if (!(len2 > 0.0)) continue;
            const double len = std::sqrt(len2);
            float fx = (float)(nx / len), fy = (float)(ny / len), fz = (float)(nz / len);
            /* flipNormalTowardsViewpoint, viewpoint = origin */
            const float dotp = (fx * P[3 * p] + fy * P[3 * p + 1]) + fz * P[3 * p + 2];
            if (dotp > 0.0f) { fx = -fx; fy = -fy; fz = -fz; }
            n[0] = fx; n[1] = fy; n[2] = fz;
        }
}

/* sdf.cpp:224-305; the colour part (:294-304) runs when an RGB image is given (rgb: h*w*3 bytes, the
 * r,g,b of pcl::PointXYZRGB at (col,row)) */
static int64_t fuse_cloud(Oracle* o, const float* cloud, const float* normals, const uint8_t* rgb = nullptr) {
    const int Wd = o->cfg.image_width, Hd = o->cfg.image_height;
    const float distance_epsilon = o->cfg.distance_epsilon, distance_delta = o->cfg.distance_delta;
    const int metric = o->cfg.metric;
    int64_t n_updated = 0;
    const bool table = !o->global_coords.empty();
#pragma omp parallel for reduction(+ : n_updated)
    for (int64_t idx = 0; idx < o->number_of_voxels; idx++) {
        V3 g;
        if (table) {
            g = o->global_coords[idx];                                   /* sdf.cpp:244 */
        } else {
            int ijk[3];
            get_voxel_coordinates_idx(o, idx, ijk);
            g = get_global_coordinates(o, ijk);
        }
        /* camera_tracking.cpp:51-54 */
        V3 cam = matvec(o->rot_inv, g);
        cam.x = cam.x + o->rot_inv_trans[0];
        cam.y = cam.y + o->rot_inv_trans[1];
        cam.z = cam.z + o->rot_inv_trans[2];
        if (cam.z < 0) continue;                                          /* sdf.cpp:247 */
        /* camera_tracking.cpp:40-47 */
        V3 ij = matvec(o->K, cam);
        double iu = ij.x / ij.z, iv = ij.y / ij.z;
        int i_image = cast_int(iu), j_image = cast_int(iv);               /* sdf.cpp:251-252 */
        if (i_image >= Wd || j_image >= Hd || i_image < 0 || j_image < 0) continue;
        const float* pt = cloud + 3 * ((size_t)j_image * Wd + i_image);   /* at(col,row) */
        float d_new;
        if (metric == 0) {
            const float* nm = normals + 3 * ((size_t)j_image * Wd + i_image);
            if (std::isnan(pt[0]) || std::isnan(pt[1]) || std::isnan(nm[0]) || std::isnan(nm[1]) || std::isnan(nm[2]))
                continue;                                                 /* sdf.cpp:260 */
            /* sdf.h:177-181 : (p_img - p_voxel) . n, Eigen dot = c0 + (c1 + c2) */
            double dx = (double)pt[0] - cam.x, dy = (double)pt[1] - cam.y, dz = (double)pt[2] - cam.z;
            double pointToPlane = dx * (double)nm[0] + (dy * (double)nm[1] + dz * (double)nm[2]);
            d_new = (float)pointToPlane;                                  /* sdf.cpp:274 */
        } else {
            if (std::isnan(pt[0]) || std::isnan(pt[1])) continue;
            /* sdf.h:169-172 : voxel depth - observed depth */
            double pointToPoint = cam.z - (double)pt[2];
            d_new = (float)pointToPoint;
        }
        float w_new = 1.0f;                                               /* sdf.cpp:276 */
        if (d_new >= distance_epsilon && d_new <= distance_delta) {
            w_new = (float)std::exp(-0.5 * (d_new - distance_epsilon) * (d_new - distance_epsilon));
        }
        if (d_new > distance_delta) continue;                             /* sdf.cpp:280-283 */
        if (d_new < -distance_delta) d_new = -distance_delta;             /* sdf.cpp:285-287 */
        float w_old = o->W[idx];
        o->W[idx] = w_old + w_new;                                        /* sdf.cpp:290 */
        o->D[idx] = (w_old * o->D[idx] + w_new * d_new) / o->W[idx];      /* sdf.cpp:292 */
        n_updated++;
        if (rgb && metric == 0) {
            /* sdf.cpp:294-304.  cam_vect = (0,0,1) (:235); Eigen 3-vector reductions are c0 + (c1 + c2) */
            const float* nm = normals + 3 * ((size_t)j_image * Wd + i_image);
            const double nx = (double)nm[0], ny = (double)nm[1], nz = (double)nm[2];
            const double dotv = 0.0 * nx + (0.0 * ny + 1.0 * nz);
            const double norm = std::sqrt(nx * nx + (ny * ny + nz * nz));
            const double cosine = std::fabs(dotv) / norm;                 /* :294 */
            const uint8_t* c = rgb + 3 * ((size_t)j_image * Wd + i_image);
            w_old = o->CW[idx];                                           /* :298 */
            w_new = (float)(w_new * cosine);                              /* :299 float = float * double */
            o->CW[idx] = w_old + w_new;                                   /* :300 */
            o->R[idx] = (w_old * o->R[idx] + w_new * c[0]) / o->CW[idx];  /* :302-304, uint8 promoted to int then float */
            o->G[idx] = (w_old * o->G[idx] + w_new * c[1]) / o->CW[idx];
            o->B[idx] = (w_old * o->B[idx] + w_new * c[2]) / o->CW[idx];
        }
    }
    return n_updated;
}

struct PerturbedRot { double r[6][9]; };   /* r1p r1m r2p r2m r3p r3m */

/* camera_tracking.cpp:92-145 */
static void build_perturbed(const Oracle* o, PerturbedRot& P) {
    const double w_h = (double)o->w_h;                    /* float promoted on assignment */
    double Rd[9];
    auto ident = [&]() { Rd[0] = 1; Rd[1] = 0; Rd[2] = 0; Rd[3] = 0; Rd[4] = 1; Rd[5] = 0; Rd[6] = 0; Rd[7] = 0; Rd[8] = 1; };
    ident(); Rd[5] = -w_h; Rd[7] = w_h;  matmul3(Rd, o->rot, P.r[0]);
    ident(); Rd[5] = w_h;  Rd[7] = -w_h; matmul3(Rd, o->rot, P.r[1]);
    ident(); Rd[2] = w_h;  Rd[6] = -w_h; matmul3(Rd, o->rot, P.r[2]);
    ident(); Rd[2] = -w_h; Rd[6] = w_h;  matmul3(Rd, o->rot, P.r[3]);
    ident(); Rd[1] = -w_h; Rd[3] = w_h;  matmul3(Rd, o->rot, P.r[4]);
    ident(); Rd[1] = w_h;  Rd[3] = -w_h; matmul3(Rd, o->rot, P.r[5]);
}

/* camera_tracking.cpp:246-363.  Returns flag: 1 ok, 2 out of volume (TRAP 5: treated as
 * invalid instead of reproducing the stale-state re-add), 3 a sample was not interpolated. */
static int get_partial_derivative(const Oracle* o, const PerturbedRot& P, const V3& cp, float J[6], float& sdf_val) {
    bool is_interpolated;
    V3 w = matvec(o->rot, cp);                                            /* :259, :55-58 */
    w.x = w.x + o->trans[0]; w.y = w.y + o->trans[1]; w.z = w.z + o->trans[2];
    V3 cv = get_voxel_coordinates(o, w);                                  /* :260 */
    if (cv.x < 0 || cv.y < 0 || cv.z < 0) return 2;                       /* :261-264 */
    if (cv.x >= o->m || cv.y >= o->m || cv.z >= o->m) return 2;           /* :265-268 */
    sdf_val = interpolate_distance(o, cv, is_interpolated);               /* :269 */
    if (!is_interpolated) return 3;
    const double v_h = (double)o->v_h;
    const float steps[3] = {o->v_h2_width, o->v_h2_height, o->v_h2_depth};
    for (int a = 0; a < 3; a++) {                                         /* :273-316 */
        V3 pp = cv, pm = cv;
        if (a == 0) { pp.x += v_h; pm.x -= v_h; }
        if (a == 1) { pp.y += v_h; pm.y -= v_h; }
        if (a == 2) { pp.z += v_h; pm.z -= v_h; }
        float plus_h = interpolate_distance(o, pp, is_interpolated);
        if (!is_interpolated) return 3;
        float minus_h = interpolate_distance(o, pm, is_interpolated);
        if (!is_interpolated) return 3;
        J[a] = (plus_h - minus_h) / steps[a];
    }
    const float two_w_h = 2 * (o->w_h);                                   /* :331 (float) */
    for (int a = 0; a < 3; a++) {                                         /* :318-361 */
        V3 wp = matvec(P.r[2 * a], cp);
        wp.x = wp.x + o->trans[0]; wp.y = wp.y + o->trans[1]; wp.z = wp.z + o->trans[2];
        V3 wm = matvec(P.r[2 * a + 1], cp);
        wm.x = wm.x + o->trans[0]; wm.y = wm.y + o->trans[1]; wm.z = wm.z + o->trans[2];
        V3 vp = get_voxel_coordinates(o, wp);
        V3 vm = get_voxel_coordinates(o, wm);
        float plus_h = interpolate_distance(o, vp, is_interpolated);
        if (!is_interpolated) return 3;
        float minus_h = interpolate_distance(o, vm, is_interpolated);
        if (!is_interpolated) return 3;
        J[3 + a] = (plus_h - minus_h) / two_w_h;
    }
    return 1;
}

/* The pixel loop of camera_tracking.cpp:160-184, storing each pixel's record; the sums are
 * then taken sequentially in the reference's loop order (i outer, j inner) in double, which
 * is the canonical order (the reference's own result depends on its OpenMP thread count). */
static int linearize_pixels(Oracle* o, const float* cloud) {
    const int Wd = o->cfg.image_width, Hd = o->cfg.image_height, s = o->cfg.pixel_stride;
    const int ni = (Wd + s - 1) / s, nj = (Hd + s - 1) / s;
    const int n = ni * nj;
    o->pxJ.assign((size_t)n * 6, 0.0f);
    o->pxPsi.assign(n, 0.0f);
    o->pxFlag.assign(n, 0);
    PerturbedRot P;
    build_perturbed(o, P);
#pragma omp parallel for schedule(dynamic, 4)
    for (int ii = 0; ii < ni; ii++) {
        for (int jj = 0; jj < nj; jj++) {
            const int i = ii * s, j = jj * s;
            const float* pt = cloud + 3 * ((size_t)j * Wd + i);           /* at(i = col, j = row) */
            const int p = ii * nj + jj;
            if (std::isnan(pt[0]) || std::isnan(pt[1]) || std::isnan(pt[2])) { o->pxFlag[p] = 0; continue; }
            V3 cp = {(double)pt[0], (double)pt[1], (double)pt[2]};
            float J[6] = {0, 0, 0, 0, 0, 0}, psi = 0;
            int flag = get_partial_derivative(o, P, cp, J, psi);
            o->pxFlag[p] = (uint8_t)flag;
            if (flag == 1) {
                for (int a = 0; a < 6; a++) o->pxJ[(size_t)p * 6 + a] = J[a];
                o->pxPsi[p] = psi;
            }
        }
    }
    return n;
}

static void accumulate(const Oracle* o, int n, double A[36], double b[6], orc_track_stats* st) {
    for (int q = 0; q < 36; q++) A[q] = 0;
    for (int q = 0; q < 6; q++) b[q] = 0;
    int n_valid = 0, n_oob = 0;
    double res = 0;
    for (int p = 0; p < n; p++) {
        if (o->pxFlag[p] == 2) n_oob++;
        if (o->pxFlag[p] != 1) continue;
        n_valid++;
        double J[6];
        for (int a = 0; a < 6; a++) J[a] = (double)o->pxJ[(size_t)p * 6 + a];
        double psi = (double)o->pxPsi[p];
        for (int r = 0; r < 6; r++)
            for (int c = 0; c < 6; c++) A[6 * r + c] = A[6 * r + c] + J[r] * J[c];   /* :181 */
        for (int r = 0; r < 6; r++) b[r] = b[r] + psi * J[r];                          /* :182 */
        res += psi * psi;
    }
    if (st) { st->n_valid = n_valid; st->n_oob = n_oob; st->residual = res; }
}

/* 6x6 partial-pivot LU solve (stands in for Eigen's A.inverse()*b, :191) */
static int solve6(const double Ain[36], const double bin[6], double x[6]) {
    double A[36], b[6];
    memcpy(A, Ain, sizeof A);
    memcpy(b, bin, sizeof b);
    int singular = 0;
    double inv[6] = {0, 0, 0, 0, 0, 0};
    for (int c = 0; c < 6; c++) {
        int piv = c;
        double best = std::fabs(A[6 * c + c]);
        for (int r = c + 1; r < 6; r++) {
            double v = std::fabs(A[6 * r + c]);
            if (v > best) { best = v; piv = r; }
        }
        if (!(best > 0.0)) { singular = 1; continue; }
        if (piv != c) {
            for (int q = 0; q < 6; q++) { double t = A[6 * c + q]; A[6 * c + q] = A[6 * piv + q]; A[6 * piv + q] = t; }
            double t = b[c]; b[c] = b[piv]; b[piv] = t;
        }
        inv[c] = 1.0 / A[6 * c + c];                      /* one reciprocal per pivot, reused by the back substitution */
        for (int r = c + 1; r < 6; r++) {
            double f = A[6 * r + c] * inv[c];
            A[6 * r + c] = f;
            for (int q = c + 1; q < 6; q++) A[6 * r + q] = A[6 * r + q] - f * A[6 * c + q];
            b[r] = b[r] - f * b[c];
        }
    }
    for (int r = 5; r >= 0; r--) {
        double s = b[r];
        for (int q = r + 1; q < 6; q++) s = s - A[6 * r + q] * x[q];
        x[r] = s * inv[r];
    }
    for (int q = 0; q < 6; q++) if (!std::isfinite(x[q])) singular = 1;
    return singular;
}

/* eigen_utils.cpp:40-59 */
static const double ang_min_sinc = 1.0e-8;
static const double ang_min_mc = 2.5e-4;
static double f_sinc(double sinx, double x) { if (std::fabs(x) < ang_min_sinc) return 1.0; else return (sinx / x); }
static double f_mcosc(double cosx, double x) { if (std::fabs(x) < ang_min_mc) return 0.5; else return ((1.0 - cosx) / x / x); }
static double f_msinc(double sinx, double x) { if (std::fabs(x) < ang_min_mc) return (1. / 6.0); else return ((1.0 - sinx / x) / x / x); }

/* eigen_utils.cpp:61-128, delta_t = 1 */
static void direct_exponential_map(const double v[6], double rd[9], double dt[3]) {
    double u[3] = {v[3], v[4], v[5]};
    double theta = std::sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    double si = std::sin(theta), co = std::cos(theta);
    double sinc = f_sinc(si, theta), mcosc = f_mcosc(co, theta), msinc = f_msinc(si, theta);
    rd[0] = co + mcosc * u[0] * u[0];
    rd[1] = -sinc * u[2] + mcosc * u[0] * u[1];
    rd[2] = sinc * u[1] + mcosc * u[0] * u[2];
    rd[3] = sinc * u[2] + mcosc * u[1] * u[0];
    rd[4] = co + mcosc * u[1] * u[1];
    rd[5] = -sinc * u[0] + mcosc * u[1] * u[2];
    rd[6] = -sinc * u[1] + mcosc * u[2] * u[0];
    rd[7] = sinc * u[0] + mcosc * u[2] * u[1];
    rd[8] = co + mcosc * u[2] * u[2];
    dt[0] = v[0] * (sinc + u[0] * u[0] * msinc) + v[1] * (u[0] * u[1] * msinc - u[2] * mcosc) + v[2] * (u[0] * u[2] * msinc + u[1] * mcosc);
    dt[1] = v[0] * (u[0] * u[1] * msinc + u[2] * mcosc) + v[1] * (sinc + u[1] * u[1] * msinc) + v[2] * (u[1] * u[2] * msinc - u[0] * mcosc);
    dt[2] = v[0] * (u[0] * u[2] * msinc - u[1] * mcosc) + v[1] * (u[1] * u[2] * msinc + u[0] * mcosc) + v[2] * (sinc + u[2] * u[2] * msinc);
}

/* camera_tracking.cpp:191-192, 237-239 */
static int apply_update(Oracle* o, const double A[36], const double b[6], double twist[6]) {
    int singular = solve6(A, b, twist);
    if (singular) return 1;                       /* the reference is unguarded (NaN pose); we report */
    double rd[9], td[3];
    direct_exponential_map(twist, rd, td);
    double rdT[9] = {rd[0], rd[3], rd[6], rd[1], rd[4], rd[7], rd[2], rd[5], rd[8]};
    double newrot[9];
    matmul3(rdT, o->rot, newrot);                                         /* :237 */
    V3 tdv = {td[0], td[1], td[2]};
    V3 rt = matvec(rdT, tdv);
    double newt[3] = {o->trans[0] - rt.x, o->trans[1] - rt.y, o->trans[2] - rt.z};   /* :238 */
    set_camera_transformation(o, newrot, newt);                           /* :239 */
    return 0;
}

}  // namespace

extern "C" {

void orc_default_config(orc_config* c) {
    /* sdf_reconstruction.cpp:83-88 */
    c->m = 256; c->width = 6.0f; c->height = 6.0f; c->depth = 3.5f;
    c->origin[0] = -3.0; c->origin[1] = -3.0; c->origin[2] = -0.5;
    c->distance_delta = 0.3f; c->distance_epsilon = 0.025f;
    c->gauss_newton_max_iteration = 20; c->maximum_twist_diff = 0.001f;
    c->v_h = 1.0f; c->w_h = 0.01f; c->pixel_stride = 3; c->metric = 0;
    c->image_width = 640; c->image_height = 480; c->use_coord_table = 1; c->preprocess = 0;
}

void* orc_create(const orc_config* cfg) {
    Oracle* o = new Oracle();
    o->cfg = *cfg;
    o->m = cfg->m;
    o->m_squared = (int64_t)cfg->m * cfg->m;
    o->number_of_voxels = o->m_squared * cfg->m;                          /* sdf.cpp:9 (64-bit here) */
    o->m_div_height = cfg->m / cfg->height;                               /* sdf.cpp:19-21 (fp32) */
    o->m_div_width = cfg->m / cfg->width;
    o->m_div_depth = cfg->m / cfg->depth;
    for (int q = 0; q < 3; q++) o->origin[q] = cfg->origin[q];
    o->v_h = cfg->v_h; o->w_h = cfg->w_h;                                 /* camera_tracking.cpp:11-17 */
    o->v_h2 = 2 * o->v_h; o->w_h2 = 2 * o->w_h;
    o->v_h2_width = o->v_h2 / o->m_div_width;
    o->v_h2_height = o->v_h2 / o->m_div_height;
    o->v_h2_depth = o->v_h2 / o->m_div_depth;
    o->D.resize(o->number_of_voxels);
    o->W.resize(o->number_of_voxels);
    o->isKFilled = false;
    for (int q = 0; q < 9; q++) o->K[q] = 0;
    orc_reset(o);
    if (cfg->use_coord_table) {
        o->global_coords.resize(o->number_of_voxels);                     /* sdf.cpp:11,40-41 */
#pragma omp parallel for
        for (int64_t idx = 0; idx < o->number_of_voxels; idx++) {
            int ijk[3];
            get_voxel_coordinates_idx(o, idx, ijk);
            o->global_coords[idx] = get_global_coordinates(o, ijk);
        }
    }
    /* camera_tracking.cpp:5-8 initial pose */
    double R0[9] = {1, 0, 0, 0, 0, -1, 0, -1, 0};
    double t0[3] = {0, 0, 1};
    set_camera_transformation(o, R0, t0);
    return o;
}
void orc_destroy(void* h) { delete (Oracle*)h; }

void orc_reset(void* h) {
    Oracle* o = (Oracle*)h;
    const float d0 = o->cfg.width + o->cfg.height + o->cfg.depth;         /* sdf.cpp:29 */
#pragma omp parallel for
    for (int64_t i = 0; i < o->number_of_voxels; i++) { o->D[i] = d0; o->W[i] = 0; }
    if (!o->CW.empty()) {
#pragma omp parallel for
        for (int64_t i = 0; i < o->number_of_voxels; i++) { o->CW[i] = 0; o->R[i] = 0.4; o->G[i] = 0.4; o->B[i] = 0.4; }   /* sdf.cpp:30,32-34 */
    }
}

void orc_enable_color(void* h) {
    Oracle* o = (Oracle*)h;
    if (!o->CW.empty()) return;
    o->CW.assign(o->number_of_voxels, 0.0f);                              /* sdf.cpp:14-17, 30-34 */
    o->R.assign(o->number_of_voxels, 0.4);
    o->G.assign(o->number_of_voxels, 0.4);
    o->B.assign(o->number_of_voxels, 0.4);
}
float* orc_color(void* h, int which) {
    Oracle* o = (Oracle*)h;
    if (o->CW.empty()) return nullptr;
    return which == 0 ? o->CW.data() : which == 1 ? o->R.data() : which == 2 ? o->G.data() : o->B.data();
}

/* sdf.cpp:164-217 — colour at WORLD coordinates; out = r,g,b,a floats (std_msgs::ColorRGBA).  Quirks kept:
 * the exact-hit early return leaves the value unscaled (0..255) while the interpolated value is divided
 * by 255; nothing qualifying gives 0/0 = NaN. */
static void interpolate_color(const Oracle* o, const V3& global, float out[4]) {
    V3 vc = get_voxel_coordinates(o, global);                             /* :170 */
    float i = (float)vc.x, j = (float)vc.y, k = (float)vc.z;              /* :172-174 */
    float w_sum = 0.0f, aux = 0;
    float r = 0.0f, g = 0.0f, b = 0.0f;                                   /* color.r/g/b are float32 */
    out[3] = 1.0f;
    float w = 0, volume;
    for (int io = 0; io < 2; io++)
        for (int jo = 0; jo < 2; jo++)
            for (int ko = 0; ko < 2; ko++) {
                const int ci = cast_int(i) + io, cj = cast_int(j) + jo, ck = cast_int(k) + ko;   /* :186-188 */
                volume = std::fabs((float)ci - i) + std::fabs((float)cj - j) + std::fabs((float)ck - k);   /* :189 */
                const int64_t a_idx = get_array_index(o, ci, cj, ck);
                if (a_idx != -1 && o->CW[a_idx] > 0) {
                    if (volume < 0.00001) { out[0] = o->R[a_idx]; out[1] = o->G[a_idx]; out[2] = o->B[a_idx]; return; }   /* :193-198 */
                    w = (float)(1.0 / volume);                            /* :200 */
                    w_sum += w;
                    r += w * o->R[a_idx]; g += w * o->G[a_idx]; b += w * o->B[a_idx];
                }
            }
    aux = (float)(w_sum * 255.0);                                         /* :210 */
    out[0] = r / aux; out[1] = g / aux; out[2] = b / aux;
}
void orc_interpolate_color(void* h, int64_t n, const double* global_pts, float* rgba) {
    Oracle* o = (Oracle*)h;
#pragma omp parallel for
    for (int64_t q = 0; q < n; q++) {
        V3 gp = {global_pts[3 * q], global_pts[3 * q + 1], global_pts[3 * q + 2]};
        interpolate_color(o, gp, rgba + 4 * q);
    }
}

void orc_set_intrinsics(void* h, const double K[9]) {
    Oracle* o = (Oracle*)h;
    for (int q = 0; q < 9; q++) o->K[q] = K[q];
    o->isKFilled = true;
}
void orc_set_pose(void* h, const double R[9], const double t[3]) { set_camera_transformation((Oracle*)h, R, t); }
void orc_get_pose(void* h, double R[9], double t[3]) {
    Oracle* o = (Oracle*)h;
    memcpy(R, o->rot, sizeof o->rot); memcpy(t, o->trans, sizeof o->trans);
}
void orc_get_pose_inv(void* h, double Rinv[9], double tinv[3]) {
    Oracle* o = (Oracle*)h;
    memcpy(Rinv, o->rot_inv, sizeof o->rot_inv); memcpy(tinv, o->rot_inv_trans, sizeof o->rot_inv_trans);
}

void orc_backproject(void* h, const float* depth, float* cloud, float* normals) {
    backproject((Oracle*)h, depth, cloud, normals);
}
void orc_k0_normals(void* h, const float* depth_filtered, float* normals) { k0_normals((Oracle*)h, depth_filtered, normals); }
void orc_preprocess(void* h, const float* depth, float* depth_out, float* normals) {
    Oracle* o = (Oracle*)h;
    k0_bilateral(o, depth, depth_out);
    if (normals) k0_normals(o, depth_out, normals);
}

static void ensure_cloud(Oracle* o, const float* depth, bool want_normals) {
    size_t n = (size_t)o->cfg.image_width * o->cfg.image_height * 3;
    o->cloud.resize(n);
    if (want_normals) o->normals.resize(n);
    if (o->cfg.preprocess) {
        /* sdf_reconstruction.cpp:37-49: filter first, the filtered cloud feeds tracking AND fusion; normals from K0 */
        std::vector<float> zf(n / 3);
        k0_bilateral(o, depth, zf.data());
        backproject(o, zf.data(), o->cloud.data(), nullptr);
        if (want_normals) k0_normals(o, zf.data(), o->normals.data());
        return;
    }
    backproject(o, depth, o->cloud.data(), want_normals ? o->normals.data() : nullptr);
}

int64_t orc_fuse_cloud(void* h, const float* cloud, const float* normals) {
    Oracle* o = (Oracle*)h;
    if (!o->isKFilled) return -1;                                         /* sdf.cpp:227-229 */
    return fuse_cloud(o, cloud, normals);
}
int64_t orc_fuse_rgb(void* h, const float* depth, const uint8_t* rgb) {
    Oracle* o = (Oracle*)h;
    if (!o->isKFilled || o->cfg.metric != 0) return -1;
    orc_enable_color(h);
    ensure_cloud(o, depth, true);
    return fuse_cloud(o, o->cloud.data(), o->normals.data(), rgb);
}
int64_t orc_fuse(void* h, const float* depth) {
    Oracle* o = (Oracle*)h;
    if (!o->isKFilled) return -1;
    ensure_cloud(o, depth, o->cfg.metric == 0);
    return fuse_cloud(o, o->cloud.data(), o->cfg.metric == 0 ? o->normals.data() : nullptr);
}

void orc_linearize(void* h, const float* depth, double A[36], double b[6], orc_track_stats* st) {
    Oracle* o = (Oracle*)h;
    ensure_cloud(o, depth, false);
    int n = linearize_pixels(o, o->cloud.data());
    orc_track_stats tmp;
    memset(&tmp, 0, sizeof tmp);
    accumulate(o, n, A, b, &tmp);
    if (st) *st = tmp;
}

int32_t orc_linearize_pixels(void* h, const float* depth, float* J, float* psi, uint8_t* flag) {
    Oracle* o = (Oracle*)h;
    ensure_cloud(o, depth, false);
    int n = linearize_pixels(o, o->cloud.data());
    if (J) memcpy(J, o->pxJ.data(), sizeof(float) * 6 * n);
    if (psi) memcpy(psi, o->pxPsi.data(), sizeof(float) * n);
    if (flag) memcpy(flag, o->pxFlag.data(), n);
    return n;
}

int32_t orc_apply_update(void* h, const double A[36], const double b[6], double twist_out[6]) {
    return apply_update((Oracle*)h, A, b, twist_out);
}

/* camera_tracking.cpp:66-245 */
void orc_track(void* h, const float* depth, orc_track_stats* st) {
    Oracle* o = (Oracle*)h;
    orc_track_stats s;
    memset(&s, 0, sizeof s);
    ensure_cloud(o, depth, false);
    bool stop = false;
    const double maximum_twist_diff = (double)o->cfg.maximum_twist_diff;  /* float member promoted */
    for (int g = 0; g < o->cfg.gauss_newton_max_iteration && !stop; g++) {  /* :79 */
        double A[36], b[6], twist[6];
        int n = linearize_pixels(o, o->cloud.data());
        accumulate(o, n, A, b, &s);
        s.iterations = g + 1;
        if (apply_update(o, A, b, twist)) { s.singular = 1; break; }
        /* :216-224 signed test (no fabs); this iteration's update was still applied */
        if (twist[0] < maximum_twist_diff && twist[1] < maximum_twist_diff && twist[2] < maximum_twist_diff &&
            twist[3] < maximum_twist_diff && twist[4] < maximum_twist_diff && twist[5] < maximum_twist_diff) {
            stop = true;
            s.stopped = 1;
        }
    }
    if (st) *st = s;
}

void orc_interpolate(void* h, int64_t n, const double* pts, float* out, uint8_t* ok) {
    Oracle* o = (Oracle*)h;
#pragma omp parallel for
    for (int64_t q = 0; q < n; q++) {
        V3 v = {pts[3 * q], pts[3 * q + 1], pts[3 * q + 2]};
        bool is_interp;
        out[q] = interpolate_distance(o, v, is_interp);
        ok[q] = is_interp ? 1 : 0;
    }
}

/* ---- pcl::MarchingCubesSDF (marching_cubes_sdf.cpp), the mesher of SDF::visualize (sdf.cpp:327) -------------
 * The triangle table is the classic public-domain one (Bourke 1994 / Bloyd) that PCL ships and the reference
 * vendors (marching_cubes_sdf.h:107-364); it is kept packed in tracking_sdf_b200/csrc/mc_tables.inc (generated
 * by tools/gen_mc_tables.py, which also proves it consistent with the cube topology) and unpacked here.  The
 * edge mask (marching_cubes_sdf.h:73-106) is derived from the topology. */
static const unsigned long long mc_packed[256] = {
#include "../tracking_sdf_b200/csrc/mc_tables.inc"
};
static int mc_triTable[256][16];
static unsigned int mc_edgeTable[256];
static bool mc_ready = false;
static void mc_init_tables() {
    if (mc_ready) return;
    static const int ea[12] = {0, 1, 2, 3, 4, 5, 6, 7, 0, 1, 2, 3}, eb[12] = {1, 2, 3, 0, 5, 6, 7, 4, 4, 5, 6, 7};
    for (int c = 0; c < 256; c++) {
        for (int q = 0; q < 16; q++) {
            const int v = (int)((mc_packed[c] >> (4 * q)) & 0xF);
            mc_triTable[c][q] = (v == 0xF) ? -1 : v;
        }
        unsigned int mask = 0;
        for (int e = 0; e < 12; e++)
            if (((c >> ea[e]) & 1) != ((c >> eb[e]) & 1)) mask |= 1u << e;
        mc_edgeTable[c] = mask;
    }
    mc_ready = true;
}
struct V3f { float v[3]; };
/* marching_cubes_sdf.cpp:87-99 */
static void mc_interpolateEdge(const V3f& p1, const V3f& p2, float val_p1, float val_p2, float iso_level, V3f& output) {
    float mu = (iso_level - val_p1) / (val_p2 - val_p1);
    for (int c = 0; c < 3; c++) output.v[c] = p1.v[c] + mu * (p2.v[c] - p1.v[c]);
}
/* marching_cubes_sdf.cpp:101-197 */
static void mc_createSurface(const Oracle* o, float iso_level, const float leaf_node[8], const int index_3d[3], std::vector<float>& cloud) {
    int cubeindex = 0;
    V3f vertex_list[12];
    for (int n = 0; n < 8; n++)
        if (leaf_node[n] < iso_level) cubeindex |= 1 << n;                /* :107-115 */
    if (mc_edgeTable[cubeindex] == 0) return;                             /* :118 */
    const float min_p[3] = {0, 0, 0}, max_p[3] = {o->cfg.width, o->cfg.height, o->cfg.depth};   /* setBBox :52-63 */
    const float res[3] = {float(o->m), float(o->m), float(o->m)};
    V3f center;
    for (int c = 0; c < 3; c++) center.v[c] = min_p[c] + (max_p[c] - min_p[c]) * float(index_3d[c]) / res[c];   /* :123-125 */
    V3f p[8];
    for (int i = 0; i < 8; i++) {                                         /* :127-140 */
        V3f point = center;
        if (i & 0x4) point.v[1] = static_cast<float>(center.v[1] + (max_p[1] - min_p[1]) / res[1]);
        if (i & 0x2) point.v[2] = static_cast<float>(center.v[2] + (max_p[2] - min_p[2]) / res[2]);
        if ((i & 0x1) ^ ((i >> 1) & 0x1)) point.v[0] = static_cast<float>(center.v[0] + (max_p[0] - min_p[0]) / res[0]);
        p[i] = point;
    }
    static const int ea[12] = {0, 1, 2, 3, 4, 5, 6, 7, 0, 1, 2, 3}, eb[12] = {1, 2, 3, 0, 5, 6, 7, 4, 4, 5, 6, 7};   /* :146-169 */
    for (int e = 0; e < 12; e++)
        if (mc_edgeTable[cubeindex] & (1u << e)) mc_interpolateEdge(p[ea[e]], p[eb[e]], leaf_node[ea[e]], leaf_node[eb[e]], iso_level, vertex_list[e]);
    for (int i = 0; mc_triTable[cubeindex][i] != -1; i += 3)              /* :175-194 */
        for (int q = 0; q < 3; q++) {
            const V3f& v = vertex_list[mc_triTable[cubeindex][i + q]];
            cloud.push_back(v.v[0]); cloud.push_back(v.v[1]); cloud.push_back(v.v[2]);
        }
}
/* marching_cubes_sdf.cpp:203-241 */
static void mc_getNeighborList1D(const Oracle* o, float leaf[8], const int pos[3]) {
    const int64_t res_z = o->m, zy_index_offset = (int64_t)o->m * o->m;
    int64_t g0 = pos[0] * zy_index_offset + pos[1] * res_z + pos[2];
    int64_t g1 = g0 + zy_index_offset, g2 = g1 + 1, g3 = g0 + 1, g4 = g0 + res_z, g5 = g4 + zy_index_offset, g6 = g5 + 1, g7 = g4 + 1;
    const int64_t gi[8] = {g0, g1, g2, g3, g4, g5, g6, g7};
    bool all = true;
    for (int n = 0; n < 8; n++) all = all && (o->W[gi[n]] > 0);
    for (int n = 0; n < 8; n++) leaf[n] = all ? o->D[gi[n]] : o->D[g0];
}
/* marching_cubes_sdf.cpp:243-287 over the interior voxels (sdf.cpp:36-39), serial = the reference's concatenation
 * of per-thread clouds in index order.  Returns the number of vertices (3 per triangle). */
int64_t orc_mesh(void* h, float iso_level) {
    Oracle* o = (Oracle*)h;
    mc_init_tables();
    o->mesh.clear();
    if (!(iso_level >= 0 && iso_level < 1)) return 0;                     /* :248-254 */
    for (int i = 1; i < o->m - 1; i++)
        for (int j = 1; j < o->m - 1; j++)
            for (int k = 1; k < o->m - 1; k++) {
                float leaf_node[8];
                const int index_3d[3] = {i, j, k};
                mc_getNeighborList1D(o, leaf_node, index_3d);
                mc_createSurface(o, iso_level, leaf_node, index_3d, o->mesh);
            }
    return (int64_t)(o->mesh.size() / 3);
}
/* the last mesh: xyz (n*3 floats, mesher frame), and optionally the marker points / colours of
 * SDF::visualize (sdf.cpp:354-356, 380-385): world = (double)xyz + sdf_origin, rgba = interpolate_color(world) */
void orc_mesh_copy(void* h, float* xyz, double* world, float* rgba) {
    Oracle* o = (Oracle*)h;
    const int64_t n = (int64_t)(o->mesh.size() / 3);
    if (xyz) memcpy(xyz, o->mesh.data(), o->mesh.size() * sizeof(float));
    for (int64_t q = 0; q < n; q++) {
        V3 p = {o->mesh[3 * q] + o->origin[0], o->mesh[3 * q + 1] + o->origin[1], o->mesh[3 * q + 2] + o->origin[2]};
        if (world) { world[3 * q] = p.x; world[3 * q + 1] = p.y; world[3 * q + 2] = p.z; }
        if (rgba) interpolate_color(o, p, rgba + 4 * q);
    }
}

void orc_exp_map(const double twist[6], double R[9], double t[3]) { direct_exponential_map(twist, R, t); }

int64_t orc_get_array_index(void* h, int32_t i, int32_t j, int32_t k) { return get_array_index((Oracle*)h, i, j, k); }
void orc_get_voxel_coordinates_idx(void* h, int64_t idx, int32_t ijk[3]) { get_voxel_coordinates_idx((Oracle*)h, idx, ijk); }
void orc_get_voxel_coordinates(void* h, const double g[3], double v[3]) {
    V3 gg = {g[0], g[1], g[2]};
    V3 r = get_voxel_coordinates((Oracle*)h, gg);
    v[0] = r.x; v[1] = r.y; v[2] = r.z;
}
void orc_get_global_coordinates(void* h, const int32_t ijk[3], double g[3]) {
    int q[3] = {ijk[0], ijk[1], ijk[2]};
    V3 r = get_global_coordinates((Oracle*)h, q);
    g[0] = r.x; g[1] = r.y; g[2] = r.z;
}

float* orc_D(void* h) { return ((Oracle*)h)->D.data(); }
float* orc_W(void* h) { return ((Oracle*)h)->W.data(); }
int64_t orc_number_of_voxels(void* h) { return ((Oracle*)h)->number_of_voxels; }

/* sdf.cpp:99-126 (D = dist - radius, W = 1; colour writes dropped) */
void orc_create_circle(void* h, float radius, float center_x, float center_y, float center_z) {
    Oracle* o = (Oracle*)h;
#pragma omp parallel for
    for (int64_t idx = 0; idx < o->number_of_voxels; idx++) {
        int ijk[3];
        get_voxel_coordinates_idx(o, idx, ijk);
        V3 g = get_global_coordinates(o, ijk);
        double x = g.x, y = g.y, z = g.z;
        double d = std::sqrt((x - center_x) * (x - center_x) + (y - center_y) * (y - center_y) + (z - center_z) * (z - center_z));
        o->D[idx] = d - radius;
        o->W[idx] = 1.0;
    }
}

void orc_get_constants(void* h, float out[6]) {
    Oracle* o = (Oracle*)h;
    out[0] = o->m_div_width; out[1] = o->m_div_height; out[2] = o->m_div_depth;
    out[3] = o->v_h2_width; out[4] = o->v_h2_height; out[5] = o->v_h2_depth;
}

int orc_num_threads(void) { return omp_get_max_threads(); }
void orc_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }

}  // extern "C"
