/*
 * tsdf_core.cuh — exact-arithmetic building blocks of the track + fuse kernels.
 *
 * Everything here is `__host__ __device__` and written as plain IEEE arithmetic: the
 * library is compiled with nvcc -fmad=false (and default -prec-div/-prec-sqrt, no
 * fast-math) so each + - * / rounds exactly as the reference's non-contracted x86 code does
 * (SURVEY.md "hard part 1").  The host compile of this header (tests/host_emul, g++
 * -ffp-contract=off) exists only so the CPU test-suite can check the kernels' per-voxel /
 * per-sample logic against the oracle without a GPU; the product never runs it.
 *
 * Reference citations are relative to /root/reference/src/.
 */
#pragma once
#include <stdint.h>
#include <limits.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define TSDF_HD __host__ __device__ __forceinline__
#else
#define TSDF_HD inline
#endif

namespace tsdf {

/* --- parameters shared by all kernels (built on the host in fp32/fp64 exactly like the
 * reference's constructors, sdf.cpp:18-21, camera_tracking.cpp:11-17) ------------------- */
struct GridParams {
    int32_t m;                 /* voxels per axis */
    int32_t ks0, ks1;          /* stored z range [ks0, ks1)  (slab + halo)  */
    int32_t ko0, ko1;          /* owned  z range [ko0, ko1)                  */
    int32_t metric;            /* 0 plane, 1 point */
    int32_t img_w, img_h;
    int32_t stride;            /* pixel stride of the tracker */
    int32_t ni, nj;            /* strided pixel grid (columns, rows) */
    float m_div_width, m_div_height, m_div_depth;      /* sdf.cpp:19-21 (fp32) */
    double m_div_d[3];         /* the same three fp32 values widened once on the host: a constant-bank operand instead of a
                                * float -> double conversion per use */
    float vs_x, vs_y, vs_z;    /* extent / (float)m, the fp32 quotient of sdf.h:154-156 */
    float delta, eps;          /* sdf.cpp:8 */
    float v_h, w_h;            /* camera_tracking.cpp:11-12 */
    float v_h2_width, v_h2_height, v_h2_depth;         /* camera_tracking.cpp:15-17 */
    float two_w_h;             /* 2 * w_h as float, camera_tracking.cpp:331 */
    float max_twist_diff;
    int32_t max_iter;
    double origin[3];
    double K[9];
    int32_t k_simple;          /* 1 when K = [fx 0 cx; 0 fy cy; 0 0 1]: the zero products are exact no-ops */
    int32_t pad;
};

/* pose block living in device memory; written by set_pose (host) and by the tracker's
 * on-device update (camera_tracking.cpp:237-239) */
struct PoseState {
    double R[9], t[3];         /* rot, trans            camera_tracking.h:42-47 */
    double Rinv[9], tinv[3];   /* rot_inv, rot_inv_trans */
    double twist[6];
    double sums[30];           /* last reduced normal equations: 21 A (upper, row-major), 6 b, psi^2, n_valid, n_oob */
    int32_t iterations, stopped, singular, halo_miss;
};

struct PixRec { float z, nx, ny, nz; };     /* per-pixel record written by K1 (16 B) */

/* ---- bit access to doubles (device: register moves; host: memcpy) */
TSDF_HD int dbl_hi(double x) {
#if defined(__CUDA_ARCH__)
    return __double2hiint(x);
#else
    long long b; memcpy(&b, &x, 8); return (int)(b >> 32);
#endif
}
TSDF_HD int dbl_lo(double x) {
#if defined(__CUDA_ARCH__)
    return __double2loint(x);
#else
    long long b; memcpy(&b, &x, 8); return (int)(b & 0xffffffffll);
#endif
}
/* correctly rounded fp32 reciprocal: identical bits to 1.0f / x, fewer instructions on the device */
TSDF_HD float rcp_rn(float x) {
#if defined(__CUDA_ARCH__)
    return __frcp_rn(x);
#else
    return 1.0f / x;
#endif
}
/* The same for operands known to be normal and in [2^-17, 4] (the L1 "volume" of sdf.cpp:146
 * lies in (1e-5, 3]): hardware reciprocal + one FMA Newton step, without __frcp_rn's range
 * checks.  Verified EXHAUSTIVELY against 1.0f/x over every float of that range on the device
 * (tests/test_gpu_parity.py::test_fast_reciprocal_is_exact_on_its_range). */
TSDF_HD float rcp_rn_small(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    const float e = fmaf(-x, r, 1.0f);
    return fmaf(r, e, r);
#else
    return 1.0f / x;
#endif
}

/* ---- the reference's (int) casts: x86 cvttss2si — truncation toward zero, NaN and
 * out-of-range give INT_MIN.  sdf.cpp:143-145 ---- */
TSDF_HD int trunc_f2i(float v) {
    /* |v| < 2^31, else INT_MIN (v = -2^31 converts to INT_MIN either way; NaN fails the compare): one compare + select */
    return (fabsf(v) < 2147483648.0f) ? (int)v : INT_MIN;
}

/* `volume < 0.00001` compares the float promoted to double against a double literal
 * (sdf.cpp:151).  float(1e-5) = 9.99999974737875e-06 < 1e-5, so the test is `volume <=`
 * that float. */
#define TSDF_VOL_EXACT_F 9.99999974737875163555145263671875e-06f

/* sdf.h:143-147 */
TSDF_HD void world_to_voxel(const GridParams& g, double wx, double wy, double wz,
                            double& vx, double& vy, double& vz) {
    vx = ((wx - g.origin[0]) * g.m_div_d[0] - 0.5);
    vy = ((wy - g.origin[1]) * g.m_div_d[1] - 0.5);
    vz = ((wz - g.origin[2]) * g.m_div_d[2] - 0.5);
}
/* sdf.h:153-157, one axis */
TSDF_HD double voxel_centre(float vs, int idx, double org) {
    return (double)vs * ((double)idx + 0.5) + org;
}

/* Eigen fixed-size product coefficient: ((a0*b0 + a1*b1) + a2*b2) */
TSDF_HD double dot3_seq(double a0, double a1, double a2, double b0, double b1, double b2) {
    return (a0 * b0 + a1 * b1) + a2 * b2;
}
TSDF_HD void matvec3(const double* M, double x, double y, double z, double& ox, double& oy, double& oz) {
    ox = dot3_seq(M[0], M[1], M[2], x, y, z);
    oy = dot3_seq(M[3], M[4], M[5], x, y, z);
    oz = dot3_seq(M[6], M[7], M[8], x, y, z);
}
TSDF_HD void matmul3(const double* A, const double* B, double* C) {
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++)
            C[3 * r + c] = (A[3 * r + 0] * B[0 + c] + A[3 * r + 1] * B[3 + c]) + A[3 * r + 2] * B[6 + c];
}
/* Matrix3d::inverse(): cofactors * (1/det)  (camera_tracking.cpp:62) */
TSDF_HD void inverse3(const double* M, double* inv) {
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
/* camera_tracking.cpp:59-65 */
TSDF_HD void pose_set(PoseState& p, const double* R, const double* t) {
#pragma unroll
    for (int q = 0; q < 9; q++) p.R[q] = R[q];
#pragma unroll
    for (int q = 0; q < 3; q++) p.t[q] = t[q];
    inverse3(p.R, p.Rinv);
    double x, y, z;
    matvec3(p.Rinv, p.t[0], p.t[1], p.t[2], x, y, z);
    p.tinv[0] = -1 * x; p.tinv[1] = -1 * y; p.tinv[2] = -1 * z;
}

/* K1 back-projection (depth_image_proc convention, fp32): x = (u - cx) * z * (1/fx) */
struct K1Params { float cx, cy, inv_fx, inv_fy; };
TSDF_HD K1Params k1_params(const double* K) {
    K1Params p;
    p.cx = (float)K[2]; p.cy = (float)K[5];
    p.inv_fx = 1.0f / (float)K[0]; p.inv_fy = 1.0f / (float)K[4];
    return p;
}
TSDF_HD bool depth_valid(float z) { return (z > 0.0f) && (z <= 3.402823466e+38f); }   /* finite, > 0; false for NaN */
TSDF_HD void backproject_px(const K1Params& p, int u, int v, float z, float& x, float& y) {
    x = ((float)u - p.cx) * z * p.inv_fx;
    y = ((float)v - p.cy) * z * p.inv_fy;
}
/* normal from central differences of the 4-neighbourhood; false = not computable (NaN) */
TSDF_HD bool normal_px(const K1Params& p, int u, int v, float zc, float zl, float zr, float zu, float zd,
                       float& nx, float& ny, float& nz) {
    if (!(depth_valid(zc) && depth_valid(zl) && depth_valid(zr) && depth_valid(zu) && depth_valid(zd))) return false;
    float thr = 0.02f * zc;        /* PCL MaxDepthChangeFactor, sdf_reconstruction.cpp:46 */
    if (fabsf(zl - zc) > thr || fabsf(zr - zc) > thr || fabsf(zu - zc) > thr || fabsf(zd - zc) > thr) return false;
    float xc, yc, xl, yl, xr, yr, xu, yu, xd, yd;
    backproject_px(p, u, v, zc, xc, yc);
    backproject_px(p, u - 1, v, zl, xl, yl);
    backproject_px(p, u + 1, v, zr, xr, yr);
    backproject_px(p, u, v - 1, zu, xu, yu);
    backproject_px(p, u, v + 1, zd, xd, yd);
    float ax = xr - xl, ay = yr - yl, az = zr - zl;
    float bx = xd - xu, by = yd - yu, bz = zd - zu;
    float cx_ = ay * bz - az * by;
    float cy_ = az * bx - ax * bz;
    float cz_ = ax * by - ay * bx;
    float len = sqrtf((cx_ * cx_ + cy_ * cy_) + cz_ * cz_);
    if (!(len > 0.0f)) return false;
    cx_ = cx_ / len; cy_ = cy_ / len; cz_ = cz_ / len;
    float dotp = (cx_ * xc + cy_ * yc) + cz_ * zc;
    if (dotp > 0.0f) { cx_ = -cx_; cy_ = -cy_; cz_ = -cz_; }
    nx = cx_; ny = cy_; nz = cz_;
    return true;
}

/* ---- SDF::interpolate_distance, sdf.cpp:127-163 -----------------------------------------
 * The Fetch object gives access to the voxel store:
 *   fetch(ci,cj,ck,D,W) -> true when the voxel is inside the grid (sdf.h:113-119)
 *   fetch.interior(bi,bj,bk) -> all 8 cells bi..bi+1 x bj..bj+1 x bk..bk+1 are inside (and held)
 *   fetch.load8(bi,bj,bk,d,w) -> their D,W in the reference's neighbour order (i outer, j, k inner)
 * Semantics kept exactly: neighbour order, fp32 L1 "volume" = (|di|+|dj|)+|dk|, w = 1/volume,
 * sequential fp32 sums, the early return when a W>0 neighbour has volume < 1e-5, and
 * is_interpolated = any neighbour with W>0.  The six |c - x| terms are computed once (they are
 * the same fp32 values the reference recomputes per neighbour).  The interior case first issues
 * all eight loads, then does the arithmetic; the early-return test is hoisted: it can only
 * fire when one candidate distance per axis is <= 1e-5. */
#ifndef RCP_VOLUME
#define RCP_VOLUME rcp_rn_small
#endif
template <bool CHECK_EXACT>
TSDF_HD float interp_accumulate(const float* d, const float* w, const bool* inb,
                                const float* fx, const float* fy, const float* fz, bool& is_interpolated) {
    float w_sum = 0.0f, sum_d = 0.0f;
    bool any = false, exact = false;
    float exact_val = 0.0f;
    if (!CHECK_EXACT) {
        /* no neighbour can take the early return: straight-line, select instead of branch.
         * Skipping a neighbour leaves both sums untouched, exactly like the reference's `if`. */
#pragma unroll
        for (int n = 0; n < 8; n++) {
            const float volume = (fx[n >> 2] + fy[(n >> 1) & 1]) + fz[n & 1];
            const bool use = inb[n] & (w[n] > 0.0f);
            const float wt = RCP_VOLUME(volume);  /* == (float)(1.0 / (double)volume), sdf.cpp:154 */
            const float ws = w_sum + wt, sd = sum_d + wt * d[n];
            w_sum = use ? ws : w_sum;
            sum_d = use ? sd : sum_d;
            any = any | use;
        }
        is_interpolated = any;
        return sum_d / w_sum;
    }
#pragma unroll
    for (int n = 0; n < 8; n++) {
        const float volume = (fx[n >> 2] + fy[(n >> 1) & 1]) + fz[n & 1];
        const bool use = inb[n] && w[n] > 0.0f && !exact;
        if (use) {
            any = true;
            if (volume <= TSDF_VOL_EXACT_F) {
                exact = true;
                exact_val = d[n];
            } else {
                const float wt = rcp_rn(volume);      /* == (float)(1.0 / (double)volume), sdf.cpp:154 */
                w_sum = w_sum + wt;
                sum_d = sum_d + wt * d[n];
            }
        }
    }
    is_interpolated = any;
    return exact ? exact_val : sum_d / w_sum;
}

template <class Fetch>
TSDF_HD float interpolate_distance(double vx, double vy, double vz, Fetch&& fetch, bool& is_interpolated) {
    const float i = (float)vx, j = (float)vy, k = (float)vz;
    const int bi = trunc_f2i(i), bj = trunc_f2i(j), bk = trunc_f2i(k);
    float fx[2], fy[2], fz[2];
    fx[0] = fabsf((float)bi - i); fx[1] = fabsf((float)(bi + 1) - i);
    fy[0] = fabsf((float)bj - j); fy[1] = fabsf((float)(bj + 1) - j);
    fz[0] = fabsf((float)bk - k); fz[1] = fabsf((float)(bk + 1) - k);
    float d[8], w[8];
    bool inb[8];
    if (fetch.interior(bi, bj, bk)) {
        fetch.load8(bi, bj, bk, d, w);
#pragma unroll
        for (int n = 0; n < 8; n++) inb[n] = true;
        const bool maybe_exact = (fminf(fx[0], fx[1]) <= TSDF_VOL_EXACT_F) && (fminf(fy[0], fy[1]) <= TSDF_VOL_EXACT_F) &&
                                 (fminf(fz[0], fz[1]) <= TSDF_VOL_EXACT_F);
        if (!maybe_exact) return interp_accumulate<false>(d, w, inb, fx, fy, fz, is_interpolated);
        return interp_accumulate<true>(d, w, inb, fx, fy, fz, is_interpolated);
    }
#pragma unroll
    for (int n = 0; n < 8; n++) {
        d[n] = 0.0f; w[n] = 0.0f;
        inb[n] = fetch(bi + (n >> 2), bj + ((n >> 1) & 1), bk + (n & 1), d[n], w[n]);
    }
    return interp_accumulate<true>(d, w, inb, fx, fy, fz, is_interpolated);
}

/* ---- fusion, per voxel: sdf.cpp:245-292 ---------------------------------------------------
 * Stage 1 (projection): camera-space voxel centre -> pixel, or reject. */
/* reciprocal good to ~2e-14 relative for y in the float range: device = fp32 MUFU.RCP of the
 * rounded operand (2^-23) + one Newton step in double (squares it); y beyond the float range
 * gives 0/inf and the caller's exponent guard sends the voxel to the exact path.  host = 1/y */
TSDF_HD double rcp_fast(double y) {
#if defined(__CUDA_ARCH__)
    float rf;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rf) : "f"(__double2float_rn(y)));
    const double r = (double)rf;
    const double e = fma(-y, r, 1.0);
    return fma(r, e, r);
#else
    return 1.0 / y;
#endif
}
/* the reference's arithmetic, verbatim: two IEEE double divisions and (int) truncation */
TSDF_HD bool project_exact(const GridParams& g, double ij0, double ij1, double ij2, int& iu, int& iv) {
    const double u = ij0 / ij2, v = ij1 / ij2;                  /* camera_tracking.cpp:45-46 */
    /* (int) truncation then 0 <= i < width  <=>  -1 < u < width (false for NaN/inf);  sdf.cpp:251-254 */
    if (!(u > -1.0 && u < (double)g.img_w && v > -1.0 && v < (double)g.img_h)) return false;
    iu = (int)u; iv = (int)v;
    return true;
}
TSDF_HD int imin(int a, int b) { return a < b ? a : b; }
TSDF_HD int imax(int a, int b) { return a > b ? a : b; }
/* camera-space centre -> (ij0, ij1, ij2) = K * cam  (camera_tracking.cpp:44) */
TSDF_HD void project_ij(const GridParams& g, double cx, double cy, double cz, double& ij0, double& ij1, double& ij2) {
    if (g.k_simple) {       /* 0*x and +0 are exact no-ops for finite operands */
        ij0 = g.K[0] * cx + g.K[2] * cz;
        ij1 = g.K[4] * cy + g.K[5] * cz;
        ij2 = cz;
    } else {
        ij0 = dot3_seq(g.K[0], g.K[1], g.K[2], cx, cy, cz);
        ij1 = dot3_seq(g.K[3], g.K[4], g.K[5], cx, cy, cz);
        ij2 = dot3_seq(g.K[6], g.K[7], g.K[8], cx, cy, cz);
    }
}
/* Branch-free projection.  Only trunc(fl(ij0/ij2)) and trunc(fl(ij1/ij2)) are consumed
 * (sdf.cpp:251-252).  With an approximate quotient q (relative error <= ~1e-12; fl() adds
 * 1.1e-16) the truncated value is certain unless q lies within TOL of an integer; those, and
 * |q| >= 2^30 or non-finite q (voxel practically in the camera plane) set need_exact and the
 * caller runs project_exact: results are identical to the reference's two IEEE divisions.
 * Round-to-nearest via the 2^52+2^51 constant: the low word of q + MAGIC is the integer,
 * d = q - rint(q) is exact, floor(q) = rint(q) - (d < 0).
 * iu, iv always come back clamped into the image so the pixel fetch is unconditional. */
TSDF_HD void fuse_project_flags(const GridParams& g, double cx, double cy, double cz,
                                int& iu, int& iv, bool& ok, bool& need_exact) {
    double ij0, ij1, ij2;
    project_ij(g, cx, cy, cz, ij0, ij1, ij2);
    const double TOL = 1e-7, MAGIC = 6755399441055744.0;
    const double r = rcp_fast(ij2);
    const double qu = ij0 * r, qv = ij1 * r;
    const int eu = dbl_hi(qu) & 0x7ff00000, ev = dbl_hi(qv) & 0x7ff00000;
    const double tu = qu + MAGIC, tv = qv + MAGIC;
    const double du = qu - (tu - MAGIC), dv = qv - (tv - MAGIC);
    const bool safe = (eu < 0x41d00000) & (ev < 0x41d00000) & (fabs(du) > TOL) & (fabs(dv) > TOL);
    const int fu = dbl_lo(tu) + (dbl_hi(du) >> 31), fv = dbl_lo(tv) + (dbl_hi(dv) >> 31);
    /* floor in [-1, W-1] <=> truncation toward zero in [0, W-1]; (-1,0) maps to 0 */
    const bool inimg = ((unsigned)(fu + 1) <= (unsigned)g.img_w) & ((unsigned)(fv + 1) <= (unsigned)g.img_h);
    const bool zpos = !(cz < 0);                                /* sdf.cpp:247 */
    ok = zpos & safe & inimg;
    need_exact = zpos & !safe;
    iu = imin(imax(fu, 0), g.img_w - 1);
    iv = imin(imax(fv, 0), g.img_h - 1);
}
TSDF_HD bool fuse_project(const GridParams& g, double cx, double cy, double cz, int& iu, int& iv) {
    bool ok, need_exact;
    fuse_project_flags(g, cx, cy, cz, iu, iv, ok, need_exact);
    if (need_exact) {
        double ij0, ij1, ij2;
        project_ij(g, cx, cy, cz, ij0, ij1, ij2);
        return project_exact(g, ij0, ij1, ij2, iu, iv);
    }
    return ok;
}

/* exp(x) for the weight of sdf.cpp:278.  For x in [-0.04, 0] a degree-9 Taylor polynomial in
 * double is accurate to < 2e-16 relative, i.e. it rounds to the same float as a correctly
 * rounded exp except with probability ~1e-8 per call; outside that range call exp(). */
TSDF_HD double weight_exp(double x) {
    if (x >= -0.04) {
        double p = 1.0 / 362880.0;
        p = fma(p, x, 1.0 / 40320.0);
        p = fma(p, x, 1.0 / 5040.0);
        p = fma(p, x, 1.0 / 720.0);
        p = fma(p, x, 1.0 / 120.0);
        p = fma(p, x, 1.0 / 24.0);
        p = fma(p, x, 1.0 / 6.0);
        p = fma(p, x, 0.5);
        p = fma(p, x, 1.0);
        p = fma(p, x, 1.0);
        return p;
    }
    return exp(x);
}

/* Stage 2 (distance), branch-free: returns false when the voxel is skipped.  rec = the pixel's
 * {z,n}; px,py = the pixel's back-projected x,y (backproject_px, bit-identical to K1).  d_out is
 * already truncated (Eq. 28); band says the exponential weight applies, with e_band = d - eps. */
TSDF_HD bool fuse_distance_flags(const GridParams& g, double cx, double cy, double cz,
                                 float px, float py, const PixRec& rec, float& d_out, float& e_band, bool& band) {
    float d_new;
    bool valid;
    if (g.metric == 0) {
        /* sdf.cpp:260: isnan(point.x)|isnan(point.y)|isnan(normal.*) — z NaN makes x,y NaN */
        valid = (rec.z == rec.z) & (rec.nx == rec.nx);
        /* sdf.h:177-181, Eigen dot = c0 + (c1 + c2) */
        const double dx = (double)px - cx, dy = (double)py - cy, dz = (double)rec.z - cz;
        const double pp = dx * (double)rec.nx + (dy * (double)rec.ny + dz * (double)rec.nz);
        d_new = (float)pp;                                      /* sdf.cpp:274 */
    } else {
        valid = (rec.z == rec.z);
        const double pp = cz - (double)rec.z;                   /* sdf.h:169-172 */
        d_new = (float)pp;
    }
    band = (d_new >= g.eps) & (d_new <= g.delta);               /* sdf.cpp:277 */
    e_band = d_new - g.eps;
    valid = valid & !(d_new > g.delta);                         /* sdf.cpp:280-283 */
    d_out = (d_new < -g.delta) ? -g.delta : d_new;              /* sdf.cpp:285-287 */
    return valid;
}
/* weight of sdf.cpp:276-279 */
TSDF_HD float fuse_weight(bool band, float e_band) {
    return band ? (float)weight_exp(-0.5 * (double)e_band * (double)e_band) : 1.0f;
}
TSDF_HD bool fuse_distance(const GridParams& g, double cx, double cy, double cz,
                           float px, float py, const PixRec& rec, float& d_out, float& w_out) {
    float e; bool band;
    const bool valid = fuse_distance_flags(g, cx, cy, cz, px, py, rec, d_out, e, band);
    if (!valid) return false;
    w_out = fuse_weight(band, e);
    return true;
}
/* Stage 3 (running weighted mean): sdf.cpp:289-292 */
TSDF_HD void fuse_apply(float& D, float& W, float d_new, float w_new) {
    const float w_old = W;
    W = w_old + w_new;
    D = (w_old * D + w_new * d_new) / W;
}
/* same, predicated without a branch (w_new must be > 0 even when !upd so the quotient is benign) */
TSDF_HD void fuse_apply_sel(float& D, float& W, float d_new, float w_new, bool upd) {
    const float w_old = W;
    const float Wn = w_old + w_new;
    const float Dn = (w_old * D + w_new * d_new) / Wn;
    W = upd ? Wn : w_old;
    D = upd ? Dn : D;
}

/* ---- colour fusion, sdf.cpp:294-304 --------------------------------------------------------------
 * cosine = |cam_vect . n| / ||n|| with cam_vect = (0,0,1) (sdf.cpp:235): the dot product is n_z
 * exactly (the two zero products only add a signed zero); Eigen's norm is sqrt(c0 + (c1 + c2)).
 * The colour weight is the D/W weight times the cosine: float = float * double (:299). */
TSDF_HD double color_cosine(float nx, float ny, float nz) {           /* per pixel: K1 tabulates it */
    const double x = (double)nx, y = (double)ny, z = (double)nz;
    const double norm = sqrt(x * x + (y * y + z * z));
    return fabs(z) / norm;
}
TSDF_HD float color_weight(float w_new, double cosine) { return (float)((double)w_new * cosine); }
/* running mean of one voxel's colour: the uint8 channel is promoted to int, then to float (:302-304) */
TSDF_HD void color_apply(float& CW, float& R, float& G, float& B, float w_c, int r, int g, int b) {
    const float w_old = CW;
    CW = w_old + w_c;
    R = (w_old * R + w_c * (float)r) / CW;
    G = (w_old * G + w_c * (float)g) / CW;
    B = (w_old * B + w_c * (float)b) / CW;
}
/* SDF::interpolate_color, sdf.cpp:164-217, for continuous voxel coordinates (the caller applies
 * get_voxel_coordinates, :170).  fetch(ci,cj,ck, cw,r,g,b) -> false when outside the grid. */
template <class Fetch>
TSDF_HD void interpolate_color(double vx, double vy, double vz, Fetch&& fetch, float out[4]) {
    const float i = (float)vx, j = (float)vy, k = (float)vz;
    const int bi = trunc_f2i(i), bj = trunc_f2i(j), bk = trunc_f2i(k);
    float w_sum = 0.0f, r = 0.0f, g = 0.0f, b = 0.0f;
    out[3] = 1.0f;
    for (int n = 0; n < 8; n++) {
        const int ci = bi + (n >> 2), cj = bj + ((n >> 1) & 1), ck = bk + (n & 1);
        const float volume = (fabsf((float)ci - i) + fabsf((float)cj - j)) + fabsf((float)ck - k);
        float cw, cr, cg, cb;
        if (!fetch(ci, cj, ck, cw, cr, cg, cb)) continue;
        if (!(cw > 0.0f)) continue;
        if (volume <= TSDF_VOL_EXACT_F) { out[0] = cr; out[1] = cg; out[2] = cb; return; }   /* unscaled, :193-198 */
        const float w = rcp_rn(volume);                              /* (float)(1.0 / volume), :200 */
        w_sum = w_sum + w;
        r = r + w * cr; g = g + w * cg; b = b + w * cb;
    }
    const float aux = (float)((double)w_sum * 255.0);                /* :210 */
    out[0] = r / aux; out[1] = g / aux; out[2] = b / aux;
}

/* ---- constants of the certified fp32 evaluations below */
#define FAST_ZMIN 0.05f                     /* closer to the camera plane than this: exact path */
#define FAST_DMARG 1e-4f                    /* margin on signed distances, metres (fp32 error <= ~1.1e-5) */
TSDF_HD float rcp_approx32(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}

/* ---- hierarchical certificates for fusion ------------------------------------------------------
 * Per pixel p (with the point P = z*ray and, for the plane metric, the unit normal n facing the
 * camera, g0 = ray.n <= 0) the signed distance of ANY camera-space point c that the reference maps
 * to p (its projection truncates to p) is  d = P.n - c_z * (g0 + e),  |e| <= ea = |nx|/fx + |ny|/fy,
 * so with a = |g0|:
 *      c_z <  zfree(p)   = (a*z - delta - M) / (a + ea)   =>  d < -delta - M   (updated with d = -delta, w = 1)
 *      c_z >  zbehind(p) = (a*z + delta + M) / (a - ea)   =>  d >  delta + M   (skipped)
 * (M = FAST_DMARG absorbs the rounding of the exact path's own d; the fp32 evaluation of the two
 * bounds is padded by 1e-5 relative + 1e-4 m).  An invalid pixel is skipped whatever c_z is:
 * zfree = -inf, zbehind = -inf.  Point metric: zfree = z - delta - M, zbehind = z + delta + M.
 * A min-pyramid of zfree and a max-pyramid of zbehind then certify a whole group of voxels whose
 * pixels lie in a bounding box: all updated-as-free-space, or all skipped.  Everything not
 * certified takes the exact path; the device self-check compares every certified voxel with it. */
#define CERT_LEVELS 11                      /* levels 0..10: texels of 1..1024 pixels */
struct CertPyramid {
    int32_t w[CERT_LEVELS], h[CERT_LEVELS];
    int64_t off[CERT_LEVELS];               /* element offsets (float2) of each level in one buffer */
};
TSDF_HD void cert_pixel(const GridParams& g, const K1Params& kp, int u, int v, const PixRec& rec, float& zfree, float& zbehind) {
    const float NINF = -3.402823466e+38f, PINF = 3.402823466e+38f;
    zfree = NINF; zbehind = NINF;
    /* The outermost pixel ring never takes part in a free-space certificate (unit_certificate asks for
     * a footprint strictly inside it), so it is neutral for the min-pyramid; otherwise its missing
     * normals would poison every coarse texel that touches the image border. */
    const bool border = (u == 0) | (v == 0) | (u == g.img_w - 1) | (v == g.img_h - 1);
    if (!(rec.z == rec.z)) { if (border) zfree = PINF; return; }
    if (g.metric != 0) {
        zfree = border ? PINF : (rec.z - g.delta - FAST_DMARG) * (1.0f - 1e-5f) - 1e-4f;
        zbehind = (rec.z + g.delta + FAST_DMARG) * (1.0f + 1e-5f) + 1e-4f;
        return;
    }
    if (!(rec.nx == rec.nx)) { if (border) zfree = PINF; return; }
    const float rx = ((float)u - kp.cx) * kp.inv_fx, ry = ((float)v - kp.cy) * kp.inv_fy;
    const float a = -(rx * rec.nx + ry * rec.ny + rec.nz);            /* |ray.n| */
    const float ea = (fabsf(rec.nx) * kp.inv_fx + fabsf(rec.ny) * kp.inv_fy) * 1.001f;   /* cell of +-1 pixel */
    zbehind = PINF;
    if (!(a > 1e-4f)) return;                                         /* grazing: no certificate, never skipped */
    const float num = a * rec.z - g.delta - FAST_DMARG;
    if (border) zfree = PINF; else if (num > 0.0f) zfree = (num / (a + ea)) * (1.0f - 1e-5f) - 1e-4f;
    const float den = a - ea;
    if (den > 1e-4f) zbehind = ((a * rec.z + g.delta + FAST_DMARG) / den) * (1.0f + 1e-5f) + 1e-4f;
}

enum { UNIT_UNKNOWN = 0, UNIT_FRONT = 1, UNIT_SKIP = 2 };
TSDF_HD int bit_length(int x) {             /* number of bits needed for x >= 0 */
#if defined(__CUDA_ARCH__)
    return 32 - __clz(x);
#else
    int n = 0;
    while ((x >> n) > 0) n++;
    return n;
#endif
}

/* Verdict for a group of voxels whose camera-space centres lie on the segment between cA and cB
 * (the lane's four consecutive x voxels): project both ends in fp32 (error < 3e-4 px), take the
 * pixel bounding box dilated by one pixel, query the pyramid at the level where the box spans at
 * most 2x2 texels.  fetch(level, x, y) -> (min zfree, max zbehind) of that texel. */
/* core: float end points, plus an absolute pad on the two depth comparisons (0 for end points that are the
 * rounded double centres; > 0 when the end points come from the fp32 affine evaluation below) */
template <class TexFetch>
TSDF_HD int unit_certificate_f(const GridParams& g, const CertPyramid& P, float XA, float YA, float ZA,
                               float XB, float YB, float ZB, float zpad, TexFetch&& fetch) {
    if (!g.k_simple) return UNIT_UNKNOWN;
    const float zmin = fminf(ZA, ZB), zmax = fmaxf(ZA, ZB);
    if (zmax < -FAST_ZMIN) return UNIT_SKIP;                          /* all behind the camera, sdf.cpp:247 */
    if (!(zmin >= FAST_ZMIN)) return UNIT_UNKNOWN;
    const float ra = rcp_approx32(ZA), rb = rcp_approx32(ZB);
    const float fxf = (float)g.K[0], fyf = (float)g.K[4], cxf = (float)g.K[2], cyf = (float)g.K[5];
    const float ua = fmaf(fxf, XA * ra, cxf), va = fmaf(fyf, YA * ra, cyf);
    const float ub = fmaf(fxf, XB * rb, cxf), vb = fmaf(fyf, YB * rb, cyf);
    if (!(fabsf(ua) < 1e6f && fabsf(ub) < 1e6f && fabsf(va) < 1e6f && fabsf(vb) < 1e6f)) return UNIT_UNKNOWN;
    int u0 = (int)floorf(fminf(ua, ub)) - 1, u1 = (int)floorf(fmaxf(ua, ub)) + 1;
    int v0 = (int)floorf(fminf(va, vb)) - 1, v1 = (int)floorf(fmaxf(va, vb)) + 1;
    if (u1 < -1 || v1 < -1 || u0 > g.img_w || v0 > g.img_h) return UNIT_SKIP;   /* certainly outside the image, sdf.cpp:254 */
    const bool inside = (u0 >= 1) & (v0 >= 1) & (u1 <= g.img_w - 2) & (v1 <= g.img_h - 2);   /* strictly inside the border ring */
    /* clamp BOTH ends into the image: u1 = -1 / u0 = img_w (box touching the image from outside)
     * pass the test above and must not index outside the pyramid */
    u0 = imin(imax(u0, 0), g.img_w - 1); v0 = imin(imax(v0, 0), g.img_h - 1);
    u1 = imax(imin(u1, g.img_w - 1), 0); v1 = imax(imin(v1, g.img_h - 1), 0);
    const int ext = imax(u1 - u0, v1 - v0);                           /* extent - 1 */
    const int level = bit_length(ext);                                /* 2^level >= extent: the box spans <= 2 texels */
    if (level >= CERT_LEVELS) return UNIT_UNKNOWN;
    const int x0 = u0 >> level, x1 = u1 >> level, y0 = v0 >> level, y1 = v1 >> level;
    float f00, b00, f10, b10, f01, b01, f11, b11;
    fetch(level, x0, y0, f00, b00); fetch(level, x1, y0, f10, b10);
    fetch(level, x0, y1, f01, b01); fetch(level, x1, y1, f11, b11);
    const float zfree = fminf(fminf(f00, f10), fminf(f01, f11));
    const float zbehind = fmaxf(fmaxf(b00, b10), fmaxf(b01, b11));
    if (inside && zmax * (1.0f + 1e-6f) + zpad < zfree) return UNIT_FRONT;
    if (zmin * (1.0f - 1e-6f) - zpad > zbehind) return UNIT_SKIP;    /* pixels outside the image skip anyway */
    return UNIT_UNKNOWN;
}
template <class TexFetch>
TSDF_HD int unit_certificate(const GridParams& g, const CertPyramid& P, double ax, double ay, double az,
                             double bx, double by, double bz, TexFetch&& fetch) {
    return unit_certificate_f(g, P, (float)ax, (float)ay, (float)az, (float)bx, (float)by, (float)bz, 0.0f, fetch);
}

/* fp32 affine evaluation of a row's camera-space centres for the per-unit certificates.  Along a grid row the
 * exact centre is affine in i:  c(i) = c(0) + i * step,  step = Rinv(:,0) * vs_x  (the tables hold Rinv(r,0) * gx(i)
 * with gx(i) = vs_x (i + 0.5) + origin_x).  c0 = fl32(c(0)) per row, step = fl32(step) per frame, and
 * c~(i) = fmaf(i, step, c0).  Error: |c0| rounding <= ulp/2, step rounding <= 6e-8 * |i step|, fmaf rounding
 * <= ulp/2: for centres within +-32 m and rows of <= 4092 voxels that is < 8e-6 m per coordinate.  It enters the
 * certificate in two places: the pixel positions (f * 8e-6 / FAST_ZMIN < 0.1 px for f < 600: absorbed by the
 * one-pixel dilation of the box) and the two depth comparisons (padded by AFFINE_ZPAD = 2.5e-5 m).  The device
 * self-check (tsdf_debug_fuse_check) and the CPU emulation compare every certified voxel with the exact path. */
#define AFFINE_ZPAD 2.5e-5f
#define AFFINE_MAX_COORD 32.0         /* metres */
#define AFFINE_MAX_FOCAL 1200.0       /* pixels: 8e-6 m * (1 + |x/z|) * f / FAST_ZMIN stays below half a pixel */
/* per frame: every voxel centre lies within AFFINE_MAX_COORD of the camera (a rigid transform keeps distances, so
 * the farthest corner of the volume bounds every camera-space coordinate) and the focal lengths are moderate;
 * otherwise the certificates take the double-precision end points */
TSDF_HD bool affine_ok(const GridParams& g, const double* t) {
    if (!g.k_simple || g.K[0] > AFFINE_MAX_FOCAL || g.K[4] > AFFINE_MAX_FOCAL) return false;
    double far2 = 0.0;
    for (int q = 0; q < 8; q++) {
        const double x = g.origin[0] + ((q & 1) ? (double)g.vs_x * g.m : 0.0) - t[0];
        const double y = g.origin[1] + ((q & 2) ? (double)g.vs_y * g.m : 0.0) - t[1];
        const double z = g.origin[2] + ((q & 4) ? (double)g.vs_z * g.m : 0.0) - t[2];
        far2 = fmax(far2, x * x + y * y + z * z);
    }
    return far2 <= AFFINE_MAX_COORD * AFFINE_MAX_COORD;
}
TSDF_HD void affine_step(const GridParams& g, const double* Rinv, float& sx, float& sy, float& sz) {
    sx = (float)(Rinv[0] * (double)g.vs_x); sy = (float)(Rinv[3] * (double)g.vs_x); sz = (float)(Rinv[6] * (double)g.vs_x);
}
template <class TexFetch>
TSDF_HD int unit_certificate_affine(const GridParams& g, const CertPyramid& P, float c0x, float c0y, float c0z,
                                    float sx, float sy, float sz, int x0, TexFetch&& fetch) {
    const float ia = (float)x0, ib = (float)(x0 + 3);
    return unit_certificate_f(g, P, fmaf(ia, sx, c0x), fmaf(ia, sy, c0y), fmaf(ia, sz, c0z),
                              fmaf(ib, sx, c0x), fmaf(ib, sy, c0y), fmaf(ib, sz, c0z), AFFINE_ZPAD, fetch);
}

/* Verdict for a BRICK of voxels (an axis-aligned box of the grid): the camera-space centres of all its voxels lie
 * in the convex hull of its eight corner centres c[0..7], and projection maps a convex set in front of the camera
 * into the convex hull of the projected corners, so the pixel bounding box of the corners (dilated by one pixel, as
 * in unit_certificate) contains every voxel's pixel and [zmin, zmax] of the corners contains every voxel's depth.
 * Same pyramid query and the same two tests as unit_certificate. */
template <class TexFetch>
TSDF_HD int box_certificate(const GridParams& g, const CertPyramid& P, const float* cx, const float* cy, const float* cz, TexFetch&& fetch) {
    if (!g.k_simple) return UNIT_UNKNOWN;
    float zmin = cz[0], zmax = cz[0];
#pragma unroll
    for (int q = 1; q < 8; q++) { zmin = fminf(zmin, cz[q]); zmax = fmaxf(zmax, cz[q]); }
    if (zmax < -FAST_ZMIN) return UNIT_SKIP;                          /* all behind the camera, sdf.cpp:247 */
    if (!(zmin >= FAST_ZMIN)) return UNIT_UNKNOWN;
    const float fxf = (float)g.K[0], fyf = (float)g.K[4], cxf = (float)g.K[2], cyf = (float)g.K[5];
    float umin = 3.0e38f, umax = -3.0e38f, vmin = 3.0e38f, vmax = -3.0e38f;
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const float r = rcp_approx32(cz[q]);
        const float u = fmaf(fxf, cx[q] * r, cxf), v = fmaf(fyf, cy[q] * r, cyf);
        umin = fminf(umin, u); umax = fmaxf(umax, u); vmin = fminf(vmin, v); vmax = fmaxf(vmax, v);
    }
    if (!(fabsf(umin) < 1e6f && fabsf(umax) < 1e6f && fabsf(vmin) < 1e6f && fabsf(vmax) < 1e6f)) return UNIT_UNKNOWN;
    int u0 = (int)floorf(umin) - 1, u1 = (int)floorf(umax) + 1;
    int v0 = (int)floorf(vmin) - 1, v1 = (int)floorf(vmax) + 1;
    if (u1 < -1 || v1 < -1 || u0 > g.img_w || v0 > g.img_h) return UNIT_SKIP;   /* certainly outside the image, sdf.cpp:254 */
    const bool inside = (u0 >= 1) & (v0 >= 1) & (u1 <= g.img_w - 2) & (v1 <= g.img_h - 2);
    u0 = imin(imax(u0, 0), g.img_w - 1); v0 = imin(imax(v0, 0), g.img_h - 1);
    u1 = imax(imin(u1, g.img_w - 1), 0); v1 = imax(imin(v1, g.img_h - 1), 0);
    const int ext = imax(u1 - u0, v1 - v0);
    const int level = bit_length(ext);
    if (level >= CERT_LEVELS) return UNIT_UNKNOWN;
    const int x0 = u0 >> level, x1 = u1 >> level, y0 = v0 >> level, y1 = v1 >> level;
    float f00, b00, f10, b10, f01, b01, f11, b11;
    fetch(level, x0, y0, f00, b00); fetch(level, x1, y0, f10, b10);
    fetch(level, x0, y1, f01, b01); fetch(level, x1, y1, f11, b11);
    const float zfree = fminf(fminf(f00, f10), fminf(f01, f11));
    const float zbehind = fmaxf(fmaxf(b00, b10), fmaxf(b01, b11));
    if (inside && zmax * (1.0f + 1e-6f) < zfree) return UNIT_FRONT;
    if (zmin * (1.0f - 1e-6f) > zbehind) return UNIT_SKIP;
    return UNIT_UNKNOWN;
}

/* ---- scan-line clipping for the fusion kernel ------------------------------------------------
 * Along a grid row (fixed j,k; i = 0..m-1) the camera-space centre is affine in i, so each of
 * the five acceptance tests of sdf.cpp:247-254 (z >= 0, -1 < u < width, -1 < v < height, the
 * last four multiplied through by z > 0) is a linear inequality f(i) >= 0 and the accepted
 * voxels form one interval.  This computes a CONSERVATIVE superset [ilo, ihi) of it (tolerance
 * far above the rounding error of the evaluation, interval widened by a voxel); the exact
 * per-voxel test still decides.  py*, pz* = Rinv(r,1)*gy, Rinv(r,2)*gz. */
TSDF_HD void clip_constraint(double f0, double f1, double span, double& lo, double& hi, bool& empty) {
    const double tol = 1e-6 * (fabs(f0) + fabs(f1)) + 1e-9;
    const double g0 = f0 + tol, g1 = f1 + tol;
    if (g0 < 0.0 && g1 < 0.0) { empty = true; return; }
    if (g0 >= 0.0 && g1 >= 0.0) return;
    const double sx = g0 / (g0 - g1) * span;
    if (g0 < 0.0) lo = fmax(lo, sx); else hi = fmin(hi, sx);
}
TSDF_HD void row_clip(const GridParams& g, const double* Ri, const double* ti,
                      double py0, double py1, double py2, double pz0, double pz1, double pz2,
                      int& ilo, int& ihi) {
    const int m = g.m;
    ilo = 0; ihi = 0;
    double lo = 0.0, hi = (double)(m - 1);
    bool empty = false;
    const double gx0 = voxel_centre(g.vs_x, 0, g.origin[0]), gx1 = voxel_centre(g.vs_x, m - 1, g.origin[0]);
    const double ax = ((Ri[0] * gx0 + py0) + pz0) + ti[0], bx = ((Ri[0] * gx1 + py0) + pz0) + ti[0];
    const double ay = ((Ri[3] * gx0 + py1) + pz1) + ti[1], by = ((Ri[3] * gx1 + py1) + pz1) + ti[1];
    const double az = ((Ri[6] * gx0 + py2) + pz2) + ti[2], bz = ((Ri[6] * gx1 + py2) + pz2) + ti[2];
    const double span = (double)(m - 1);
    const double Wd = (double)g.img_w, Hd = (double)g.img_h;
    clip_constraint(az, bz, span, lo, hi, empty);                               /* z >= 0 */
    if (g.k_simple) {
        /* all five constraints are evaluated without early exits: their divisions are independent and overlap
         * (early returns were measured slower: they serialise the chain in a latency-bound kernel) */
        const double a0 = g.K[0] * ax + g.K[2] * az, b0 = g.K[0] * bx + g.K[2] * bz;   /* ij0 */
        const double a1 = g.K[4] * ay + g.K[5] * az, b1 = g.K[4] * by + g.K[5] * bz;   /* ij1 */
        clip_constraint(a0 + az, b0 + bz, span, lo, hi, empty);                 /* u > -1     */
        clip_constraint(Wd * az - a0, Wd * bz - b0, span, lo, hi, empty);       /* u < width  */
        clip_constraint(a1 + az, b1 + bz, span, lo, hi, empty);                 /* v > -1     */
        clip_constraint(Hd * az - a1, Hd * bz - b1, span, lo, hi, empty);       /* v < height */
    }
    if (!empty && lo <= hi) {
        int l = (int)floor(lo) - 1, h = (int)ceil(hi) + 2;
        l = l < 0 ? 0 : l; h = h > m ? m : h;
        ilo = l & ~3;                       /* whole 32-byte sectors (m is a multiple of 4) */
        ihi = (h + 3) & ~3;
    }
}

/* ---- tracker: sample coordinates, camera_tracking.cpp:259-260, 273-361 -----------------------
 * s = 0 centre; 1..6 = +x,-x,+y,-y,+z,-z voxel steps; 7..12 = r1p,r1m,r2p,r2m,r3p,r3m.
 * M = rot for s < 7, the perturbed rotation otherwise.  cvx.. returns the UN-stepped voxel
 * coordinate (the centre for s < 7) for the bounds test of :261-268. */
TSDF_HD void sample_offsets(const GridParams& g, int s, double& ox, double& oy, double& oz) {
    /* +-v_h on one axis for s = 1..6 (x + (-v_h) == x - v_h; x + 0.0 == x) */
    const double vh = (double)g.v_h;
    ox = (s == 1) ? vh : (s == 2) ? -vh : 0.0;
    oy = (s == 3) ? vh : (s == 4) ? -vh : 0.0;
    oz = (s == 5) ? vh : (s == 6) ? -vh : 0.0;
}
TSDF_HD void sample_coords_off(const GridParams& g, const double* M, const double* t, double ox, double oy, double oz,
                               double px, double py, double pz, double& vx, double& vy, double& vz) {
    double wx, wy, wz;
    matvec3(M, px, py, pz, wx, wy, wz);
    wx = wx + t[0]; wy = wy + t[1]; wz = wz + t[2];
    world_to_voxel(g, wx, wy, wz, vx, vy, vz);
    vx = vx + ox; vy = vy + oy; vz = vz + oz;
}
TSDF_HD void sample_coords(const GridParams& g, const double* M, const double* t, int s,
                           double px, double py, double pz,
                           double& vx, double& vy, double& vz) {
    double ox, oy, oz;
    sample_offsets(g, s, ox, oy, oz);
    sample_coords_off(g, M, t, ox, oy, oz, px, py, pz, vx, vy, vz);
}
/* camera_tracking.cpp:92-145: perturbed rotation q (0..5) = (I +- w_h [e_k]x) * rot */
TSDF_HD void perturbed_rot(const GridParams& g, const double* rot, int q, double* out) {
    const double w_h = (double)g.w_h;
    double Rd[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    switch (q) {
        case 0: Rd[5] = -w_h; Rd[7] = w_h; break;
        case 1: Rd[5] = w_h; Rd[7] = -w_h; break;
        case 2: Rd[2] = w_h; Rd[6] = -w_h; break;
        case 3: Rd[2] = -w_h; Rd[6] = w_h; break;
        case 4: Rd[1] = -w_h; Rd[3] = w_h; break;
        default: Rd[1] = w_h; Rd[3] = -w_h; break;
    }
    matmul3(Rd, rot, out);
}

/* ---- k_linearize's work distribution (tsdf_kernels.cu: "Work distribution"): the strided pixel grid (ni columns x nj
 * rows) is cut into 4 x 4 micro-tiles numbered column by column; a sweep of a block covers mt_sweep micro-tiles, one
 * pixel per thread pair.  lin_layout: the numbers every block derives from the launch; lin_pixel_of: strided pixel
 * (ii, jj) of slot tl (0 .. 16 * mt_sweep - 1) in sweep q of block b, false when the slot is empty.
 *   unsharded: sweep q of block b = the mt_sweep ADJACENT micro-tiles (q * nblocks + b) * mt_sweep ...
 *   sharded:   slot sl of block b = micro-tile b + sl * nblocks (spread: a rank owns a compact image region)
 * Shared with the host build (tests/host_emul): every pixel must be covered exactly once for any image size. */
struct LinLayout { int mty, n_micro, n_slots, n_sweeps; };
TSDF_HD LinLayout lin_layout(int ni, int nj, int nblocks, int mt_sweep) {
    LinLayout L;
    L.mty = (nj + 3) >> 2;
    L.n_micro = ((ni + 3) >> 2) * L.mty;
    L.n_slots = (L.n_micro + nblocks - 1) / nblocks;                         /* micro-tiles per block */
    L.n_sweeps = mt_sweep > 0 ? (L.n_slots + mt_sweep - 1) / mt_sweep : 0;
    return L;
}
TSDF_HD bool lin_pixel_of(const LinLayout& L, int ni, int nj, int nblocks, int b, int mt_sweep, bool sharded, int q, int tl, int& ii, int& jj) {
    const int sl = q * mt_sweep + (tl >> 4);
    const int mt = sharded ? b + sl * nblocks : (q * nblocks + b) * mt_sweep + (tl >> 4);
    const int mx = mt / L.mty, my = mt - mx * L.mty;
    ii = (mx << 2) + ((tl & 15) >> 2); jj = (my << 2) + (tl & 3);
    return (sharded ? sl < L.n_slots : true) & (mt < L.n_micro) & (ii < ni) & (jj < nj);
}

/* slot layout of the reduced normal equations */
enum { SLOT_A = 0, SLOT_B = 21, SLOT_RES = 27, SLOT_NVALID = 28, SLOT_NOOB = 29, N_SLOTS = 30, SLOT_MISS = 30 /* local, not exchanged */ };

/* ---- 6x6 partial-pivot LU solve, stands in for Eigen's A.inverse()*b (camera_tracking.cpp:191) */
/* one elimination column; C is a template parameter so that every array index below is a
 * compile-time constant and A, b stay in registers on the device (the pivot row is applied with
 * conditional swaps instead of a dynamic index) */
template <int C>
struct Solve6Col {
    static TSDF_HD void run(double (&A)[6][6], double (&b)[6], double (&inv)[6], int& singular) {
        int piv = C;
        double best = fabs(A[C][C]);
#pragma unroll
        for (int r = C + 1; r < 6; r++) {
            const double v = fabs(A[r][C]);
            if (v > best) { best = v; piv = r; }
        }
        const bool good = (best > 0.0);
        if (!good) singular = 1;
#pragma unroll
        for (int r = C + 1; r < 6; r++) {
            const bool sw = good && (piv == r);
#pragma unroll
            for (int q = C; q < 6; q++) {
                const double u = A[C][q], v = A[r][q];
                A[C][q] = sw ? v : u; A[r][q] = sw ? u : v;
            }
            const double u = b[C], v = b[r];
            b[C] = sw ? v : u; b[r] = sw ? u : v;
        }
        inv[C] = 1.0 / A[C][C];                          /* one reciprocal per pivot, reused by the back substitution */
#pragma unroll
        for (int r = C + 1; r < 6; r++) {
            const double f = A[r][C] * inv[C];
#pragma unroll
            for (int q = C + 1; q < 6; q++) A[r][q] = good ? (A[r][q] - f * A[C][q]) : A[r][q];
            b[r] = good ? (b[r] - f * b[C]) : b[r];
        }
        Solve6Col<C + 1>::run(A, b, inv, singular);
    }
};
template <>
struct Solve6Col<6> {
    static TSDF_HD void run(double (&)[6][6], double (&)[6], double (&)[6], int&) {}
};
template <int R>
struct Solve6Back {
    static TSDF_HD void run(const double (&A)[6][6], const double (&b)[6], const double (&inv)[6], double* x) {
        double sacc = b[R];
#pragma unroll
        for (int q = R + 1; q < 6; q++) sacc = sacc - A[R][q] * x[q];
        x[R] = sacc * inv[R];
        Solve6Back<R - 1>::run(A, b, inv, x);
    }
};
template <>
struct Solve6Back<-1> {
    static TSDF_HD void run(const double (&)[6][6], const double (&)[6], const double (&)[6], double*) {}
};
TSDF_HD int solve6(const double* Ain, const double* bin, double* x) {
    double A[6][6], b[6];
#pragma unroll
    for (int r = 0; r < 6; r++) {
#pragma unroll
        for (int c = 0; c < 6; c++) A[r][c] = Ain[6 * r + c];
        b[r] = bin[r];
    }
    int singular = 0;
    double inv[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    Solve6Col<0>::run(A, b, inv, singular);
    Solve6Back<5>::run(A, b, inv, x);
#pragma unroll
    for (int q = 0; q < 6; q++) if (!(fabs(x[q]) <= 1.7976931348623157e308)) singular = 1;
    return singular;
}

/* ---- eigen_utils.cpp:40-128 (delta_t = 1) */
TSDF_HD void exp_map(const double* v, double* rd, double* dt) {
    const double ang_min_sinc = 1.0e-8, ang_min_mc = 2.5e-4;
    const double u0 = v[3], u1 = v[4], u2 = v[5];
    const double theta = sqrt(u0 * u0 + u1 * u1 + u2 * u2);
    double si, co;
    sincos(theta, &si, &co);
    const double sinc = (fabs(theta) < ang_min_sinc) ? 1.0 : (si / theta);
    const double mcosc = (fabs(theta) < ang_min_mc) ? 0.5 : ((1.0 - co) / theta / theta);
    const double msinc = (fabs(theta) < ang_min_mc) ? (1. / 6.0) : ((1.0 - si / theta) / theta / theta);
    rd[0] = co + mcosc * u0 * u0;
    rd[1] = -sinc * u2 + mcosc * u0 * u1;
    rd[2] = sinc * u1 + mcosc * u0 * u2;
    rd[3] = sinc * u2 + mcosc * u1 * u0;
    rd[4] = co + mcosc * u1 * u1;
    rd[5] = -sinc * u0 + mcosc * u1 * u2;
    rd[6] = -sinc * u1 + mcosc * u2 * u0;
    rd[7] = sinc * u0 + mcosc * u2 * u1;
    rd[8] = co + mcosc * u2 * u2;
    dt[0] = v[0] * (sinc + u0 * u0 * msinc) + v[1] * (u0 * u1 * msinc - u2 * mcosc) + v[2] * (u0 * u2 * msinc + u1 * mcosc);
    dt[1] = v[0] * (u0 * u1 * msinc + u2 * mcosc) + v[1] * (sinc + u1 * u1 * msinc) + v[2] * (u1 * u2 * msinc - u0 * mcosc);
    dt[2] = v[0] * (u0 * u2 * msinc - u1 * mcosc) + v[1] * (u1 * u2 * msinc + u0 * mcosc) + v[2] * (sinc + u2 * u2 * msinc);
}

/* Gauss-Newton step: solve, exp, pose update, signed stop test.
 * camera_tracking.cpp:191-192, 216-224, 237-239.  sums = the 30 reduced slots. */
TSDF_HD void gn_update(const GridParams& g, PoseState& p, const double* sums) {
    double A[36], b[6];
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
        for (int c = r; c < 6; c++) {
            const int q = r * 6 - (r * (r - 1)) / 2 + (c - r);       /* index in the row-major upper triangle */
            A[6 * r + c] = sums[SLOT_A + q]; A[6 * c + r] = sums[SLOT_A + q];
        }
#pragma unroll
    for (int r = 0; r < 6; r++) b[r] = sums[SLOT_B + r];
    for (int s = 0; s < N_SLOTS; s++) p.sums[s] = sums[s];
    double tw[6];
    const int singular = solve6(A, b, tw);
    p.iterations = p.iterations + 1;
    if (singular) { p.singular = 1; p.stopped = 1; return; }    /* keep the previous pose, report */
#pragma unroll
    for (int s = 0; s < 6; s++) p.twist[s] = tw[s];
    double rd[9], td[3];
    exp_map(tw, rd, td);
    const double rdT[9] = {rd[0], rd[3], rd[6], rd[1], rd[4], rd[7], rd[2], rd[5], rd[8]};
    double newR[9];
    matmul3(rdT, p.R, newR);                                     /* :237 */
    double x, y, z;
    matvec3(rdT, td[0], td[1], td[2], x, y, z);
    const double newt[3] = {p.t[0] - x, p.t[1] - y, p.t[2] - z};  /* :238 */
    pose_set(p, newR, newt);                                     /* :239 */
    const double mtd = (double)g.max_twist_diff;
    if (tw[0] < mtd && tw[1] < mtd && tw[2] < mtd && tw[3] < mtd && tw[4] < mtd && tw[5] < mtd) p.stopped = 1;   /* :216-224 */
}

}  // namespace tsdf
