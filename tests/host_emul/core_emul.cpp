/*
 * core_emul.cpp — HOST compile of the kernels' __host__ __device__ core (tsdf_core.cuh).
 *
 * TEST INFRASTRUCTURE ONLY (tests/ -m "not gpu").  There is no GPU in the build container, so
 * this lets the CPU suite check the exact-arithmetic building blocks the CUDA kernels are made
 * of — the hoisted camera-space sums, the scan-line clip, the per-sample interpolation, the
 * J/psi construction, the 6x6 solve / exp map / pose update — against the oracle, bit for bit.
 * The loops below mirror the kernels' work decomposition serially.  It is never loaded by the
 * tracking_sdf_b200 package and is not a fallback.
 */
#include <cstdint>
#include <cstring>
#include <vector>
#include "tsdf_core.cuh"
#include "mc_core.cuh"

using namespace tsdf;

extern "C" {

/* fills a GridParams the same way tsdf_abi.cu:build_params does (kept in sync by
 * tests/test_core_emul.py::test_params_match_oracle_constants) */
void emul_params(int m, float width, float height, float depth, const double origin[3], float delta, float eps,
                 float v_h, float w_h, int stride, int metric, int img_w, int img_h, const double K[9],
                 float max_twist_diff, int max_iter, GridParams* g) {
    memset(g, 0, sizeof *g);
    g->m = m; g->ks0 = 0; g->ks1 = m; g->ko0 = 0; g->ko1 = m;
    g->metric = metric; g->img_w = img_w; g->img_h = img_h; g->stride = stride;
    g->ni = (img_w + stride - 1) / stride; g->nj = (img_h + stride - 1) / stride;
    g->m_div_height = m / height; g->m_div_width = m / width; g->m_div_depth = m / depth;
    g->m_div_d[0] = (double)g->m_div_width; g->m_div_d[1] = (double)g->m_div_height; g->m_div_d[2] = (double)g->m_div_depth;
    g->vs_x = width / ((float)m); g->vs_y = height / ((float)m); g->vs_z = depth / ((float)m);
    g->delta = delta; g->eps = eps; g->v_h = v_h; g->w_h = w_h;
    const float v_h2 = 2 * v_h;
    g->v_h2_width = v_h2 / g->m_div_width; g->v_h2_height = v_h2 / g->m_div_height; g->v_h2_depth = v_h2 / g->m_div_depth;
    g->two_w_h = 2 * (w_h);
    g->max_twist_diff = max_twist_diff; g->max_iter = max_iter;
    for (int q = 0; q < 3; q++) g->origin[q] = origin[q];
    for (int q = 0; q < 9; q++) g->K[q] = K[q];
    g->k_simple = (K[1] == 0.0 && K[3] == 0.0 && K[6] == 0.0 && K[7] == 0.0 && K[8] == 1.0) ? 1 : 0;
}
int emul_sizeof_params() { return (int)sizeof(GridParams); }
int emul_sizeof_pose() { return (int)sizeof(PoseState); }

void emul_pose_set(PoseState* p, const double R[9], const double t[3]) { memset(p, 0, sizeof *p); pose_set(*p, R, t); }
void emul_pose_get(const PoseState* p, double R[9], double t[3], double Rinv[9], double tinv[3]) {
    memcpy(R, p->R, sizeof p->R); memcpy(t, p->t, sizeof p->t);
    memcpy(Rinv, p->Rinv, sizeof p->Rinv); memcpy(tinv, p->tinv, sizeof p->tinv);
}

/* K1 */
void emul_prep(const GridParams* g, const float* depth, float* pix /* [h*w*4] */) {
    const K1Params kp = k1_params(g->K);
    const float qnan = std::nanf("");
    for (int v = 0; v < g->img_h; v++)
        for (int u = 0; u < g->img_w; u++) {
            const size_t o = (size_t)v * g->img_w + u;
            const float zc = depth[o];
            float r[4] = {depth_valid(zc) ? zc : qnan, qnan, qnan, qnan};
            if (u > 0 && v > 0 && u < g->img_w - 1 && v < g->img_h - 1) {
                float nx, ny, nz;
                if (normal_px(kp, u, v, zc, depth[o - 1], depth[o + 1], depth[o - g->img_w], depth[o + g->img_w], nx, ny, nz)) {
                    r[1] = nx; r[2] = ny; r[3] = nz;
                }
            }
            memcpy(pix + 4 * o, r, sizeof r);
        }
}
void emul_cloud(const GridParams* g, const float* pix, float* cloud, float* normals) {
    const K1Params kp = k1_params(g->K);
    for (int o = 0; o < g->img_w * g->img_h; o++) {
        const int u = o % g->img_w, v = o / g->img_w;
        float x, y;
        backproject_px(kp, u, v, pix[4 * o], x, y);
        cloud[3 * o] = x; cloud[3 * o + 1] = y; cloud[3 * o + 2] = pix[4 * o];
        normals[3 * o] = pix[4 * o + 1]; normals[3 * o + 2] = pix[4 * o + 3]; normals[3 * o + 1] = pix[4 * o + 2];
    }
}

/* grid here: x-fastest interleaved {D,W}, like the device store */
struct HostFetch {
    const float* grid; int m;
    bool operator()(int ci, int cj, int ck, float& d, float& w) const {
        if ((unsigned)ci >= (unsigned)m || (unsigned)cj >= (unsigned)m || (unsigned)ck >= (unsigned)m) return false;
        const size_t o = (((size_t)ck * m + cj) * m + ci) * 2;
        d = grid[o]; w = grid[o + 1];
        return true;
    }
    bool interior(int bi, int bj, int bk) const {
        return bi >= 0 && bj >= 0 && bk >= 0 && bi < m - 1 && bj < m - 1 && bk < m - 1;
    }
    void load8(int bi, int bj, int bk, float* d, float* w) const {
        for (int n = 0; n < 8; n++) (*this)(bi + (n >> 2), bj + ((n >> 1) & 1), bk + (n & 1), d[n], w[n]);
    }
};

void emul_interpolate(const GridParams* g, const float* grid, int64_t n, const double* pts, float* out, uint8_t* ok) {
    HostFetch f{grid, g->m};
    for (int64_t q = 0; q < n; q++) {
        bool is_interp;
        out[q] = interpolate_distance(pts[3 * q], pts[3 * q + 1], pts[3 * q + 2], f, is_interp);
        ok[q] = is_interp;
    }
}

/* K3, mirroring the fusion kernels: rows clipped (k_fuse_plan), lane units of four voxels certified
 * against the pyramid (k_prep level 0 + k_pyramid + k_fuse_cert), the rest through the exact path
 * (k_fuse_exact), products hoisted exactly like the kernels.
 * use_clip = 0 visits every voxel; use_cert = 0 sends every unit through the exact path. */
static int64_t g_last_fast = 0, g_rows_front = 0, g_rows_skip = 0, g_rows_unknown = 0;
void emul_row_stats(int64_t out[3]) { out[0] = g_rows_front; out[1] = g_rows_skip; out[2] = g_rows_unknown; }
int64_t emul_last_fast_count() { return g_last_fast; }
int64_t emul_fuse_rgb(const GridParams* gp, float* grid, const float* pix, const PoseState* pose, int use_clip, int use_cert,
                      float* color /* [k][j][i]{Color_W,R,G,B} or NULL */, const uint8_t* rgb3 /* [h*w*3] or NULL */);
int64_t emul_fuse(const GridParams* gp, float* grid, const float* pix, const PoseState* pose, int use_clip, int use_cert) {
    return emul_fuse_rgb(gp, grid, pix, pose, use_clip, use_cert, nullptr, nullptr);
}
/* with colour (k_fuse_cert in queue_front mode + k_fuse_exact<colour>): the free-space certificate no longer
 * updates in place, the exact pass takes the unit WITH its certificate (d = -delta, w = 1 without the distance
 * arithmetic) and every updated voxel gets the colour running mean from its pixel's tabulated cosine and rgb */
int64_t emul_fuse_rgb(const GridParams* gp, float* grid, const float* pix, const PoseState* pose, int use_clip, int use_cert,
                      float* color, const uint8_t* rgb3) {
    const GridParams& g = *gp;
    const int m = g.m;
    const double* Ri = pose->Rinv; const double* ti = pose->tinv;
    const K1Params kp = k1_params(g.K);
    /* certificate pyramid */
    CertPyramid P;
    std::vector<std::vector<float>> zf(CERT_LEVELS), zb(CERT_LEVELS);
    for (int l = 0; l < CERT_LEVELS; l++) {
        P.w[l] = (g.img_w + (1 << l) - 1) >> l; P.h[l] = (g.img_h + (1 << l) - 1) >> l; P.off[l] = 0;
        zf[l].assign((size_t)P.w[l] * P.h[l], 3.402823466e+38f); zb[l].assign((size_t)P.w[l] * P.h[l], -3.402823466e+38f);
    }
    for (int v = 0; v < g.img_h; v++)
        for (int u = 0; u < g.img_w; u++) {
            const float* q = pix + 4 * ((size_t)v * g.img_w + u);
            PixRec r; r.z = q[0]; r.nx = q[1]; r.ny = q[2]; r.nz = q[3];
            cert_pixel(g, kp, u, v, r, zf[0][(size_t)v * P.w[0] + u], zb[0][(size_t)v * P.w[0] + u]);
        }
    for (int l = 1; l < CERT_LEVELS; l++)
        for (int y = 0; y < P.h[l - 1]; y++)
            for (int x = 0; x < P.w[l - 1]; x++) {
                const size_t o = (size_t)(y >> 1) * P.w[l] + (x >> 1), i = (size_t)y * P.w[l - 1] + x;
                zf[l][o] = fminf(zf[l][o], zf[l - 1][i]); zb[l][o] = fmaxf(zb[l][o], zb[l - 1][i]);
            }
    auto fetch = [&](int level, int x, int y, float& f, float& b) { f = zf[level][(size_t)y * P.w[level] + x]; b = zb[level][(size_t)y * P.w[level] + x]; };
    int64_t n_updated = 0, n_fast = 0;
    g_rows_front = g_rows_skip = g_rows_unknown = 0;
    /* per-unit certificates by the fp32 affine form of the row, exactly as k_fuse_cert evaluates them */
    const bool affine = affine_ok(g, pose->t);
    float stx, sty, stz;
    affine_step(g, Ri, stx, sty, stz);
#pragma omp parallel for reduction(+ : n_updated, n_fast) schedule(dynamic, 1)
    for (int k = g.ks0; k < g.ks1; k++) {
        const double gz = voxel_centre(g.vs_z, k, g.origin[2]);
        const double pz0 = Ri[2] * gz, pz1 = Ri[5] * gz, pz2 = Ri[8] * gz;
        for (int j = 0; j < m; j++) {
            const double gy = voxel_centre(g.vs_y, j, g.origin[1]);
            const double py0 = Ri[1] * gy, py1 = Ri[4] * gy, py2 = Ri[7] * gy;
            int ilo = 0, ihi = m;
            if (use_clip) row_clip(g, Ri, ti, py0, py1, py2, pz0, pz1, pz2, ilo, ihi);
            int rowv = UNIT_UNKNOWN, itemv = UNIT_UNKNOWN;
            const double gx0 = voxel_centre(g.vs_x, 0, g.origin[0]);
            const float c0x = (float)(((Ri[0] * gx0 + py0) + pz0) + ti[0]), c0y = (float)(((Ri[3] * gx0 + py1) + pz1) + ti[1]),
                        c0z = (float)(((Ri[6] * gx0 + py2) + pz2) + ti[2]);
            if (use_cert && ihi > ilo) {
                const double gxa = voxel_centre(g.vs_x, ilo, g.origin[0]), gxb = voxel_centre(g.vs_x, ihi - 1, g.origin[0]);
                rowv = unit_certificate(g, P, ((Ri[0] * gxa + py0) + pz0) + ti[0], ((Ri[3] * gxa + py1) + pz1) + ti[1], ((Ri[6] * gxa + py2) + pz2) + ti[2],
                                        ((Ri[0] * gxb + py0) + pz0) + ti[0], ((Ri[3] * gxb + py1) + pz1) + ti[1], ((Ri[6] * gxb + py2) + pz2) + ti[2], fetch);
#pragma omp critical
                { if (rowv == UNIT_FRONT) g_rows_front++; else if (rowv == UNIT_SKIP) g_rows_skip++; else g_rows_unknown++; }
                if (rowv == UNIT_SKIP) { n_fast += ihi - ilo; continue; }
            }
            for (int x0 = ilo; x0 < ihi; x0 += 4) {
                /* item-level certificate (128 voxels), like k_fuse_plan */
                if (use_cert && rowv == UNIT_UNKNOWN && ((x0 - ilo) & 127) == 0) {
                    const int xa = x0, xb = (xa + 127 < ihi - 1) ? xa + 127 : ihi - 1;
                    const double gxa = voxel_centre(g.vs_x, xa, g.origin[0]), gxb = voxel_centre(g.vs_x, xb, g.origin[0]);
                    itemv = unit_certificate(g, P, ((Ri[0] * gxa + py0) + pz0) + ti[0], ((Ri[3] * gxa + py1) + pz1) + ti[1], ((Ri[6] * gxa + py2) + pz2) + ti[2],
                                             ((Ri[0] * gxb + py0) + pz0) + ti[0], ((Ri[3] * gxb + py1) + pz1) + ti[1], ((Ri[6] * gxb + py2) + pz2) + ti[2], fetch);
                }
                double cx[4], cy[4], cz[4];
                for (int v = 0; v < 4; v++) {
                    const double gx = voxel_centre(g.vs_x, x0 + v, g.origin[0]);
                    cx[v] = ((Ri[0] * gx + py0) + pz0) + ti[0];
                    cy[v] = ((Ri[3] * gx + py1) + pz1) + ti[1];
                    cz[v] = ((Ri[6] * gx + py2) + pz2) + ti[2];
                }
                const int verdict = !use_cert ? UNIT_UNKNOWN : (rowv != UNIT_UNKNOWN ? rowv : itemv != UNIT_UNKNOWN ? itemv :
                                    affine ? unit_certificate_affine(g, P, c0x, c0y, c0z, stx, sty, stz, x0, fetch)
                                           : unit_certificate(g, P, cx[0], cy[0], cz[0], cx[3], cy[3], cz[3], fetch));
                if (verdict == UNIT_SKIP) { n_fast += 4; continue; }
                for (int v = 0; v < 4; v++) {
                    const size_t o = (((size_t)(k - g.ks0) * m + j) * m + x0 + v) * 2;
                    if (verdict == UNIT_FRONT && !color) {
                        fuse_apply(grid[o], grid[o + 1], -g.delta, 1.0f);
                        n_updated++; n_fast++;
                        continue;
                    }
                    int iu, iv;
                    bool ok, need_exact;
                    fuse_project_flags(g, cx[v], cy[v], cz[v], iu, iv, ok, need_exact);
                    if (need_exact) {
                        double ij0, ij1, ij2;
                        project_ij(g, cx[v], cy[v], cz[v], ij0, ij1, ij2);
                        int eu_ = 0, ev_ = 0;
                        ok = project_exact(g, ij0, ij1, ij2, eu_, ev_);
                        if (ok) { iu = eu_; iv = ev_; }
                    }
                    const float* rr = pix + 4 * ((size_t)iv * g.img_w + iu);      /* unconditional, clamped */
                    PixRec rec; rec.z = rr[0]; rec.nx = rr[1]; rec.ny = rr[2]; rec.nz = rr[3];
                    float dn, wn;
                    if (verdict == UNIT_FRONT) {                                  /* colour mode: certified unit in the exact pass */
                        dn = -g.delta; wn = 1.0f; n_fast++;
                    } else {
                        float fx_, fy_, eb;
                        bool band;
                        backproject_px(kp, iu, iv, rec.z, fx_, fy_);
                        const bool upd = fuse_distance_flags(g, cx[v], cy[v], cz[v], fx_, fy_, rec, dn, eb, band) & ok;
                        if (!upd) continue;
                        wn = fuse_weight(band, eb);
                    }
                    fuse_apply(grid[o], grid[o + 1], dn, wn);
                    n_updated++;
                    if (color) {
                        float* c = color + 2 * o;                                 /* 4 floats per voxel */
                        const uint8_t* px = rgb3 + 3 * ((size_t)iv * g.img_w + iu);
                        const float wc = color_weight(wn, color_cosine(rec.nx, rec.ny, rec.nz));   /* K1 tabulates the cosine per pixel */
                        color_apply(c[0], c[1], c[2], c[3], wc, (int)px[0], (int)px[1], (int)px[2]);
                    }
                }
            }
        }
    }
    g_last_fast = n_fast;
    return n_updated;
}

/* K2, mirroring k_linearize's lanes serially: per pixel 13 samples, J, psi, flag; sums in double */
void emul_linearize(const GridParams* gp, const float* grid, const float* pix, const PoseState* pose,
                    float* J /* [P*6] */, float* psi /* [P] */, uint8_t* flag /* [P] */, double* sums /* [30] */) {
    const GridParams& g = *gp;
    double M[7][9];
    for (int q = 0; q < 9; q++) M[0][q] = pose->R[q];
    for (int q = 0; q < 6; q++) perturbed_rot(g, pose->R, q, M[1 + q]);
    const K1Params kp = k1_params(g.K);
    HostFetch f{grid, g.m};
    const int P = g.ni * g.nj;
    for (int q = 0; q < N_SLOTS; q++) sums[q] = 0.0;
    for (int p = 0; p < P; p++) {
        const int ii = p / g.nj, jj = p - ii * g.nj;
        const int u = ii * g.stride, v = jj * g.stride;
        const float z = pix[4 * ((size_t)v * g.img_w + u)];
        for (int a = 0; a < 6; a++) J[(size_t)p * 6 + a] = 0.0f;
        psi[p] = 0.0f;
        if (!(z == z)) { flag[p] = 0; continue; }
        float x, y;
        backproject_px(kp, u, v, z, x, y);
        float val[13]; bool ok[13]; bool oob = false;
        for (int s = 0; s < 13; s++) {
            double vx, vy, vz;
            sample_coords(g, M[(s < 7) ? 0 : (s - 6)], pose->t, s, (double)x, (double)y, (double)z, vx, vy, vz);
            if (s == 0) { const double dm = (double)g.m; oob = (vx < 0 || vy < 0 || vz < 0 || vx >= dm || vy >= dm || vz >= dm); }
            val[s] = interpolate_distance(vx, vy, vz, f, ok[s]);
        }
        bool allok = true;
        for (int s = 0; s < 13; s++) allok = allok && ok[s];
        const int fl = oob ? 2 : (allok ? 1 : 3);
        flag[p] = (uint8_t)fl;
        if (fl == 2) sums[SLOT_NOOB] += 1.0;
        if (fl != 1) continue;
        double xv[7];
        for (int a = 0; a < 6; a++) {
            const float step = (a == 0) ? g.v_h2_width : (a == 1) ? g.v_h2_height : (a == 2) ? g.v_h2_depth : g.two_w_h;
            const float Ja = (val[2 * a + 1] - val[2 * a + 2]) / step;
            J[(size_t)p * 6 + a] = Ja; xv[a] = (double)Ja;
        }
        psi[p] = val[0]; xv[6] = (double)val[0];
        int q = 0;
        for (int r = 0; r < 6; r++) for (int c = r; c < 6; c++) { sums[SLOT_A + q] += xv[r] * xv[c]; q++; }
        for (int r = 0; r < 6; r++) sums[SLOT_B + r] += xv[6] * xv[r];
        sums[SLOT_RES] += xv[6] * xv[6];
        sums[SLOT_NVALID] += 1.0;
    }
}

/* certificate statistics of one frame (no grid access): how the 128-voxel items of k_fuse_cert split
 * into verdict classes.  out: [0] items total, [1] items in FRONT rows, [2] items FRONT at item level,
 * [3] items SKIP at item level, [4] per-lane items (row+item UNKNOWN), [5] of those: every active unit UNKNOWN,
 * [6] of those: no unit UNKNOWN, [7] units UNKNOWN, [8] units FRONT (lane level), [9] units SKIP (lane level),
 * [10] per-lane items whose centre-pixel band test predicts "hopeless", [11] of those really all-UNKNOWN,
 * [12] units certified inside predicted-hopeless items (the cost of a wrong prediction) */
void emul_fuse_stats(const GridParams* gp, const float* pix, const PoseState* pose, float band_margin, int64_t out[16]) {
    const GridParams& g = *gp;
    const int m = g.m;
    const double* Ri = pose->Rinv; const double* ti = pose->tinv;
    const K1Params kp = k1_params(g.K);
    CertPyramid P;
    std::vector<std::vector<float>> zf(CERT_LEVELS), zb(CERT_LEVELS);
    for (int l = 0; l < CERT_LEVELS; l++) {
        P.w[l] = (g.img_w + (1 << l) - 1) >> l; P.h[l] = (g.img_h + (1 << l) - 1) >> l; P.off[l] = 0;
        zf[l].assign((size_t)P.w[l] * P.h[l], 3.402823466e+38f); zb[l].assign((size_t)P.w[l] * P.h[l], -3.402823466e+38f);
    }
    for (int v = 0; v < g.img_h; v++)
        for (int u = 0; u < g.img_w; u++) {
            const float* q = pix + 4 * ((size_t)v * g.img_w + u);
            PixRec r; r.z = q[0]; r.nx = q[1]; r.ny = q[2]; r.nz = q[3];
            cert_pixel(g, kp, u, v, r, zf[0][(size_t)v * P.w[0] + u], zb[0][(size_t)v * P.w[0] + u]);
        }
    for (int l = 1; l < CERT_LEVELS; l++)
        for (int y = 0; y < P.h[l - 1]; y++)
            for (int x = 0; x < P.w[l - 1]; x++) {
                const size_t o = (size_t)(y >> 1) * P.w[l] + (x >> 1), i = (size_t)y * P.w[l - 1] + x;
                zf[l][o] = fminf(zf[l][o], zf[l - 1][i]); zb[l][o] = fmaxf(zb[l][o], zb[l - 1][i]);
            }
    auto fetch = [&](int level, int x, int y, float& f, float& b) { f = zf[level][(size_t)y * P.w[level] + x]; b = zb[level][(size_t)y * P.w[level] + x]; };
    int64_t c[16] = {0};
#pragma omp parallel
    {
        int64_t lc[16] = {0};
#pragma omp for schedule(dynamic, 1)
        for (int k = g.ks0; k < g.ks1; k++) {
            const double gz = voxel_centre(g.vs_z, k, g.origin[2]);
            const double pz0 = Ri[2] * gz, pz1 = Ri[5] * gz, pz2 = Ri[8] * gz;
            for (int j = 0; j < m; j++) {
                const double gy = voxel_centre(g.vs_y, j, g.origin[1]);
                const double py0 = Ri[1] * gy, py1 = Ri[4] * gy, py2 = Ri[7] * gy;
                int ilo = 0, ihi = m;
                row_clip(g, Ri, ti, py0, py1, py2, pz0, pz1, pz2, ilo, ihi);
                if (ihi <= ilo) continue;
                auto cam = [&](int x, double& X, double& Y, double& Z) {
                    const double gx = voxel_centre(g.vs_x, x, g.origin[0]);
                    X = ((Ri[0] * gx + py0) + pz0) + ti[0]; Y = ((Ri[3] * gx + py1) + pz1) + ti[1]; Z = ((Ri[6] * gx + py2) + pz2) + ti[2];
                };
                double ax, ay, az, bx, by, bz;
                cam(ilo, ax, ay, az); cam(ihi - 1, bx, by, bz);
                const int rowv = unit_certificate(g, P, ax, ay, az, bx, by, bz, fetch);
                if (rowv == UNIT_SKIP) continue;
                for (int xs = ilo; xs < ihi; xs += 128) {
                    lc[0]++;
                    if (rowv == UNIT_FRONT) { lc[1]++; continue; }
                    const int xb = (xs + 127 < ihi - 1) ? xs + 127 : ihi - 1;
                    cam(xs, ax, ay, az); cam(xb, bx, by, bz);
                    const int itemv = unit_certificate(g, P, ax, ay, az, bx, by, bz, fetch);
                    if (itemv == UNIT_FRONT) { lc[2]++; continue; }
                    if (itemv == UNIT_SKIP) { lc[3]++; continue; }
                    lc[4]++;
                    int nu = 0, nf = 0, ns = 0;
                    for (int x0 = xs; x0 <= xb; x0 += 4) {
                        cam(x0, ax, ay, az); cam(x0 + 3, bx, by, bz);
                        const int v = unit_certificate(g, P, ax, ay, az, bx, by, bz, fetch);
                        if (v == UNIT_UNKNOWN) nu++; else if (v == UNIT_FRONT) nf++; else ns++;
                    }
                    lc[7] += nu; lc[8] += nf; lc[9] += ns;
                    if (nf + ns == 0) lc[5]++;
                    if (nu == 0) lc[6]++;
                    /* predictor: both ends and the middle of the item project to pixels whose band
                     * [zfree, zbehind] contains the voxel depth with a margin */
                    bool hopeless = true;
                    const int probes[3] = {xs, (xs + xb) >> 1, xb};
                    for (int q = 0; q < 3 && hopeless; q++) {
                        cam(probes[q], ax, ay, az);
                        if (!(az > 0.05)) { hopeless = false; break; }
                        const float uu = (float)(g.K[0] * ax / az + g.K[2]), vv = (float)(g.K[4] * ay / az + g.K[5]);
                        const int iu = (int)floorf(uu + 0.5f), iv = (int)floorf(vv + 0.5f);
                        if (iu < 1 || iv < 1 || iu > g.img_w - 2 || iv > g.img_h - 2) { hopeless = false; break; }
                        const float f = zf[0][(size_t)iv * P.w[0] + iu], b = zb[0][(size_t)iv * P.w[0] + iu];
                        if (!((float)az > f + band_margin && (float)az < b - band_margin)) hopeless = false;
                    }
                    if (hopeless) { lc[10]++; if (nf + ns == 0) lc[11]++; lc[12] += nf + ns; }
                }
            }
        }
#pragma omp critical
        for (int q = 0; q < 16; q++) c[q] += lc[q];
    }
    for (int q = 0; q < 16; q++) out[q] = c[q];
}

/* brick-level certificate statistics: bricks of bx*by*bz voxels.  out: [0] bricks, [1] FRONT, [2] SKIP, [3] UNKNOWN,
 * [4..6] units UNKNOWN/FRONT/SKIP inside UNKNOWN bricks, [7] units in FRONT bricks, [8] units in SKIP bricks that are in view
 * (unit verdict not SKIP-by-geometry), [9] certified-brick units whose own verdict CONTRADICTS the brick (must be 0) */
void emul_brick_stats(const GridParams* gp, const float* pix, const PoseState* pose, int bx, int by, int bz, int64_t out[16]) {
    const GridParams& g = *gp;
    const int m = g.m;
    const double* Ri = pose->Rinv; const double* ti = pose->tinv;
    const K1Params kp = k1_params(g.K);
    CertPyramid P;
    std::vector<std::vector<float>> zf(CERT_LEVELS), zb(CERT_LEVELS);
    for (int l = 0; l < CERT_LEVELS; l++) {
        P.w[l] = (g.img_w + (1 << l) - 1) >> l; P.h[l] = (g.img_h + (1 << l) - 1) >> l; P.off[l] = 0;
        zf[l].assign((size_t)P.w[l] * P.h[l], 3.402823466e+38f); zb[l].assign((size_t)P.w[l] * P.h[l], -3.402823466e+38f);
    }
    for (int v = 0; v < g.img_h; v++)
        for (int u = 0; u < g.img_w; u++) {
            const float* q = pix + 4 * ((size_t)v * g.img_w + u);
            PixRec r; r.z = q[0]; r.nx = q[1]; r.ny = q[2]; r.nz = q[3];
            cert_pixel(g, kp, u, v, r, zf[0][(size_t)v * P.w[0] + u], zb[0][(size_t)v * P.w[0] + u]);
        }
    for (int l = 1; l < CERT_LEVELS; l++)
        for (int y = 0; y < P.h[l - 1]; y++)
            for (int x = 0; x < P.w[l - 1]; x++) {
                const size_t o = (size_t)(y >> 1) * P.w[l] + (x >> 1), i = (size_t)y * P.w[l - 1] + x;
                zf[l][o] = fminf(zf[l][o], zf[l - 1][i]); zb[l][o] = fmaxf(zb[l][o], zb[l - 1][i]);
            }
    auto fetch = [&](int level, int x, int y, float& f, float& b) { f = zf[level][(size_t)y * P.w[level] + x]; b = zb[level][(size_t)y * P.w[level] + x]; };
    auto cam = [&](int x, int j, int k, double& X, double& Y, double& Z) {
        const double gx = voxel_centre(g.vs_x, x, g.origin[0]), gy = voxel_centre(g.vs_y, j, g.origin[1]), gz = voxel_centre(g.vs_z, k, g.origin[2]);
        X = ((Ri[0] * gx + Ri[1] * gy) + Ri[2] * gz) + ti[0]; Y = ((Ri[3] * gx + Ri[4] * gy) + Ri[5] * gz) + ti[1]; Z = ((Ri[6] * gx + Ri[7] * gy) + Ri[8] * gz) + ti[2];
    };
    int64_t c[16] = {0};
#pragma omp parallel
    {
        int64_t lc[16] = {0};
#pragma omp for schedule(dynamic, 1)
        for (int k0 = g.ks0; k0 < g.ks1; k0 += bz)
            for (int j0 = 0; j0 < m; j0 += by)
                for (int i0 = 0; i0 < m; i0 += bx) {
                    const int i1 = std::min(i0 + bx, m) - 1, j1 = std::min(j0 + by, m) - 1, k1 = std::min(k0 + bz, g.ks1) - 1;
                    float X[8], Y[8], Z[8];
                    for (int q = 0; q < 8; q++) {
                        double x, y, z;
                        cam((q & 1) ? i1 : i0, (q & 2) ? j1 : j0, (q & 4) ? k1 : k0, x, y, z);
                        X[q] = (float)x; Y[q] = (float)y; Z[q] = (float)z;
                    }
                    const int bv = box_certificate(g, P, X, Y, Z, fetch);
                    lc[0]++; lc[1 + (bv == UNIT_FRONT ? 0 : bv == UNIT_SKIP ? 1 : 2)]++;
                    for (int k = k0; k <= k1; k++)
                        for (int j = j0; j <= j1; j++)
                            for (int x0 = i0; x0 <= i1; x0 += 4) {
                                double ax, ay, az, bx_, by_, bz_;
                                cam(x0, j, k, ax, ay, az); cam(x0 + 3, j, k, bx_, by_, bz_);
                                const int v = unit_certificate(g, P, ax, ay, az, bx_, by_, bz_, fetch);
                                if (bv == UNIT_UNKNOWN) lc[4 + (v == UNIT_UNKNOWN ? 0 : v == UNIT_FRONT ? 1 : 2)]++;
                                else if (bv == UNIT_FRONT) { lc[7]++; if (v == UNIT_SKIP) lc[9]++; }
                                else { lc[8]++; if (v == UNIT_FRONT) lc[9]++; }
                            }
                }
#pragma omp critical
        for (int q = 0; q < 16; q++) c[q] += lc[q];
    }
    for (int q = 0; q < 16; q++) out[q] = c[q];
}

/* how many UNKNOWN units would per-VOXEL certificates resolve?  out: [0] units UNKNOWN at unit level, [1] of those: all four
 * voxels individually certified SKIP, [2] all four individually certified (skip or front), [3] all four FRONT */
void emul_voxel_cert_stats(const GridParams* gp, const float* pix, const PoseState* pose, int64_t out[8]) {
    const GridParams& g = *gp;
    const int m = g.m;
    const double* Ri = pose->Rinv; const double* ti = pose->tinv;
    const K1Params kp = k1_params(g.K);
    CertPyramid P;
    std::vector<std::vector<float>> zf(CERT_LEVELS), zb(CERT_LEVELS);
    for (int l = 0; l < CERT_LEVELS; l++) {
        P.w[l] = (g.img_w + (1 << l) - 1) >> l; P.h[l] = (g.img_h + (1 << l) - 1) >> l; P.off[l] = 0;
        zf[l].assign((size_t)P.w[l] * P.h[l], 3.402823466e+38f); zb[l].assign((size_t)P.w[l] * P.h[l], -3.402823466e+38f);
    }
    for (int v = 0; v < g.img_h; v++)
        for (int u = 0; u < g.img_w; u++) {
            const float* q = pix + 4 * ((size_t)v * g.img_w + u);
            PixRec r; r.z = q[0]; r.nx = q[1]; r.ny = q[2]; r.nz = q[3];
            cert_pixel(g, kp, u, v, r, zf[0][(size_t)v * P.w[0] + u], zb[0][(size_t)v * P.w[0] + u]);
        }
    for (int l = 1; l < CERT_LEVELS; l++)
        for (int y = 0; y < P.h[l - 1]; y++)
            for (int x = 0; x < P.w[l - 1]; x++) {
                const size_t o = (size_t)(y >> 1) * P.w[l] + (x >> 1), i = (size_t)y * P.w[l - 1] + x;
                zf[l][o] = fminf(zf[l][o], zf[l - 1][i]); zb[l][o] = fmaxf(zb[l][o], zb[l - 1][i]);
            }
    auto fetch = [&](int level, int x, int y, float& f, float& b) { f = zf[level][(size_t)y * P.w[level] + x]; b = zb[level][(size_t)y * P.w[level] + x]; };
    int64_t c[8] = {0};
#pragma omp parallel
    {
        int64_t lc[8] = {0};
#pragma omp for schedule(dynamic, 1)
        for (int k = g.ks0; k < g.ks1; k++) {
            const double gz = voxel_centre(g.vs_z, k, g.origin[2]);
            const double pz0 = Ri[2] * gz, pz1 = Ri[5] * gz, pz2 = Ri[8] * gz;
            for (int j = 0; j < m; j++) {
                const double gy = voxel_centre(g.vs_y, j, g.origin[1]);
                const double py0 = Ri[1] * gy, py1 = Ri[4] * gy, py2 = Ri[7] * gy;
                int ilo = 0, ihi = m;
                row_clip(g, Ri, ti, py0, py1, py2, pz0, pz1, pz2, ilo, ihi);
                auto cam = [&](int x, double& X, double& Y, double& Z) {
                    const double gx = voxel_centre(g.vs_x, x, g.origin[0]);
                    X = ((Ri[0] * gx + py0) + pz0) + ti[0]; Y = ((Ri[3] * gx + py1) + pz1) + ti[1]; Z = ((Ri[6] * gx + py2) + pz2) + ti[2];
                };
                for (int x0 = ilo; x0 < ihi; x0 += 4) {
                    double ax, ay, az, bx, by, bz;
                    cam(x0, ax, ay, az); cam(x0 + 3, bx, by, bz);
                    if (unit_certificate(g, P, ax, ay, az, bx, by, bz, fetch) != UNIT_UNKNOWN) continue;
                    lc[0]++;
                    int ns = 0, nf = 0;
                    for (int v = 0; v < 4; v++) {
                        cam(x0 + v, ax, ay, az);
                        const int vv = unit_certificate(g, P, ax, ay, az, ax, ay, az, fetch);
                        ns += vv == UNIT_SKIP; nf += vv == UNIT_FRONT;
                    }
                    if (ns == 4) lc[1]++;
                    if (ns + nf == 4) lc[2]++;
                    if (nf == 4) lc[3]++;
                }
            }
        }
#pragma omp critical
        for (int q = 0; q < 8; q++) c[q] += lc[q];
    }
    for (int q = 0; q < 8; q++) out[q] = c[q];
}

/* SDF::interpolate_color through the kernels' core (k_sample_color) */
void emul_interpolate_color(const GridParams* gp, const float* color, int64_t n, const double* gpts, float* rgba) {
    const GridParams& g = *gp;
    const int m = g.m;
    auto fetch = [&](int ci, int cj, int ck, float& cw, float& r, float& gg, float& b) {
        if ((unsigned)ci >= (unsigned)m || (unsigned)cj >= (unsigned)m || (unsigned)ck >= (unsigned)m) return false;
        const float* c = color + 4 * (((size_t)ck * m + cj) * m + ci);
        cw = c[0]; r = c[1]; gg = c[2]; b = c[3];
        return true;
    };
    for (int64_t q = 0; q < n; q++) {
        const double vx = ((gpts[3 * q] - g.origin[0]) * (double)g.m_div_width - 0.5);
        const double vy = ((gpts[3 * q + 1] - g.origin[1]) * (double)g.m_div_height - 0.5);
        const double vz = ((gpts[3 * q + 2] - g.origin[2]) * (double)g.m_div_depth - 0.5);
        interpolate_color(vx, vy, vz, fetch, rgba + 4 * q);
    }
}

/* the mesher through the kernels' core (k_mc_sweep): cells in (i,j,k) order, vertices via mc_edge_vertex.
 * First call with xyz = NULL to get the vertex count. */
int64_t emul_mesh(const GridParams* gp, float width, float height, float depth, float iso, const float* grid, float* xyz) {
    const GridParams& g = *gp;
    const int m = g.m;
    if (!(iso >= 0.0f && iso < 1.0f)) return 0;
    McParams P; P.width = width; P.height = height; P.depth = depth; P.iso = iso;
    const float fm = (float)m;
    int64_t n = 0;
    auto at = [&](int i, int j, int k) { return grid + 2 * (((size_t)k * m + j) * m + i); };
    for (int i = 1; i <= m - 2; i++)
        for (int j = 1; j <= m - 2; j++)
            for (int k = 1; k <= m - 2; k++) {
                const float* c[8] = {at(i, j, k), at(i + 1, j, k), at(i + 1, j, k + 1), at(i, j, k + 1),
                                     at(i, j + 1, k), at(i + 1, j + 1, k), at(i + 1, j + 1, k + 1), at(i, j + 1, k + 1)};
                float d[8], w[8];
                for (int q = 0; q < 8; q++) { d[q] = c[q][0]; w[q] = c[q][1]; }
                const int ci = mc_cube_index(d, w, P.iso);
                if (ci == 0 || ci == 255) continue;
                const unsigned long long row = c_mc_tri[ci];
                const int nv = mc_vertex_count(row);
                if (xyz)
                    for (int q = 0; q < nv; q++) mc_edge_vertex(P, fm, i, j, k, (int)((row >> (4 * q)) & 0xFull), d, xyz + 3 * (n + q));
                n += nv;
            }
    return n;
}

void emul_gn_update(const GridParams* g, PoseState* pose, const double* sums) { gn_update(*g, *pose, sums); }
void emul_pose_stats(const PoseState* p, int32_t out[4], double twist[6]) {
    out[0] = p->iterations; out[1] = p->stopped; out[2] = p->singular; out[3] = p->halo_miss;
    memcpy(twist, p->twist, sizeof p->twist);
}
void emul_exp_map(const double twist[6], double R[9], double t[3]) { exp_map(twist, R, t); }
double emul_weight_exp(double x) { return weight_exp(x); }
int emul_trunc_f2i(float v) { return trunc_f2i(v); }

/* k_linearize's work distribution (tsdf_core.cuh: lin_layout / lin_pixel_of), walked exactly as the kernel's blocks do:
 * count[ii * nj + jj] += 1 for every (block, sweep, slot) that yields a pixel.  Returns the number of sweeps. */
int emul_lin_coverage(int ni, int nj, int nblocks, int mt_sweep, int sharded, int32_t* count) {
    const LinLayout L = lin_layout(ni, nj, nblocks, mt_sweep);
    for (int b = 0; b < nblocks; b++)
        for (int q = 0; q < L.n_sweeps; q++)
            for (int tl = 0; tl < 16 * mt_sweep; tl++) {
                int ii, jj;
                if (lin_pixel_of(L, ni, nj, nblocks, b, mt_sweep, sharded != 0, q, tl, ii, jj)) count[ii * nj + jj]++;
            }
    return L.n_sweeps;
}

}  // extern "C"
