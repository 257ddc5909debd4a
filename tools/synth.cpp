/*
 * synth.cpp — deterministic synthetic depth frames for tests and bench (host tool, double
 * precision, not part of the product path and not part of the oracle).
 *
 * The TUM RGB-D bags the reference was run on are not available offline, so depth frames
 * are raycast from an analytic scene along the fr1/plant ground-truth camera path
 * (BASELINE.json north_star).  Scene: a closed room whose walls sit 0.5 m
 * inside the reference's default volume ([-3,3]x[-3,3]x[-0.5,3], sdf_reconstruction.cpp:83-85),
 * so every back-projected point stays inside the volume (SURVEY.md TRAP 5), plus a "plant on
 * a pedestal" at the point the trajectory's optical axes converge on, (0.32,-0.90,1.05), and
 * some furniture for 6-DoF observability.
 *
 * Depth convention: metres along the camera z axis (not range), float32; every pixel hits
 * something, so there are no invalid pixels unless the caller injects them.
 */
#include <cmath>
#include <cstdint>
#include <limits>

namespace {

struct Sphere { double c[3], r; };
struct Box { double lo[3], hi[3]; };

const double ROOM_LO[3] = {-2.4, -2.4, 0.0};
const double ROOM_HI[3] = {2.4, 2.4, 2.5};

const Sphere SPHERES[] = {
    {{0.32, -0.90, 0.80}, 0.22},   // pot
    {{0.32, -0.90, 1.15}, 0.18},   // plant
    {{-1.20, 1.40, 0.45}, 0.45},
    {{1.90, -1.90, 1.60}, 0.35},
    {{-1.70, -0.60, 1.90}, 0.30},
};
const Box BOXES[] = {
    {{0.17, -1.05, 0.0}, {0.47, -0.75, 0.62}},    // pedestal
    {{-2.4, -2.4, 0.0}, {-1.6, -1.2, 1.1}},
    {{1.7, 0.5, 0.0}, {2.4, 1.9, 0.8}},
    {{-0.8, 1.9, 0.0}, {0.6, 2.4, 1.8}},
    {{1.3, -2.4, 0.0}, {2.4, -2.0, 2.5}},
};
const int N_SPHERES = sizeof(SPHERES) / sizeof(SPHERES[0]);
const int N_BOXES = sizeof(BOXES) / sizeof(BOXES[0]);

inline double hit_room(const double o[3], const double d[3]) {
    double best = std::numeric_limits<double>::infinity();
    for (int a = 0; a < 3; a++) {
        if (d[a] > 0) { double l = (ROOM_HI[a] - o[a]) / d[a]; if (l < best) best = l; }
        else if (d[a] < 0) { double l = (ROOM_LO[a] - o[a]) / d[a]; if (l < best) best = l; }
    }
    return best;
}
inline double hit_sphere(const Sphere& s, const double o[3], const double d[3]) {
    double oc[3] = {o[0] - s.c[0], o[1] - s.c[1], o[2] - s.c[2]};
    double a = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
    double b = oc[0] * d[0] + oc[1] * d[1] + oc[2] * d[2];
    double c = oc[0] * oc[0] + oc[1] * oc[1] + oc[2] * oc[2] - s.r * s.r;
    double disc = b * b - a * c;
    if (disc < 0) return std::numeric_limits<double>::infinity();
    double l = (-b - std::sqrt(disc)) / a;
    return l > 0 ? l : std::numeric_limits<double>::infinity();
}
inline double hit_box(const Box& bx, const double o[3], const double d[3]) {
    double tmin = 0, tmax = std::numeric_limits<double>::infinity();
    for (int a = 0; a < 3; a++) {
        if (d[a] == 0) {
            if (o[a] < bx.lo[a] || o[a] > bx.hi[a]) return std::numeric_limits<double>::infinity();
        } else {
            double t1 = (bx.lo[a] - o[a]) / d[a], t2 = (bx.hi[a] - o[a]) / d[a];
            if (t1 > t2) { double t = t1; t1 = t2; t2 = t; }
            if (t1 > tmin) tmin = t1;
            if (t2 < tmax) tmax = t2;
            if (tmin > tmax) return std::numeric_limits<double>::infinity();
        }
    }
    return tmin > 0 ? tmin : std::numeric_limits<double>::infinity();
}

}  // namespace

extern "C" {

/* R: camera->world rotation (row-major), t: camera centre in world, K row-major 3x3 */
void synth_render_depth(const double R[9], const double t[3], const double K[9], int w, int h, float* depth) {
    const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
#pragma omp parallel for schedule(static)
    for (int v = 0; v < h; v++) {
        for (int u = 0; u < w; u++) {
            double dc[3] = {(u - cx) / fx, (v - cy) / fy, 1.0};
            double d[3] = {R[0] * dc[0] + R[1] * dc[1] + R[2] * dc[2],
                           R[3] * dc[0] + R[4] * dc[1] + R[5] * dc[2],
                           R[6] * dc[0] + R[7] * dc[1] + R[8] * dc[2]};
            double best = hit_room(t, d);
            for (int s = 0; s < N_SPHERES; s++) { double l = hit_sphere(SPHERES[s], t, d); if (l < best) best = l; }
            for (int b = 0; b < N_BOXES; b++) { double l = hit_box(BOXES[b], t, d); if (l < best) best = l; }
            depth[(size_t)v * w + u] = (float)best;   /* camera-z of the hit, since dc.z == 1 */
        }
    }
}

/* distance from a point to the nearest object surface (not the room): used to check that
 * the camera path keeps clear of the furniture */
double synth_clearance(const double p[3]) {
    double best = std::numeric_limits<double>::infinity();
    for (int s = 0; s < N_SPHERES; s++) {
        double dx = p[0] - SPHERES[s].c[0], dy = p[1] - SPHERES[s].c[1], dz = p[2] - SPHERES[s].c[2];
        double d = std::sqrt(dx * dx + dy * dy + dz * dz) - SPHERES[s].r;
        if (d < best) best = d;
    }
    for (int b = 0; b < N_BOXES; b++) {
        double q[3], out2 = 0, in = -std::numeric_limits<double>::infinity();
        for (int a = 0; a < 3; a++) {
            double lo = BOXES[b].lo[a] - p[a], hi = p[a] - BOXES[b].hi[a];
            q[a] = lo > hi ? lo : hi;
            if (q[a] > 0) out2 += q[a] * q[a];
            if (q[a] > in) in = q[a];
        }
        double d = out2 > 0 ? std::sqrt(out2) : in;
        if (d < best) best = d;
    }
    for (int a = 0; a < 3; a++) {
        double d1 = p[a] - ROOM_LO[a], d2 = ROOM_HI[a] - p[a];
        if (d1 < best) best = d1;
        if (d2 < best) best = d2;
    }
    return best;
}

}  // extern "C"
