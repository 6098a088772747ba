// Host build of the product's __host__ __device__ math headers (exact_math.cuh, geometry3d.cuh)
// so their arithmetic can be checked against the oracle on a machine without a GPU.
// Test infrastructure only; built by tests/test_host_math.py with g++ -ffp-contract=off -mfma.
#include <cstring>
#include "../nirrt_star_b200/csrc/exact_math.cuh"
#include "../nirrt_star_b200/csrc/glibc_trig.cuh"
#include "../nirrt_star_b200/csrc/geometry3d.cuh"
#include "../nirrt_star_b200/csrc/geometry2d.cuh"

using namespace nirrt;

extern "C" {
double hh_hypot3(double a, double b, double c) { return hypot3(a, b, c); }
double hh_hypot2(double a, double b) { return hypot2(a, b); }
double hh_rownorm3(double a, double b, double c) { return rownorm3(a, b, c); }
double hh_vecnorm3(double a, double b, double c) { return vecnorm3(a, b, c); }
void hh_cr_sincos(double x, double *s, double *c) { cr_sincos(x, s, c); }
// vectorised: out_s[i] = glibc_sin(x[i]), out_c[i] = glibc_cos(x[i])
void hh_glibc_sincos(long n, const double *x, double *out_s, double *out_c) {
    for (long i = 0; i < n; i++) { out_s[i] = glibc_sin(x[i]); out_c[i] = glibc_cos(x[i]); }
}
double hh_sqrt_le_threshold(double r) { return sqrt_le_threshold(r); }
double hh_pairwise_sum(const double *a, long n) { return pairwise_sum(a, n); }

static Geom3 g;
void hh_set_geom(int nb, const double *balls, const double *r2, int nx, const double *boxes, double cl, const double *range6) {
    memset(&g, 0, sizeof(g));
    g.n_balls = nb; g.n_boxes = nx; g.clearance = cl;
    memcpy(g.range, range6, 48);
    for (int k = 0; k < nb; k++) { memcpy(g.balls[k], balls + 4 * k, 32); g.ball_r2[k] = r2[k]; }
    for (int k = 0; k < nx; k++) memcpy(g.boxes[k], boxes + 6 * k, 48);
}
void hh_collide(long m, const double *edges, unsigned char *out) {
    for (long i = 0; i < m; i++) out[i] = seg_collides(g, edges + 6 * i, edges + 6 * i + 3);
}
void hh_inside(long m, const double *p, unsigned char *out) { for (long i = 0; i < m; i++) out[i] = point_inside_obs(g, p + 3 * i); }
void hh_valid(long m, const double *p, unsigned char *out) { for (long i = 0; i < m; i++) out[i] = point_valid(g, p + 3 * i); }

// ---- 2D
double hh_np_hypot(double a, double b) { return np_hypot(a, b); }
double hh_cr_atan2(double y, double x) { return cr_atan2(y, x); }
void hh_glibc_atan2(long n, const double *y, const double *x, double *out) { for (long i = 0; i < n; i++) out[i] = glibc_atan2(y[i], x[i]); }
double hh_vecnorm2(double a, double b) { return vecnorm2(a, b); }
double hh_rownorm2(double a, double b) { return rownorm2(a, b); }
static Geom2 g2;
void hh_set_geom2(int nc, const double *circles, int nr, const double *rects, double cl, const double *range4) {
    memset(&g2, 0, sizeof(g2));
    g2.n_circles = nc; g2.n_rects = nr; g2.clearance = cl;
    memcpy(g2.range, range4, 32);
    for (int k = 0; k < nc; k++) memcpy(g2.circles[k], circles + 3 * k, 24);
    for (int k = 0; k < nr; k++) memcpy(g2.rects[k], rects + 4 * k, 32);
}
void hh_collide2(long m, const double *edges, unsigned char *out) {
    for (long i = 0; i < m; i++) out[i] = seg_collides(g2, edges + 4 * i, edges + 4 * i + 2);
}
void hh_inside2(long m, const double *p, unsigned char *out) { for (long i = 0; i < m; i++) out[i] = point_inside_obs(g2, p + 2 * i); }
void hh_valid2(long m, const double *p, unsigned char *out) { for (long i = 0; i < m; i++) out[i] = point_valid(g2, p + 2 * i); }
}
