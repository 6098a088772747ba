// Host build of the product's __host__ __device__ math headers (exact_math.cuh, geometry3d.cuh)
// so their arithmetic can be checked against the oracle on a machine without a GPU.
// Test infrastructure only; built by tests/test_host_math.py with g++ -ffp-contract=off -mfma.
#include <cstring>
#include "../nirrt_star_b200/csrc/exact_math.cuh"
#include "../nirrt_star_b200/csrc/geometry3d.cuh"

using namespace nirrt;

extern "C" {
double hh_hypot3(double a, double b, double c) { return hypot3(a, b, c); }
double hh_hypot2(double a, double b) { return hypot2(a, b); }
double hh_rownorm3(double a, double b, double c) { return rownorm3(a, b, c); }
double hh_vecnorm3(double a, double b, double c) { return vecnorm3(a, b, c); }
void hh_cr_sincos(double x, double *s, double *c) { cr_sincos(x, s, c); }
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
}
