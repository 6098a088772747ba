// planner3d.cu -- lock-step batched RRT*/IRRT*/NIRRT* loop body for 3D worlds on sm_100a, and the
// C ABI over it (include/nirrt_b200.h).
//
// One "iteration" advances every running planning problem (env) by one loop body of the reference
// (rrt_star_3d.py:37-55 == irrt_star_3d.py:50-71) with TWO kernels per group of problems:
//
//   k_nearest_m  (chunks x E)  the one HBM-bound pass: Nearest filter over the 2-byte fixed-point mirror
//                              of the coordinates + speculative collection of the Near ball around x_rand
//   k_expand     1 CTA / env   steer_body (exact Nearest finish, Steer, steer-edge collision, insert),
//                              exact Near test + sort + edge collision filter, cost walks, ChooseParent,
//                              Rewire, goal bookkeeping / records, top_body (driver phase machine, c_best
//                              refresh for IRRT*, MT19937 sampling of the NEXT iteration)
//
// (k_top opens a run; k_steer / k_nearest / k_near are the unfused f64 kernels behind NIRRT_SCAN=f64, the
// stand-alone query entry points and the per-kernel attribution of nirrt_batch_run_profiled_sync.)
//
// HBM layout per env (flat, capacity-strided):
//   vx[], vy[], vz[]   f64 SoA       -- source of truth of the coordinates (exact re-checks)
//   ux[], uy[], uz[]   u16 SoA       -- fixed-point mirror, what the scan streams (6 B / vertex / iteration)
//   nodes[]            {x,y,z,parent} 32 B records -- coordinates of candidates, what read_trees returns
//   links[]            {math.hypot(v - parent), parent, Near stamp} 16 B -- what a cost walk reads
//   hints[]            8 ancestor indices, 32 B -- eight hops of a cost walk per round trip
// All index-deciding arithmetic is IEEE-exact float64 in the reference's operand order
// (exact_math.cuh, geometry3d.cuh); the mirror, the hints and the stamps only select what is evaluated.
// There is no CPU fallback anywhere in this file.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <limits.h>
#include <map>
#include <mutex>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../include/nirrt_b200.h"
#include "exact_math.cuh"
#include "glibc_trig.cuh"
#include "geometry3d.cuh"
#include "geometry2d.cuh"
#include "mt19937.cuh"

using namespace nirrt;

// ------------------------------------------------------------------------------------------------
// error plumbing
#include "errors.h"
#include "devmem.h"
static thread_local std::string g_err;
int nirrt_set_error(int code, const std::string &msg) { g_err = msg; return code; }
static int fail(int code, const std::string &msg) { return nirrt_set_error(code, msg); }
#define CUDA_TRY(expr)                                                                                \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess)                                                                        \
            return fail(NIRRT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));          \
    } while (0)

extern "C" const char *nirrt_last_error(void) { return g_err.c_str(); }
extern "C" int nirrt_version(void) { return 100; }
extern "C" int nirrt_device_count(void) {
    // answered once per process: the entry points call this on every invocation, and a full property query costs
    // milliseconds per device
    static int cached = -1;
    if (cached >= 0) return cached;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    int ok = 0;
    for (int i = 0; i < n; i++) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, i) == cudaSuccess && major == 10) ok++;
    }
    cached = ok;
    return ok;
}

// ------------------------------------------------------------------------------------------------
// device memory cache (devmem.h)
namespace {
struct DevCache {
    std::mutex mu;
    std::unordered_map<void *, std::pair<size_t, int>> live;                   // block -> (bytes, device)
    std::multimap<std::pair<int, size_t>, void *> idle;                         // (device, bytes) -> block
    size_t idle_bytes = 0, cap = (size_t)32 << 30;
    DevCache() {
        const char *e = getenv("NIRRT_CACHE_GB");                               // 0 disables the cache
        if (e) cap = (size_t)(atof(e) * 1073741824.0);
    }
    void drop_idle() {                                                          // mu held
        for (auto &kv : idle) cudaFree(kv.second);
        idle.clear(); idle_bytes = 0;
    }
};
DevCache &dev_cache() { static DevCache c; return c; }
}  // namespace

cudaError_t nirrt_dev_malloc(void **p, size_t bytes) {
    DevCache &c = dev_cache();
    bytes = (bytes + 255) & ~(size_t)255;
    if (bytes == 0) bytes = 256;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    {
        std::lock_guard<std::mutex> g(c.mu);
        auto it = c.idle.find({dev, bytes});
        if (it != c.idle.end()) {
            *p = it->second;
            c.idle.erase(it);
            c.idle_bytes -= bytes;
            c.live[*p] = {bytes, dev};
        } else *p = nullptr;
    }
    if (*p) {                                       // a recycled block looks like a fresh one: zero-filled
        e = cudaMemset(*p, 0, bytes);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        return e;
    }
    e = cudaMalloc(p, bytes);
    if (e == cudaErrorMemoryAllocation) {           // give the cached blocks back and try once more
        cudaGetLastError();
        { std::lock_guard<std::mutex> g(c.mu); c.drop_idle(); }
        e = cudaMalloc(p, bytes);
    }
    if (e == cudaSuccess) { std::lock_guard<std::mutex> g(c.mu); c.live[*p] = {bytes, dev}; }
    return e;
}

void nirrt_dev_free(void *p) {
    if (!p) return;
    DevCache &c = dev_cache();
    cudaDeviceSynchronize();                        // cudaFree's contract: nothing in flight uses the block any more
    std::lock_guard<std::mutex> g(c.mu);
    auto it = c.live.find(p);
    if (it == c.live.end()) { cudaFree(p); return; }
    const size_t bytes = it->second.first;
    const int dev = it->second.second;
    c.live.erase(it);
    if (c.idle_bytes + bytes > c.cap) { cudaFree(p); return; }
    c.idle.insert({{dev, bytes}, p});
    c.idle_bytes += bytes;
}

extern "C" int nirrt_release_cached_memory(void) {
    DevCache &c = dev_cache();
    std::lock_guard<std::mutex> g(c.mu);
    c.drop_idle();
    return NIRRT_OK;
}

// ------------------------------------------------------------------------------------------------
// device-side state

struct __align__(32) Node {
    double x, y, z;
    long long parent;
};

enum { ST_DONE = 0, ST_PHASE1 = 1, ST_PHASE2 = 2, ST_WAIT_CLOUD = 3 };
enum { ERR_NEAR_OVERFLOW = 1, ERR_SOL_OVERFLOW = 2, ERR_VERTEX_OVERFLOW = 4, ERR_EMPTY_CLOUD = 8,
       ERR_PATH_DEPTH = 16, ERR_RECORD_OVERFLOW = 32, ERR_GOAL_OVERFLOW = 64, ERR_OUT_OF_RANGE = 128, ERR_CHILD_LISTS = 256 };

// what a mirror-scan CTA needs before it can start streaming (written by k_top / k_steer)
struct __align__(16) ScanHdr {
    int go, n;
    float qx, qy, qz;   // the query in the mirror's coordinates
    float thr;          // Near: squared-distance threshold
    float band;         // Nearest: 2 * margin
    int base;           // Nearest: value of the iteration copy's (never reset) spec_cnt when the sample was drawn;
                        // list position of a speculative candidate = atomicAdd(spec_cnt) - base
};

// What one iteration hands from Steer to the expansion (and to the trace readers).  Two copies per problem:
// in the pipelined RRT* driver the next iteration's scan + Steer + sample run while this iteration's
// expansion still reads its copy (View::par selects the copy, 0 everywhere else).
struct __align__(16) IterScratch {
    ScanHdr hdr1;       // Near scan header (query x_new) for the fallback scan
    int go, skip, nearest, new_idx, inserted, near_cnt;
    int spec_cnt;     // speculative Near candidates appended by the Nearest scans so far (never reset)
    int spec_base;    // its value when this iteration's list started: the list holds spec_cnt - spec_base entries
    int fb_cnt;       // matches of the fallback Near scan (k_expand), its own counter: spec_cnt is read by the next sample
    int use_spec;     // steer: x_new == x_rand (up to rounding), the speculative list is the Near superset
    int need_scan;    // steer: x_new != x_rand, k_expand runs the Near scan itself before filtering
    double x_new[3];
    double r, T_near;
    double cnew_default;  // cost(new) if ChooseParent keeps the steer parent: the walk from x_new, leaf -> root
    float near_thr;   // mirror scan: a <= near_thr selects the Near candidates that get the exact f64 test
    int pad1;
};

// Driver parameters of the current run.  They live in the per-problem control block (written by k_set_budget at
// the start of every nirrt_batch_run), NOT in the by-value View: changing them never invalidates a captured
// CUDA graph (see ensure_graph).
struct RunCfg {
    int iter_max, iter_after;
    int n_limit;       // > 0: problems whose tree reached n_limit vertices idle (benchmark pre-growth)
    int pad;
    double stop_below; // phase 1 ends when the recorded value drops below this (+inf: planning_random; finite: planning_block_gap)
    double pc_rate, pc_ratio;
};

struct __align__(16) EnvCtl {
    ScanHdr hdr0;       // Nearest scan header (query x_rand)
    IterScratch s[2];
    RunCfg cfg;
    // problem
    double start[3], goal[3];
    double step_len, search_radius;
    double c_min, center[3], C[9];
    double T_goal;  // rownorm(goal - v) <= step_len  <=>  sq <= T_goal
    // tree
    int n;
    // driver
    int state, saved_state, p1_done, left, budget, n_rec, resumed;
    // iteration scratch
    int cand_cnt;
    int scan_min;     // SAD scan: smallest L1 cell distance any chunk of this problem has seen so far in this iteration
    double x_rand[3];
    double curr_cost, c_best, c_update;
    double margin;    // bound on |mirror distance - f64 distance| for vertices inside the world range, in the
                      // mirror's units (f32 mirror: world units; u16 mirror: grid cells)
    double qlo[3], qscale;  // u16 mirror: cell = rint((x - qlo[d]) * qscale), 65535 cells over the longest world edge
    int fallbacks;    // Nearest scans whose in-band candidate list overflowed (k_steer re-scanned the env in f64)
    // goal bookkeeping
    int n_sol, n_goal, tree_changed, n_pc;
    long long last_gp;
    double last_len;
    int best_k, pad_k;   // RRT* eval driver: slot of the current goal parent in the goal-candidate list (-1: none)
    double best_val;     // its cost(vertex) + goal distance
    int err;
    unsigned stamp;   // k_expand invocations so far: tag of the Near stamps in the walk records
    // work counters (nirrt_batch_work_stats_sync): expansions, sum |Near|, sum candidates, goal-tracking traversal
    // rounds, full goal evaluations, refreshed goal candidates, re-parented seeds, sum of list length at full evaluations
    unsigned long long work[8];
};

struct View {
    int E, cap, stride, chunks, near_cap, rec_cap, sol_cap, pc_cap, path_cap;
    int env0;        // first problem of the group an iteration kernel works on (see nirrt_batch_run)
    int fuse_top;    // k_expand ends with the next iteration's k_top work (driver, c_best refresh, sampling)
    int fuse_steer;  // k_expand starts with k_steer's work (warp 0): one kernel for everything between two scans
    int par;         // which IterScratch copy / cand2 half this launch works on (iteration parity when pipelined, else 0)
    int pipe;        // 1: pipelined RRT* driver (k_front does Steer + accounting + next sample, k_expand only expands)
    int variant, mode;   // planner family / loop driver: select code paths, part of the graph cache key
    double *vx, *vy, *vz;
    float *fx, *fy, *fz;   // f32 mirror of the coordinates (scan layout, 4 B per coordinate); null = not in use
    unsigned short *ux, *uy, *uz;   // u16 fixed-point mirror (scan layout, 2 B per coordinate); null = not in use
    int tma;                        // u16 mirror: Nearest pass by the TMA-staged kernel (k_nearest_t); 0: LDG kernel (NIRRT_SCAN=u16ldg)
    int sad;                        // u8 mirror: Nearest pass by the SAD kernel (k_nearest_s, L1 filter); 0: DP4A kernel (NIRRT_SCAN=u8)
    unsigned *m8;                   // u8 fixed-point mirror: one word {x8, y8, z8, 0} per vertex (NIRRT_SCAN=u8); null = not in use
    Node *nodes;
    struct Hint *hints;  // [E][stride] ancestor hints of the cost walks (see walk_to_root)
    struct Link *links;  // [E][stride] {edge length to the parent, parent}: what a cost walk reads
    Geom3 *geom;
    Geom2 *geom2;    // 2D worlds (dim == 2)
    int dim;
    MtState *mt;
    MtState *mt_py;  // CPython `random` stream (2D informed sampling, irrt_star_2d.py:146-151)
    EnvCtl *ctl;
    double *part_s;
    int *part_i;
    int *cand;       // [E][near_cap] unordered Nearest candidates (k_nearest_m -> k_steer), then Near candidates of a fallback scan
    int *cand2;      // [E][near_cap] speculative Near candidates collected during the Nearest scan
    int *near_out;   // [E][near_cap] final (ordered, collision-filtered) Near list: trace
    unsigned char *big;  // [E][near_cap * kNearRowAlloc] HBM staging of k_expand for more than kNearSmem candidates; null if near_cap <= kNearSmem
    int *sol;        // [E][sol_cap]  path_solutions
    int *gc_idx;     // [E][cap]      vertices within step_len of the goal (RRT* eval driver)
    double *gc_d;    // [E][cap]      their goal distance, +inf when the goal edge collides
    double *gc_cost; // [E][gc_stride] cached cost(vertex) + goal distance (+inf when the goal edge collides), see goal_track
    int gc_stride;   // row length of gc_d / gc_cost: max(cap, sol_cap) -- the informed family keeps its candidates in `sol`
    struct Kid *kid; // [E][stride]   child lists + goal-candidate slot of every vertex (RRT* eval driver), see goal_track
    double *records; // [E][rec_cap]
    const unsigned char *occ;   // 2D guidance clouds: free-space masks [E][occ_h][occ_w], 1 = free (binary_mask), null until set
    int occ_h, occ_w;
    double *pc;      // [E][pc_cap][3]
    double *pathseg; // [E][path_cap]
    const double *near_table;
};

__host__ __device__ __forceinline__ bool has_mirror(const View &v) { return v.ux || v.fx || v.m8; }
__host__ __device__ __forceinline__ int scan_bytes_per_vertex(const View &v) { return v.m8 ? 4 : (v.ux ? 2 : (v.fx ? 4 : 8)) * v.dim; }
#define IT (c->s[v.par])     // this launch's IterScratch copy (every user names its EnvCtl pointer c and its View v)
__device__ __forceinline__ int *cand2_of(const View &v, int e) { return v.cand2 + ((size_t)v.par * v.E + e) * v.near_cap; }

// planner families: 0 RRT*, 1 IRRT*, 2 NIRRT* (informed + guidance cloud with updates),
// 3 NRRT* (RRT* driver + a fixed guidance cloud, nrrt_star_png_3d.py:52-56)
__host__ __device__ __forceinline__ bool fam_informed(int variant) { return variant == 1 || variant == 2; }
__host__ __device__ __forceinline__ bool fam_cloud(int variant) { return variant == 2 || variant == 3; }

__device__ __forceinline__ Node load_node(const Node *p) {
    // L2-coherent loads: parents are rewritten by this CTA while other threads keep walking
    const double2 a = __ldcg(reinterpret_cast<const double2 *>(p));
    const double2 b = __ldcg(reinterpret_cast<const double2 *>(p) + 1);
    Node n;
    n.x = a.x; n.y = a.y; n.z = b.x; n.parent = __double_as_longlong(b.y);
    return n;
}
__device__ __forceinline__ void store_parent(Node *p, long long parent) {
    __stcg(reinterpret_cast<long long *>(p) + 3, parent);
}

// ---- programmatic dependent launch (PDL): the kernels of an iteration are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so kernel N+1's CTAs may become resident while
// kernel N drains; every kernel therefore starts with pdl_wait() (returns once ALL memory operations
// of the preceding kernel in the stream are complete and visible).  The scans let their (small)
// successor launch right away; the small kernels never trigger early, so a scan grid is never parked
// on the SMs for longer than its predecessor's tail.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- dimension traits: which obstacle table, and which of the reference's norms evaluates what
template <int D> struct GeomOf;
template <> struct GeomOf<3> {
    typedef Geom3 type;
    static __host__ __device__ __forceinline__ Geom3 *ptr(const View &v, int e) { return v.geom + e; }
};
template <> struct GeomOf<2> {
    typedef Geom2 type;
    static __host__ __device__ __forceinline__ Geom2 *ptr(const View &v, int e) { return v.geom2 + e; }
};
template <int D> __device__ __forceinline__ void stage_geom(typename GeomOf<D>::type *dst, const View &v, int e) {
    typedef typename GeomOf<D>::type G;
    for (int i = threadIdx.x; i < (int)(sizeof(G) / sizeof(double)); i += blockDim.x)
        reinterpret_cast<double *>(dst)[i] = reinterpret_cast<const double *>(GeomOf<D>::ptr(v, e))[i];
}
// math.hypot(dx, dy[, dz]): cost() edges, Line(), steer distance (rrt_base_{2,3}d.py)
template <int D> __device__ __forceinline__ double edge_len(double dx, double dy, double dz) {
    return D == 3 ? hypot3(dx, dy, dz) : hypot2(dx, dy);
}
// squared distance used by the scans: exact decision value in 3D, pre-filter in 2D
template <int D> __device__ __forceinline__ double scan_sq(double dx, double dy, double dz) {
    return D == 3 ? sq3_rows(dx, dy, dz) : XADD(XMUL(dx, dx), XMUL(dy, dy));
}
// the vectorised distance of Near / ChooseParent / Rewire / goal scans:
// np.linalg.norm(axis=-1) in 3D (rrt_star_3d.py:82,94,103,136), np.hypot in 2D (rrt_star_2d.py:82,94,103,135)
template <int D> __device__ __forceinline__ double vec_dist(double dx, double dy, double dz) {
    return D == 3 ? rownorm3(dx, dy, dz) : np_hypot(dx, dy);
}
// np.linalg.norm(path[1:] - path[:-1], axis=1) rows (rrt_base_{2,3}d.py get_path_len)
template <int D> __device__ __forceinline__ double row_norm(double dx, double dy, double dz) {
    return D == 3 ? rownorm3(dx, dy, dz) : rownorm2(dx, dy);
}
// 2D scans decide with np.hypot (not correctly rounded, < 1 ulp): a squared-distance pre-filter
// with a 2^-49 relative safety band selects the few candidates that need the exact evaluation.
__device__ __forceinline__ double hypot_band_sq(double h) {
    return __dmul_ru(__dmul_ru(h, h), 1.0000000000000018);
}

// ---- root walks: cached edge lengths + ancestor hints
// RRTBase{2,3}D.cost (rrt_base_3d.py:60-67) walks parent pointers leaf -> root and sums
// math.hypot(v - parent(v)) in that order; the summation order rules out any re-association, and a
// plain walk is a chain of dependent DRAM round trips with a long math.hypot (CPython vector_norm:
// ~100 dependent f64 operations incl. a square root and a division) on every hop.  Both are removed
// without touching the arithmetic:
//   * links[v] = {math.hypot(v - parent(v)), parent(v)}: an edge length depends only on the two end
//     points, which never move, so it is evaluated once when the edge is created (insert,
//     ChooseParent, Rewire -- the same value cost() would recompute) and the walk only adds.
//   * hints[v] = {a1 .. a8} names the vertices that were v's ancestors 1..8 hops up when the
//     record was last written.  A walk loads the links of all eight and hints[a8] at once and verifies
//     the chain against the authoritative parent fields as the data arrives: eight hops per round
//     trip.  A stale hint (an ancestor was re-parented since) costs one ordinary hop and is repaired
//     on the spot.  Hints never decide anything.
struct __align__(16) Link {
    double elen;   // math.hypot(v - parent(v)); 0 for the root
    int parent;
    int pad;       // Near stamp: (iteration tag << kPosBits) | position in this iteration's Near list (k_expand), else stale/0
};
#ifdef NIRRT_PHASE_TIMING
__device__ unsigned long long g_walk_stats[4];
// k_expand phase clocks summed over CTAs (development build only, profiles/tools/phase_timing.py):
// [0] entry->steer done, [1] sort, [2] filter, [3] walks + ChooseParent, [4] Rewire, [5] goal bookkeeping, [6] records,
// [7] top (next sample), [8] CTA count
__device__ unsigned long long g_phase[16];
#endif
constexpr int kHintHops = 8;
struct __align__(32) Hint {     // one 32-byte sector
    int a[kHintHops];           // the vertices that were v's ancestors 1..8 hops up when the record was last written
};
struct TreeRef {
    Node *nodes;
    Link *links;
    Hint *hints;
    int cap;
};
__device__ __forceinline__ TreeRef tree_of(const View &v, int e) {
    TreeRef t;
    t.nodes = v.nodes + (size_t)e * v.stride; t.links = v.links + (size_t)e * v.stride;
    t.hints = v.hints + (size_t)e * v.stride; t.cap = v.cap;
    return t;
}
__device__ __forceinline__ Hint load_hint(const Hint *p) {
    const int4 lo = __ldcg(reinterpret_cast<const int4 *>(p)), hi = __ldcg(reinterpret_cast<const int4 *>(p) + 1);
    Hint h;
    h.a[0] = lo.x; h.a[1] = lo.y; h.a[2] = lo.z; h.a[3] = lo.w; h.a[4] = hi.x; h.a[5] = hi.y; h.a[6] = hi.z; h.a[7] = hi.w;
    return h;
}
__device__ __forceinline__ void store_hint_raw(Hint *p, const Hint &h) {
    __stcg(reinterpret_cast<int4 *>(p), make_int4(h.a[0], h.a[1], h.a[2], h.a[3]));
    __stcg(reinterpret_cast<int4 *>(p) + 1, make_int4(h.a[4], h.a[5], h.a[6], h.a[7]));
}
// hints of a vertex whose parent is a1 and whose parent's hints are `up`
__device__ __forceinline__ void store_hint(Hint *p, int a1, const Hint &up) {
    Hint h;
    h.a[0] = a1;
#pragma unroll
    for (int q = 1; q < kHintHops; q++) h.a[q] = up.a[q - 1];
    store_hint_raw(p, h);
}
__device__ __forceinline__ Hint zero_hint() {
    Hint h;
#pragma unroll
    for (int q = 0; q < kHintHops; q++) h.a[q] = 0;
    return h;
}
__device__ __forceinline__ Link load_link(const Link *p) {
    const int4 r = __ldcg(reinterpret_cast<const int4 *>(p));
    Link l;
    l.elen = __hiloint2double(r.y, r.x); l.parent = r.z; l.pad = r.w;
    return l;
}
// parent(v) = par with edge length elen: the node record (what read_trees returns) and the walk record
__device__ __forceinline__ void set_parent(const TreeRef &t, int v, int par, double elen) {
    store_parent(t.nodes + v, par);
    __stcg(reinterpret_cast<int4 *>(t.links + v), make_int4(__double2loint(elen), __double2hiint(elen), par, 0));
}

// hop(par, edge_length, pad field of par's link) is called once per edge, leaf -> root
template <typename F>
__device__ __forceinline__ void walk_to_root(const TreeRef &t, int idx, F &&hop) {
    if (idx == 0) return;
    Link cur = load_link(t.links + idx);
    Hint h = load_hint(t.hints + idx);
    int start = idx;
    for (;;) {
        const unsigned cap = (unsigned)t.cap;
        int a[kHintHops];
        Link l[kHintHops];
#pragma unroll
        for (int q = 0; q < kHintHops; q++) {
            a[q] = (unsigned)h.a[q] < cap ? h.a[q] : 0;
            l[q] = load_link(t.links + a[q]);
        }
        const Hint hn = load_hint(t.hints + a[kHintHops - 1]);
        int j = 0;                       // verified hops of this group
        int p = cur.parent;
        Link cj = cur;                   // link of the last verified vertex
        bool ok = true;
#pragma unroll
        for (int q = 0; q < kHintHops; q++) {
            if (ok) {
                if (p == a[q]) {
                    hop(p, cj.elen, l[q].pad);
                    if (p == 0) return;
                    cj = l[q]; p = cj.parent; j = q + 1;
                } else ok = false;
            }
        }
        if (ok) {
            cur = cj; start = a[kHintHops - 1]; h = hn;
#ifdef NIRRT_PHASE_TIMING
            atomicAdd(&g_walk_stats[0], 1ull);
#endif
            continue;
        }
#ifdef NIRRT_PHASE_TIMING
        atomicAdd(&g_walk_stats[1], 1ull);
#endif
        // stale from hop j+1 on (an ancestor was re-parented since hints[start] was written): one ordinary
        // hop from the last verified vertex, and hints[start] is rewritten with the verified prefix followed
        // by the true parent and that parent's own hints
        const Link lp = load_link(t.links + p);
        const Hint hp = load_hint(t.hints + p);
        Hint r;
#pragma unroll
        for (int q = 0; q < kHintHops; q++) {
            int val = q < j ? a[q] : p;
#pragma unroll
            for (int k = 0; k < kHintHops - 1; k++) if (q - j - 1 == k) val = hp.a[k];
            r.a[q] = val;
        }
        store_hint_raw(t.hints + start, r);
        hop(p, cj.elen, lp.pad);
        if (p == 0) return;
        cur = lp; start = p; h = hp;
    }
}

// RRTBase{2,3}D.cost (rrt_base_3d.py:60-67, rrt_base_2d.py:54-61): leaf -> root, math.hypot per
// edge, summed in that order
template <int D>
__device__ double cost_walk(const TreeRef &t, int idx) {
    double c = 0.0;
    walk_to_root(t, idx, [&](int, double e, int) { c = XADD(c, e); });
    return c;
}

// One walk, two sums in the reference's leaf -> root order:
//   c   = cost(idx)                                   (0 + e1) + e2 + ...
//   via = cost(x_new) if idx were x_new's parent      (first + e1) + e2 + ...   with first = hypot(x_new - v[idx])
// so neither ChooseParent's winner nor the steer parent needs a second walk for node_new_cost
// (rrt_star_3d.py:96).
template <int D>
__device__ __forceinline__ void cost_walk2(const TreeRef &t, int idx, double first, double &c_out, double &via_out) {
    double c = 0.0, a = first;
    walk_to_root(t, idx, [&](int, double e, int) { c = XADD(c, e); a = XADD(a, e); });
    c_out = c; via_out = a;
}

// lexicographic (value, index) min == np.argmin's first-minimum rule
__device__ __forceinline__ void lexmin(double &s, int &i, double os, int oi) {
    if (os < s || (os == s && oi < i)) { s = os; i = oi; }
}
__device__ __forceinline__ void warp_lexmin(double &s, int &i) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double os = __shfl_down_sync(0xffffffffu, s, off);
        const int oi = __shfl_down_sync(0xffffffffu, i, off);
        lexmin(s, i, os, oi);
    }
}
// block-wide (<= 1024 threads); result valid in thread 0 and broadcast through smem to all
__device__ void block_lexmin(double &s, int &i, double *sm_s, int *sm_i) {
    warp_lexmin(s, i);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (l == 0) { sm_s[w] = s; sm_i[w] = i; }
    __syncthreads();
    if (w == 0) {
        s = l < nw ? sm_s[l] : XINF;
        i = l < nw ? sm_i[l] : INT_MAX;
        warp_lexmin(s, i);
        if (l == 0) { sm_s[0] = s; sm_i[0] = i; }
    }
    __syncthreads();
    s = sm_s[0]; i = sm_i[0];
    __syncthreads();
}
__device__ int block_min_int(int v, int *sm_i) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = min(v, __shfl_down_sync(0xffffffffu, v, off));
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (l == 0) sm_i[w] = v;
    __syncthreads();
    if (w == 0) {
        v = l < nw ? sm_i[l] : INT_MAX;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v = min(v, __shfl_down_sync(0xffffffffu, v, off));
        if (l == 0) sm_i[0] = v;
    }
    __syncthreads();
    v = sm_i[0];
    __syncthreads();
    return v;
}

__device__ __forceinline__ void push_record(const View &v, EnvCtl *c, int e, double val) {
    if (c->n_rec < v.rec_cap) v.records[(size_t)e * v.rec_cap + c->n_rec] = val;
    else atomicOr(&c->err, ERR_RECORD_OVERFLOW);
    c->n_rec++;
}

// ---- mirror-scan helpers (see "Mirror scans" below)
constexpr double kMarginU16 = 2.0;
// u8 mirror: 255 cells over the longest world edge, vertex and query both rounded to a cell and all scan
// arithmetic in exact integers: |sqrt(d8^2) - d * qscale| <= sqrt(3)/2 + sqrt(3)/2 < kMarginU8 cells
constexpr double kMarginU8 = 1.75;
constexpr double kSpecSlack = 1e-6;   // world units: |x_new - x_rand| allowed when the speculative Near list is used (k_steer checks 1e-9 per axis)

__device__ __forceinline__ unsigned short quantize_u16(double x, double lo, double scale) {
    int q = __double2int_rn((x - lo) * scale);
    return (unsigned short)min(65535, max(0, q));
}
__device__ __forceinline__ unsigned quantize_u8(double x, double lo, double scale) {
    int q = __double2int_rn((x - lo) * scale);
    return (unsigned)min(255, max(0, q));
}
__device__ __forceinline__ void mirror_store(const View &v, const EnvCtl *c, size_t o, double x, double y, double z) {
    if (v.m8) {
        unsigned w = quantize_u8(x, c->qlo[0], c->qscale) | (quantize_u8(y, c->qlo[1], c->qscale) << 8);
        if (v.dim == 3) w |= quantize_u8(z, c->qlo[2], c->qscale) << 16;
        v.m8[o] = w;
    }
    if (v.ux) {
        v.ux[o] = quantize_u16(x, c->qlo[0], c->qscale);
        v.uy[o] = quantize_u16(y, c->qlo[1], c->qscale);
        if (v.uz) v.uz[o] = quantize_u16(z, c->qlo[2], c->qscale);
    }
    if (v.fx) {
        v.fx[o] = (float)x; v.fy[o] = (float)y;
        if (v.fz) v.fz[o] = (float)z;
    }
}
// the query in the mirror's coordinates (u16: 2^23 + nearest cell, see above)
template <bool kU16>
__device__ __forceinline__ float mirror_query(const EnvCtl *c, const double *q, int d) {
    if (kU16) return 8388608.0f + rintf((float)((q[d] - c->qlo[d]) * c->qscale));
    return (float)q[d];
}

template <int D>
__device__ __forceinline__ double exact_scan_value(const View &v, int e, int i, double qx, double qy, double qz) {
    const size_t o = (size_t)e * v.stride + i;
    const double dx = XSUB(qx, v.vx[o]), dy = XSUB(qy, v.vy[o]), dz = D == 3 ? XSUB(qz, v.vz[o]) : 0.0;
    return D == 3 ? XSQRT(sq3_rows(dx, dy, dz)) : np_hypot(dx, dy);
}

// One 32-byte record per scan holds everything a scan CTA needs to start streaming: a single round
// trip instead of a chain of dependent loads (the CTA's whole share is only ~10 us of traffic).
__device__ __forceinline__ ScanHdr load_hdr(const ScanHdr *h) {
    const int4 a = __ldcg(reinterpret_cast<const int4 *>(h));
    const int4 b = __ldcg(reinterpret_cast<const int4 *>(h) + 1);
    ScanHdr r;
    r.go = a.x; r.n = a.y; r.qx = __int_as_float(a.z); r.qy = __int_as_float(a.w);
    r.qz = __int_as_float(b.x); r.thr = __int_as_float(b.y); r.band = __int_as_float(b.z); r.base = b.w;
    return r;
}
// u8 mirror: the scan compares the exact integer a' = |m|^2 - 2 m.q8 (two DP4A per vertex) with thresholds that
// have |q8|^2 folded in; qx = packed query cells, qy = |q8|^2, thr = integer threshold - |q8|^2 (as bit patterns).
// A query outside the world range (never produced by the planners) selects everything: the exact fallbacks decide.
__device__ __forceinline__ void store_hdr_u8(ScanHdr *h, const EnvCtl *c, int dim, int go, int n, const double *q, float thr, int base) {
    ScanHdr r;
    r.go = go; r.n = n;
    unsigned packed = 0; int qq = 0; bool inside = true;
    for (int d = 0; d < dim; d++) {
        int qi = __double2int_rn((q[d] - c->qlo[d]) * c->qscale);
        if (qi < 0 || qi > 255) { inside = false; qi = min(255, max(0, qi)); }
        packed |= (unsigned)qi << (8 * d);
        qq += qi * qi;
    }
    r.qx = __uint_as_float(packed); r.qy = __int_as_float(qq); r.qz = 0.f;
    r.thr = __int_as_float(inside ? (int)fminf(thr, 1.0e9f) + 1 - qq : 0x3fffffff);
    r.band = inside ? __double2float_ru(2.0 * c->margin) : 3.0e4f;
    r.base = base;
    *h = r;
}
// SAD scan (k_nearest_s): qx = packed query cells, thr = L1 threshold in cells of the speculative Near ball.
// With vertex and query rounded to cells, every axis of the cell difference is within 1 of the scaled true difference,
// so a vertex at scaled L2 distance d has L1 cell distance <= sqrt(dim) * d + dim, and d <= L1 + dim.
__device__ __forceinline__ int sad_bound(const View &v, float d_cells) {
    return (int)(d_cells * (v.dim == 3 ? 1.7320509f : 1.4142136f) * 1.000001f) + v.dim + 1;
}
__device__ __forceinline__ void store_hdr_sad(const View &v, ScanHdr *h, const EnvCtl *c, int go, int n, const double *q, float thr_sq, int base) {
    ScanHdr r;
    r.go = go; r.n = n;
    unsigned packed = 0; bool inside = true;
    for (int d = 0; d < v.dim; d++) {
        int qi = __double2int_rn((q[d] - c->qlo[d]) * c->qscale);
        if (qi < 0 || qi > 255) { inside = false; qi = min(255, max(0, qi)); }
        packed |= (unsigned)qi << (8 * d);
    }
    r.qx = __uint_as_float(packed); r.qy = 0.f; r.qz = 0.f;
    // thr_sq = (scaled radius + margin)^2, rounded up: everything inside that ball has L1 <= sad_bound(radius)
    r.thr = __int_as_float(inside ? sad_bound(v, __fsqrt_ru(fminf(thr_sq, 1.0e9f))) : 0x3fffffff);
    r.band = 0.f; r.base = base;
    *h = r;
}
template <int kMirror>   // 2: u16 mirror, 1: f32 mirror, 0: none (f64 scans: only go / n matter)
__device__ __forceinline__ void store_hdr(ScanHdr *h, const EnvCtl *c, int go, int n, const double *q, float thr, int base) {
    ScanHdr r;
    r.go = go; r.n = n;
    r.qx = r.qy = r.qz = 0.f;
    if (kMirror) { r.qx = mirror_query<kMirror == 2>(c, q, 0); r.qy = mirror_query<kMirror == 2>(c, q, 1); r.qz = mirror_query<kMirror == 2>(c, q, 2); }
    r.thr = thr; r.band = __double2float_ru(2.0 * c->margin); r.base = base;
    *h = r;
}
__device__ __forceinline__ void write_hdr(const View &v, EnvCtl *c, int which, int go, const double *q, float thr, int base = 0) {
    ScanHdr *h = which == 0 ? &c->hdr0 : &IT.hdr1;
    if (v.m8 && v.sad && which == 0) { store_hdr_sad(v, h, c, go, c->n, q, thr, base); c->scan_min = 0x3fffffff; }
    else if (v.m8) store_hdr_u8(h, c, v.dim, go, c->n, q, thr, base);
    else if (v.ux) store_hdr<2>(h, c, go, c->n, q, thr, base);
    else if (v.fx) store_hdr<1>(h, c, go, c->n, q, thr, base);
    else store_hdr<0>(h, c, go, c->n, q, thr, base);
}

// the go flag of the NEXT iteration lives in the Nearest header only; steer_body copies it into its iteration copy
__device__ __forceinline__ void set_idle(EnvCtl *c) { c->hdr0.go = 0; }

// ------------------------------------------------------------------------------------------------
// k_top: driver phase machine + c_best refresh + sampling
//   RRTBase3D.SampleFree                 rrt_base_3d.py:49-58
//   IRRTStar3D.find_best_path_solution   irrt_star_3d.py:80-93
//   IRRTStar3D.SampleInformedSubset      irrt_star_3d.py:117-157
//   NIRRTStarPNG3D.generate_random_node  nirrt_star_png_3d.py:99-130
//   planning_random record/phase rules   irrt_star_3d.py:245-331
__device__ void sample_free(const Geom3 &g, MtStream &rng, MtStream &, double *out) {
    const double x0 = XADD(g.range[0], g.clearance), x1 = XSUB(g.range[1], g.clearance);
    const double y0 = XADD(g.range[2], g.clearance), y1 = XSUB(g.range[3], g.clearance);
    const double z0 = XADD(g.range[4], g.clearance), z1 = XSUB(g.range[5], g.clearance);
    do {
        out[0] = rng.uniform(x0, x1);
        out[1] = rng.uniform(y0, y1);
        out[2] = rng.uniform(z0, z1);
    } while (point_inside_obs(g, out));
}
// RRTBase2D.SampleFree (rrt_base_2d.py:46-52)
__device__ void sample_free(const Geom2 &g, MtStream &rng, MtStream &, double *out) {
    const double x0 = XADD(g.range[0], g.clearance), x1 = XSUB(g.range[1], g.clearance);
    const double y0 = XADD(g.range[2], g.clearance), y1 = XSUB(g.range[3], g.clearance);
    out[2] = 0.0;
    do {
        out[0] = rng.uniform(x0, x1);
        out[1] = rng.uniform(y0, y1);
    } while (point_inside_obs(g, out));
}

__device__ void sample_informed(const Geom3 &g, const EnvCtl *c, MtStream &rng, MtStream &, double c_max, double *out) {
    const double c2 = XSUB(XMUL(c_max, c_max), XMUL(c->c_min, c->c_min));
    const double eps = (c2 < 0.0) ? 1e-6 : 0.0;
    double r[3], M[9];
    r[0] = XDIV(c_max, 2.0);
    r[1] = r[2] = XDIV(XSQRT(XADD(c2, eps)), 2.0);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) M[3 * i + j] = XMUL(c->C[3 * i + j], r[j]);
    const double PI = 3.141592653589793, TWO_PI = 6.283185307179586;
    for (;;) {
        const double rr = rng.uniform(0.0, 1.0);
        const double th = rng.uniform(0.0, PI);
        const double ph = rng.uniform(0.0, TWO_PI);
        // np.sin / np.cos == glibc's sin / cos, restated operation by operation (glibc_trig.cuh)
        const double st = glibc_sin(th), ct = glibc_cos(th), sp = glibc_sin(ph), cp = glibc_cos(ph);
        const double rs = XMUL(rr, st);
        const double xb0 = XMUL(rs, cp), xb1 = XMUL(rs, sp), xb2 = XMUL(rr, ct);
        for (int i = 0; i < 3; i++)
            out[i] = XADD(XFMA(M[3 * i + 2], xb2, XFMA(M[3 * i], xb0, XMUL(M[3 * i + 1], xb1))), c->center[i]);
        if (point_valid(g, out)) break;
    }
}
// IRRTStar2D.SampleInformedSubset / SampleUnitBall (irrt_star_2d.py:121-151): the unit-disc draw
// consumes the CPython `random` stream (py), node = (C L) x_ball + x_center evaluated as numpy's
// dgemv does it (fma(M0, x, M1 * y) + centre, probed).
__device__ void sample_informed(const Geom2 &g, const EnvCtl *c, MtStream &, MtStream &py, double c_max, double *out) {
    const double c2 = XSUB(XMUL(c_max, c_max), XMUL(c->c_min, c->c_min));
    const double eps = (c2 < 0.0) ? 1e-6 : 0.0;
    const double r0 = XDIV(c_max, 2.0), r1 = XDIV(XSQRT(XADD(c2, eps)), 2.0);
    const double M00 = XMUL(c->C[0], r0), M01 = XMUL(c->C[1], r1), M10 = XMUL(c->C[3], r0), M11 = XMUL(c->C[4], r1);
    out[2] = 0.0;
    for (;;) {
        double x, y;
        do {
            x = py.uniform(-1.0, 1.0);
            y = py.uniform(-1.0, 1.0);
        } while (!(XADD(XMUL(x, x), XMUL(y, y)) < 1.0));
        out[0] = XADD(XFMA(M00, x, XMUL(M01, y)), c->center[0]);
        out[1] = XADD(XFMA(M10, x, XMUL(M11, y)), c->center[1]);
        if (point_valid(g, out)) break;
    }
}

// SampleFree (rrt_base_{2,3}d.py:46-58) evaluated speculatively by the whole CTA: attempt a consumes the words
// [2*D*a, 2*D*(a+1)) of the staged stream, one warp per attempt, the lanes of a warp split the OR over
// the obstacles; the first free attempt wins -- exactly the point and the stream position the sequential
// rejection loop arrives at.  words_out = 0: undecided inside the staged words (the caller's serial loop
// takes over).  Nothing is committed here; the sampling thread skips the words if the driver samples.
// w0: words the driver consumes before the sampler (2 for the NRRT* / NIRRT* cloud-or-not draw)
template <int D, typename G>
__device__ __forceinline__ void spec_sample_free(const G &g, const MtState *st, const uint32_t *cache_all, int w0, double *s_out, int *words_out) {
    __shared__ int s_win;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    constexpr int W = 2 * D;
    const uint32_t *cache = cache_all + w0;
    const int avail = min(kMtCache, 624 - st->pos) - w0;
    const int max_att = avail > 0 ? avail / W : 0;
    if (tid == 0) { s_win = INT_MAX; *words_out = 0; }
    __syncthreads();
    double lo[3] = {0.0, 0.0, 0.0}, hi[3] = {0.0, 0.0, 0.0};
    for (int d = 0; d < D; d++) { lo[d] = XADD(g.range[2 * d], g.clearance); hi[d] = XSUB(g.range[2 * d + 1], g.clearance); }
    const int m = n_obstacles(g);
    for (int a0 = 0; a0 < max_att; a0 += nw) {
        const int a = a0 + warp;
        double p[3] = {0.0, 0.0, 0.0};
        if (a < max_att) {
            for (int d = 0; d < D; d++)
                p[d] = XADD(lo[d], XMUL(XSUB(hi[d], lo[d]), mt_double(cache[a * W + 2 * d], cache[a * W + 2 * d + 1])));
            bool hit = false;
            for (int k = lane; k < m; k += 32) hit = hit || point_in_obstacle(g, k, p);
            if (!__any_sync(0xffffffffu, hit) && lane == 0) atomicMin(&s_win, a);
        }
        __syncthreads();
        const int win = s_win;
        if (win != INT_MAX) {
            if (a == win && lane == 0) { s_out[0] = p[0]; s_out[1] = p[1]; s_out[2] = p[2]; *words_out = (win + 1) * W; }
            break;
        }
    }
    __syncthreads();
}

// IRRTStar3D.SampleInformedSubset (irrt_star_3d.py:117-157) evaluated speculatively: thread a carries out attempt a of the
// rejection loop on the words [w0 + 6a, w0 + 6a + 6) of the staged stream -- the same operations as sample_informed, four glibc
// sin / cos evaluations and the obstacle tests included -- and the first valid attempt wins: the point and the stream position of
// the sequential loop, at the latency of ONE attempt instead of the 2-3 a cluttered world needs (measured 30 us per sample before).
__device__ __forceinline__ void spec_sample_informed(const Geom3 &g, const EnvCtl *c, const MtState *st, const uint32_t *cache_all, int w0,
                                                     double c_max, double *s_out, int *words_out) {
    __shared__ int s_win_i;
    const int tid = threadIdx.x;
    const uint32_t *cache = cache_all + w0;
    const int avail = min(kMtCache, 624 - st->pos) - w0;
    const int max_att = avail > 0 ? avail / 6 : 0;
    if (tid == 0) { s_win_i = INT_MAX; *words_out = 0; }
    __syncthreads();
    double out[3] = {0.0, 0.0, 0.0};
    if (tid < max_att) {
        const double c2 = XSUB(XMUL(c_max, c_max), XMUL(c->c_min, c->c_min));
        const double eps = (c2 < 0.0) ? 1e-6 : 0.0;
        double r[3], M[9];
        r[0] = XDIV(c_max, 2.0);
        r[1] = r[2] = XDIV(XSQRT(XADD(c2, eps)), 2.0);
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) M[3 * i + j] = XMUL(c->C[3 * i + j], r[j]);
        const double PI = 3.141592653589793, TWO_PI = 6.283185307179586;
        const uint32_t *w = cache + 6 * tid;
        // np.random.uniform(lo, hi) == lo + (hi - lo) * next_double()
        const double rr = XADD(0.0, XMUL(XSUB(1.0, 0.0), mt_double(w[0], w[1])));
        const double th = XADD(0.0, XMUL(XSUB(PI, 0.0), mt_double(w[2], w[3])));
        const double ph = XADD(0.0, XMUL(XSUB(TWO_PI, 0.0), mt_double(w[4], w[5])));
        const double st_ = glibc_sin(th), ct = glibc_cos(th), sp = glibc_sin(ph), cp = glibc_cos(ph);
        const double rs = XMUL(rr, st_);
        const double xb0 = XMUL(rs, cp), xb1 = XMUL(rs, sp), xb2 = XMUL(rr, ct);
        for (int i = 0; i < 3; i++)
            out[i] = XADD(XFMA(M[3 * i + 2], xb2, XFMA(M[3 * i], xb0, XMUL(M[3 * i + 1], xb1))), c->center[i]);
        if (point_valid(g, out)) atomicMin(&s_win_i, tid);
    }
    __syncthreads();
    if (tid == s_win_i) { s_out[0] = out[0]; s_out[1] = out[1]; s_out[2] = out[2]; *words_out = (tid + 1) * 6; }
    __syncthreads();
}
__device__ __forceinline__ void spec_sample_informed(const Geom2 &, const EnvCtl *, const MtState *, const uint32_t *, int, double, double *, int *words_out) {
    if (threadIdx.x == 0) *words_out = 0;     // 2D draws the unit disc from the CPython stream: cheap, left to the sampling thread
    __syncthreads();
}

// Called by all 128 threads of the env's CTA: from k_top (first iteration of a run) and from the tail
// of k_expand (every following iteration -- one launch and one dependent round trip less per iteration).
template <int D>
__device__ __forceinline__ void top_body(const View &v, int e, typename GeomOf<D>::type &g, bool g_staged, double *sm_s, int *sm_i) {
    EnvCtl *c = v.ctl + e;
    const int state = c->state, budget = c->budget;
    if (state == ST_DONE || state == ST_WAIT_CLOUD || budget <= 0 || (c->cfg.n_limit > 0 && c->n >= c->cfg.n_limit)) {
        if (threadIdx.x == 0) set_idle(c);
        return;
    }
    if (!g_staged) stage_geom<D>(&g, v, e);
    __shared__ uint32_t s_mt[kMtCache];
    mt_prepare_next(v.mt + e, 160);
    mt_stage_words(v.mt + e, s_mt);
    if (D == 2 && fam_informed(v.variant)) mt_prepare_next(v.mt_py + e, 160);
    __syncthreads();

    // find_best_path_solution (irrt_star_3d.py:80-93): kept current incrementally by goal_track (the reference re-walks
    // every stored solution at the top of every iteration -- thousands of walks in a dense informed tree)
    double c_best = XINF;
    if (fam_informed(v.variant) && c->n_sol > 0) c_best = c->best_val;
    // The sampler the driver will call if it samples (SampleFree, or the informed sampler once a solution exists) is evaluated
    // by all threads on the staged words; NRRT* / NIRRT* first draw whether the sample comes from the guidance cloud (2 words)
    __shared__ double s_spec[3];
    __shared__ int s_spec_words;
    const int spec_w0 = fam_cloud(v.variant) ? 2 : 0;
    const bool informed_now = fam_informed(v.variant) && c_best < XINF;
    if (informed_now) spec_sample_informed(g, c, v.mt + e, s_mt, spec_w0, c_best, s_spec, &s_spec_words);
    else spec_sample_free<D>(g, v.mt + e, s_mt, spec_w0, s_spec, &s_spec_words);
    if (threadIdx.x != 0) return;

    const bool fresh = !(v.variant == 2 && c->resumed);
    if (fam_informed(v.variant) && fresh) {
        c->c_best = c_best;
        if (v.mode == NIRRT_MODE_PLANNING_RANDOM) {
            if (c->state == ST_PHASE1) {
                if (c_best < c->cfg.stop_below) { c->state = ST_PHASE2; c->left = c->cfg.iter_after; }
                else if (c->p1_done >= c->cfg.iter_max) { push_record(v, c, e, c_best); c->state = ST_DONE; set_idle(c); return; }
            }
            if (c->state == ST_PHASE2 && c->left <= 0) { push_record(v, c, e, c_best); c->state = ST_DONE; set_idle(c); return; }
            push_record(v, c, e, c_best);
        }
    }
    if (v.variant == 2 && fresh && c_best < XMUL(c->cfg.pc_ratio, c->c_update)) {
        // update_point_cloud (nirrt_star_png_3d.py:113-115): the host runs PointNet++ and resumes us
        c->c_update = c_best;
        c->saved_state = c->state;
        c->state = ST_WAIT_CLOUD;
        c->resumed = 1;
        set_idle(c);
        return;
    }
    c->resumed = 0;

    MtStream rng(v.mt + e, s_mt);
    MtStream py(D == 2 ? v.mt_py + e : v.mt + e);
    double out[3];
    bool done = false;
    if (fam_cloud(v.variant)) {
        if (rng.next_double() < c->cfg.pc_rate) {
            if (c->n_pc <= 0) { atomicOr(&c->err, ERR_EMPTY_CLOUD); c->state = ST_DONE; set_idle(c); rng.flush(); return; }
            const long long k = rng.randint(c->n_pc);
            const double *p = v.pc + ((size_t)e * v.pc_cap + k) * 3;
            out[0] = p[0]; out[1] = p[1]; out[2] = p[2];
            done = true;
        }
    }
    if (!done) {
        if (s_spec_words > 0) { out[0] = s_spec[0]; out[1] = s_spec[1]; out[2] = s_spec[2]; rng.skip(s_spec_words); }
        else if (informed_now) sample_informed(g, c, rng, py, c_best, out);
        else sample_free(g, rng, py, out);
    }
    rng.flush();
    if (D == 2) py.flush();
    c->x_rand[0] = out[0]; c->x_rand[1] = out[1]; c->x_rand[2] = out[2];
    c->cand_cnt = 0;     // the Nearest mirror scan appends its in-band candidates here
    // the iteration this sample belongs to works on copy par ^ pipe; its speculative list starts at the copy's
    // current (never reset) counter value -- nothing of that copy is written here, its previous user may still run
    const int spec_base = c->s[v.par ^ v.pipe].spec_cnt;
    // Speculative Near: whenever the tree already reaches within step_len of x_rand (always, once it is
    // dense) Steer returns x_new == x_rand up to rounding, so the ball around x_rand with the largest
    // radius the insertion can produce (the table is evaluated at n and n + 1) is a superset of Near(x_new):
    // the Nearest scan collects it in the same pass and the second scan of the iteration disappears.
    {
        const int n = c->n;
        double rs = XMUL(c->search_radius, fmax(v.near_table[n], v.near_table[n + 1]));
        if (c->step_len < rs) rs = c->step_len;
        const double rm = ((v.ux || v.m8) ? (rs + kSpecSlack) * c->qscale : rs + kSpecSlack) + c->margin;
        write_hdr(v, c, 0, 1, c->x_rand, __double2float_ru(rm * rm * 1.000001), spec_base);
    }
}

template <int D>
__global__ void __launch_bounds__(128) k_top(View v) {
    typedef typename GeomOf<D>::type G;
    pdl_wait();
    __shared__ G g;
    __shared__ double sm_s[4];
    __shared__ int sm_i[4];
    top_body<D>(v, v.env0 + blockIdx.x, g, false, sm_s, sm_i);
}

// ------------------------------------------------------------------------------------------------
// k_nearest: RRTBase3D.nearest_neighbor (rrt_base_3d.py:100-113)
//   argmin_i sqrt((dx*dx + dy*dy) + dz*dz), first index on ties.  The square root is only taken
//   for running-minimum candidates: s2 >= RU(best_s*best_s) implies sqrt_rn(s2) >= best_s.
template <int D, bool kForce>
__global__ void __launch_bounds__(256) k_nearest(View v) {
    pdl_wait();
    pdl_launch_dependents();
    const int e = v.env0 + blockIdx.y;
    const EnvCtl *c = v.ctl + e;
    if (!kForce && !c->hdr0.go) return;
    const int n = c->n;
    const int per = (((n + (int)gridDim.x - 1) / (int)gridDim.x) + 1) & ~1;
    const int beg = blockIdx.x * per;
    const int end = min(n, beg + per);
    const double *X = v.vx + (size_t)e * v.stride, *Y = v.vy + (size_t)e * v.stride;
    const double *Z = D == 3 ? v.vz + (size_t)e * v.stride : nullptr;
    const double qx = c->x_rand[0], qy = c->x_rand[1], qz = c->x_rand[2];
    double best_s = XINF, T = XINF;
    int best_i = INT_MAX;

    // 3D: argmin of sqrt((dx*dx + dy*dy) + dz*dz); the square root is only taken for running-minimum
    // candidates (s2 >= RU(best*best) implies sqrt_rn(s2) >= best).
    // 2D: argmin of np.hypot(dx, dy) (rrt_base_2d.py:105-106); evaluated only inside the band.
#define NEAREST_ONE(xx, yy, zz, ii)                                                   \
    {                                                                                 \
        const double dx = XSUB(qx, xx), dy = XSUB(qy, yy), dz = D == 3 ? XSUB(qz, zz) : 0.0; \
        const double s2 = scan_sq<D>(dx, dy, dz);                                     \
        if (D == 3) {                                                                 \
            if (s2 < T) {                                                             \
                const double s = XSQRT(s2);                                           \
                if (s < best_s) { best_s = s; best_i = (ii); T = __dmul_ru(s, s); }   \
            }                                                                         \
        } else if (s2 <= T) {                                                         \
            const double s = np_hypot(dx, dy);                                        \
            if (s < best_s) { best_s = s; best_i = (ii); T = hypot_band_sq(s); }      \
        }                                                                             \
    }
    const int step = 2 * blockDim.x;
    int i = beg + 2 * threadIdx.x;
    for (; i + step < end; i += 2 * step) {   // two independent 16-byte loads per array in flight
        const double2 xa = __ldg(reinterpret_cast<const double2 *>(X + i));
        const double2 ya = __ldg(reinterpret_cast<const double2 *>(Y + i));
        const double2 xb = __ldg(reinterpret_cast<const double2 *>(X + i + step));
        const double2 yb = __ldg(reinterpret_cast<const double2 *>(Y + i + step));
        double2 za = make_double2(0.0, 0.0), zb = za;
        if (D == 3) {
            za = __ldg(reinterpret_cast<const double2 *>(Z + i));
            zb = __ldg(reinterpret_cast<const double2 *>(Z + i + step));
        }
        NEAREST_ONE(xa.x, ya.x, za.x, i)
        NEAREST_ONE(xa.y, ya.y, za.y, i + 1)          // i+1 < i+step < end
        NEAREST_ONE(xb.x, yb.x, zb.x, i + step)
        if (i + step + 1 < end) NEAREST_ONE(xb.y, yb.y, zb.y, i + step + 1)
    }
    for (; i < end; i += step) {
        const double2 xa = __ldg(reinterpret_cast<const double2 *>(X + i));
        const double2 ya = __ldg(reinterpret_cast<const double2 *>(Y + i));
        double2 za = make_double2(0.0, 0.0);
        if (D == 3) za = __ldg(reinterpret_cast<const double2 *>(Z + i));
        NEAREST_ONE(xa.x, ya.x, za.x, i)
        if (i + 1 < end) NEAREST_ONE(xa.y, ya.y, za.y, i + 1)
    }
#undef NEAREST_ONE
    __shared__ double sm_s[8];
    __shared__ int sm_i[8];
    warp_lexmin(best_s, best_i);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sm_s[w] = best_s; sm_i[w] = best_i; }
    __syncthreads();
    if (w == 0) {
        best_s = l < 8 ? sm_s[l] : XINF;
        best_i = l < 8 ? sm_i[l] : INT_MAX;
        warp_lexmin(best_s, best_i);
        if (l == 0) {
            v.part_s[(size_t)e * v.chunks + blockIdx.x] = best_s;
            v.part_i[(size_t)e * v.chunks + blockIdx.x] = best_i;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_steer: argmin finish + new_state (rrt_star_3d.py:67-78 / rrt_star_2d.py:67-78) + steer-edge
//          collision + duplicate guard / vertex insert (:40-51) + Near radius (:134 / :133)
// Called by ONE warp per env: the k_steer kernel, or warp 0 of k_expand when the two are fused.
template <int D>
__device__ __forceinline__ void steer_body(const View &v, int e, const typename GeomOf<D>::type *staged = nullptr) {
    typedef typename GeomOf<D>::type G;
    EnvCtl *c = v.ctl + e;
    const int lane = threadIdx.x & 31;
    const int go = __ldcg(&c->hdr0.go), spec_base = __ldcg(&c->hdr0.base), cnt = __ldcg(&c->cand_cnt);     // independent loads: one round trip (L2: cand_cnt is updated by atomics)
    if (lane == 0) { IT.go = go; IT.spec_base = spec_base; }     // what k_expand of this iteration reads
    if (!go) return;
    Node *nodes = v.nodes + (size_t)e * v.stride;
    Hint *hints = v.hints + (size_t)e * v.stride;
    double bs = XINF; int bi = INT_MAX;
    Node nn; Hint hn = zero_hint();              // record + hints of this lane's best candidate
    nn.x = nn.y = nn.z = 0.0; nn.parent = 0;
    bool have_rec = false;
    if (has_mirror(v)) {
        // candidates of the mirror scan: exact distance, lexicographic (value, index) minimum.  The exact
        // coordinates come from the 32-byte node record (the same doubles as vx/vy/vz, one sector instead
        // of three), so the winner's record is already here when Steer needs it.
        const double qx = c->x_rand[0], qy = c->x_rand[1], qz = c->x_rand[2];
        if (cnt <= v.near_cap) {
            const int *cand = v.cand + (size_t)e * v.near_cap;
            have_rec = true;
            for (int k = lane; k < cnt; k += 32) {
                const int i = __ldcg(cand + k);
                const Node nd = load_node(nodes + i);
                const Hint hh = load_hint(hints + i);
                const double dx = XSUB(qx, nd.x), dy = XSUB(qy, nd.y), dz = D == 3 ? XSUB(qz, nd.z) : 0.0;
                const double val = D == 3 ? XSQRT(sq3_rows(dx, dy, dz)) : np_hypot(dx, dy);
                if (val < bs || (val == bs && i < bi)) { bs = val; bi = i; nn = nd; hn = hh; }
            }
        } else {
            const int n = c->n;
            for (int i = lane; i < n; i += 32) lexmin(bs, bi, exact_scan_value<D>(v, e, i, qx, qy, qz), i);
            if (lane == 0) c->fallbacks++;
        }
    } else {
        for (int k = lane; k < v.chunks; k += 32) lexmin(bs, bi, v.part_s[(size_t)e * v.chunks + k], v.part_i[(size_t)e * v.chunks + k]);
    }
    const int my_bi = bi;
    warp_lexmin(bs, bi);
    const int nearest = __shfl_sync(0xffffffffu, bi, 0);
    const G &g = staged ? *staged : *GeomOf<D>::ptr(v, e);     // the persistent kernel keeps the obstacle table in shared memory
    if (have_rec) {      // broadcast the winning lane's record
        const int src = __ffs(__ballot_sync(0xffffffffu, my_bi == nearest)) - 1;
        nn.x = __shfl_sync(0xffffffffu, nn.x, src); nn.y = __shfl_sync(0xffffffffu, nn.y, src);
        nn.z = __shfl_sync(0xffffffffu, nn.z, src);
#pragma unroll
        for (int q = 0; q < kHintHops; q++) hn.a[q] = __shfl_sync(0xffffffffu, hn.a[q], src);
    } else {
        nn = load_node(nodes + nearest);
        hn = load_hint(hints + nearest);
    }
    const double xn[3] = {nn.x, nn.y, nn.z};
    // every lane computes the same x_new (cheap, avoids broadcasts)
    const double d0 = XSUB(c->x_rand[0], xn[0]), d1 = XSUB(c->x_rand[1], xn[1]), d2 = D == 3 ? XSUB(c->x_rand[2], xn[2]) : 0.0;
    double dist = edge_len<D>(d0, d1, d2);
    double xnew[3];
    if (D == 3) {
        double dir[3] = {0.0, 0.0, 0.0};
        if (dist != 0.0) { dir[0] = XDIV(d0, dist); dir[1] = XDIV(d1, dist); dir[2] = XDIV(d2, dist); }
        if (!(dist < c->step_len)) dist = c->step_len;   // min(step_len, dist)
        for (int i = 0; i < 3; i++) xnew[i] = XADD(xn[i], XMUL(dist, dir[i]));
    } else {
        // theta = math.atan2(dy, dx); node_new = start + dist * [cos(theta), sin(theta)]
        const double theta = glibc_atan2(d1, d0);      // math.atan2, glibc kernels restated (glibc_trig.cuh)
        if (!(dist < c->step_len)) dist = c->step_len;
        const double sn = glibc_sin(theta), cs = glibc_cos(theta);     // math.sin / math.cos (glibc_trig.cuh)
        xnew[0] = XADD(xn[0], XMUL(dist, cs));
        xnew[1] = XADD(xn[1], XMUL(dist, sn));
        xnew[2] = 0.0;
    }
    const int m = n_obstacles(g);
    bool hit = false;
    for (int k = lane; k < m; k += 32) hit = hit || seg_hits_obstacle(g, k, xn, xnew);
    hit = __any_sync(0xffffffffu, hit);
    if (lane != 0) return;
    IT.nearest = nearest;
    c->cand_cnt = 0;
    IT.near_cnt = 0;
    IT.inserted = 0;
    IT.hdr1.go = 0;
    if (hit) { IT.skip = 1; IT.new_idx = -1; return; }
    IT.skip = 0;
    int new_idx;
    const double dup = D == 3 ? vecnorm3(XSUB(xnew[0], xn[0]), XSUB(xnew[1], xn[1]), XSUB(xnew[2], xn[2]))
                              : vecnorm2(XSUB(xnew[0], xn[0]), XSUB(xnew[1], xn[1]));
    if (dup < 1e-8) {
        // "do not create a new node if it is actually the same point" (rrt_star_3d.py:41-45)
        xnew[0] = xn[0]; xnew[1] = xn[1]; xnew[2] = xn[2];
        new_idx = nearest;
        IT.cnew_default = -1.0;        // marker: x_new re-uses an existing vertex (k_expand walks it)
    } else {
        new_idx = c->n;
        if (new_idx >= v.cap) { atomicOr(&c->err, ERR_VERTEX_OVERFLOW); IT.skip = 1; IT.new_idx = -1; return; }
        const size_t o = (size_t)e * v.stride + new_idx;
        v.vx[o] = xnew[0]; v.vy[o] = xnew[1];
        if (D == 3) v.vz[o] = xnew[2];
        mirror_store(v, c, o, xnew[0], xnew[1], xnew[2]);
        Node nd; nd.x = xnew[0]; nd.y = xnew[1]; nd.z = xnew[2]; nd.parent = nearest;
        nodes[new_idx] = nd;
        store_hint(hints + new_idx, nearest, hn);
        c->n = new_idx + 1;
        IT.inserted = 1;
        c->tree_changed = 1;
        // Line(nearest, new) (rrt_base_3d.py:132-137); the root walk that turns it into
        // curr_node_new_cost and node_new_cost runs in k_expand together with the neighbours' walks
        IT.cnew_default = edge_len<D>(XSUB(xnew[0], xn[0]), XSUB(xnew[1], xn[1]), XSUB(xnew[2], xn[2]));
        set_parent(tree_of(v, e), new_idx, nearest, IT.cnew_default);
    }
    IT.new_idx = new_idx;
    IT.x_new[0] = xnew[0]; IT.x_new[1] = xnew[1]; IT.x_new[2] = xnew[2];
    double r = XMUL(c->search_radius, v.near_table[c->n]);
    if (c->step_len < r) r = c->step_len;            // min(gamma * f(n), step_len)
    IT.r = r;
    IT.T_near = D == 3 ? sqrt_le_threshold(r) : hypot_band_sq(r);
    {   // mirror pre-filter threshold: every vertex with f64 distance <= r has mirror squared distance <= near_thr
        const double rm = ((v.ux || v.m8) ? r * c->qscale : r) + c->margin;
        IT.near_thr = __double2float_ru(rm * rm * 1.000001);
        write_hdr(v, c, 1, 1, IT.x_new, IT.near_thr);
    }
    if (has_mirror(v)) {
        // x_new == x_rand up to rounding (the tree reaches within step_len of the sample): the ball the Nearest
        // scan collected around x_rand contains Near(x_new) -- no second scan.  Otherwise k_expand scans itself.
        bool same = fabs(XSUB(xnew[0], c->x_rand[0])) <= 1e-9 && fabs(XSUB(xnew[1], c->x_rand[1])) <= 1e-9;
        if (D == 3) same = same && fabs(XSUB(xnew[2], c->x_rand[2])) <= 1e-9;
        const int use_spec = same && __ldcg(&IT.spec_cnt) - spec_base <= v.near_cap;
        IT.use_spec = use_spec;
        IT.need_scan = !use_spec;
    } else { IT.use_spec = 0; IT.need_scan = 0; }
}

template <int D>
__global__ void __launch_bounds__(32) k_steer(View v) {
    pdl_wait();
    steer_body<D>(v, v.env0 + blockIdx.x);
}

// Pipelined RRT* driver (planning() loop body, rrt_star_3d.py:36-55): nothing the NEXT iteration's sample and
// Nearest scan need depends on this iteration's ChooseParent / Rewire -- only on the inserted vertex and the
// RNG stream.  k_front therefore does Steer of iteration i, the loop accounting and the sample of iteration
// i + 1, so that scan(i + 1) streams while k_expand(i) (launched on a second stream, View::pipe = 1) still walks.
// The two iterations in flight work on the two IterScratch copies (View::par).
template <int D>
__global__ void __launch_bounds__(128) k_front(View v) {
    typedef typename GeomOf<D>::type G;
    pdl_wait();
    const int e = v.env0 + blockIdx.x;
    EnvCtl *c = v.ctl + e;
    if (!c->hdr0.go) {
        if (threadIdx.x == 0) IT.go = 0;
        return;
    }
    __shared__ G g;
    __shared__ double sm_s[4];
    __shared__ int sm_i[4];
    if (threadIdx.x < 32) steer_body<D>(v, e);
    __syncthreads();
    if (threadIdx.x == 0) {      // the accounting k_expand does at its end in the unpipelined flow
        c->budget--;
        c->p1_done++;
        if (c->p1_done >= c->cfg.iter_max) c->state = ST_DONE;
    }
    __syncthreads();
    top_body<D>(v, e, g, false, sm_s, sm_i);
}

// ------------------------------------------------------------------------------------------------
// k_near: the distance part of find_near_neighbors (rrt_star_3d.py:134-137 / rrt_star_2d.py:133-136):
//   3D  np.where(np.linalg.norm(node_new - vertices, axis=-1) <= r)  <=>  sq <= T_near (exact, see
//       sqrt_le_threshold)
//   2D  np.where(np.hypot(dx, dy) <= r): squared pre-filter, exact np.hypot inside the band.
// Matches are sparse (tens out of 1e5) and are appended unordered; k_expand sorts them back into
// ascending index order.
template <int D, bool kForce>
__global__ void __launch_bounds__(256) k_near(View v) {
    pdl_wait();
    pdl_launch_dependents();
    const int e = v.env0 + blockIdx.y;
    EnvCtl *c = v.ctl + e;
    if (!kForce && (!IT.go || IT.skip)) return;
    const int n = c->n;
    const int per = (((n + (int)gridDim.x - 1) / (int)gridDim.x) + 1) & ~1;
    const int beg = blockIdx.x * per;
    const int end = min(n, beg + per);
    const double *X = v.vx + (size_t)e * v.stride, *Y = v.vy + (size_t)e * v.stride;
    const double *Z = D == 3 ? v.vz + (size_t)e * v.stride : nullptr;
    const double qx = IT.x_new[0], qy = IT.x_new[1], qz = IT.x_new[2];
    const double T = IT.T_near, r = IT.r;
    int *cand = v.cand + (size_t)e * v.near_cap;

#define NEAR_ONE(xx, yy, zz, ii)                                                      \
    {                                                                                 \
        const double dx = XSUB(qx, xx), dy = XSUB(qy, yy), dz = D == 3 ? XSUB(qz, zz) : 0.0; \
        if (scan_sq<D>(dx, dy, dz) <= T && (D == 3 || np_hypot(dx, dy) <= r)) {       \
            const int slot = atomicAdd(&c->cand_cnt, 1);                              \
            if (slot < v.near_cap) cand[slot] = (ii);                                 \
        }                                                                             \
    }
    const int step = 2 * blockDim.x;
    int i = beg + 2 * threadIdx.x;
    for (; i + step < end; i += 2 * step) {
        const double2 xa = __ldg(reinterpret_cast<const double2 *>(X + i));
        const double2 ya = __ldg(reinterpret_cast<const double2 *>(Y + i));
        const double2 xb = __ldg(reinterpret_cast<const double2 *>(X + i + step));
        const double2 yb = __ldg(reinterpret_cast<const double2 *>(Y + i + step));
        double2 za = make_double2(0.0, 0.0), zb = za;
        if (D == 3) {
            za = __ldg(reinterpret_cast<const double2 *>(Z + i));
            zb = __ldg(reinterpret_cast<const double2 *>(Z + i + step));
        }
        NEAR_ONE(xa.x, ya.x, za.x, i)
        NEAR_ONE(xa.y, ya.y, za.y, i + 1)
        NEAR_ONE(xb.x, yb.x, zb.x, i + step)
        if (i + step + 1 < end) NEAR_ONE(xb.y, yb.y, zb.y, i + step + 1)
    }
    for (; i < end; i += step) {
        const double2 xa = __ldg(reinterpret_cast<const double2 *>(X + i));
        const double2 ya = __ldg(reinterpret_cast<const double2 *>(Y + i));
        double2 za = make_double2(0.0, 0.0);
        if (D == 3) za = __ldg(reinterpret_cast<const double2 *>(Z + i));
        NEAR_ONE(xa.x, ya.x, za.x, i)
        if (i + 1 < end) NEAR_ONE(xa.y, ya.y, za.y, i + 1)
    }
#undef NEAR_ONE
}

// ------------------------------------------------------------------------------------------------
// Mirror scans.  The two HBM-bound passes read a compact copy of the coordinates and use it only as
// a conservative filter; the f64 SoA stays the source of truth and every decision is re-made with the
// reference's exact arithmetic on the few vertices the filter lets through.
//   u16 mirror (default, 2 B / coordinate): cell = rint((x - qlo) * qscale), 65535 cells over the
//       longest world edge.  The scan works in cell units with the query rounded to a cell as well,
//       so the subtraction is exact: q and the vertex are both integers carried in floats
//       (as_float(0x4B000000 | cell) == 2^23 + cell, one PRMT per coordinate, no conversion).
//       |sqrt(a) - d * qscale| <= sqrt(3)/2 (vertex rounding) + sqrt(3)/2 (query rounding) + 0.02
//       (three float roundings of a <= 1.3e10) < kMarginU16 = 2 cells for vertices inside the range.
//   f32 mirror (NIRRT_SCAN=f32, 4 B / coordinate): |sqrt(a) - d| < margin = 2^-19 * largest |range bound|.
// Nearest: the f64 argmin of a chunk lies within 2 * margin of the chunk's mirror minimum.  Every
//   thread tracks its best (value, index) and its second-best value; after ONE block-wide minimum
//   (redux + shared atomicMin + one barrier) the threads whose best is inside the band append it to
//   the env's candidate list, a thread whose second-best is inside the band too re-walks its own
//   vertices and appends all of them (rare).  k_steer evaluates the exact formula on the candidates
//   and takes the lexicographic (value, index) minimum == np.argmin.  If the list overflows k_steer
//   re-scans the env in f64 (counter `fallbacks`).
// Near: mirror squared distance <= near_thr is a superset of the exact set; k_expand applies the exact
//   test to every candidate before anything else looks at it.
// Results are bit-identical to the f64 scans (same parity tests).
// Calls visit(a[kVec], base) with the mirror squared distances of vertices base .. base+kVec-1 for this
// thread's share of [beg, end) of env e; slots past `end` (ragged last chunk only) hold +inf.
template <int D, bool kU16, typename F>
__device__ __forceinline__ void mirror_scan(const View &v, int e, int beg, int end, float qx, float qy, float qz, F &&visit) {
    constexpr int kVec = kU16 ? 8 : 4;
    const int step = kVec * blockDim.x;
    for (int i = beg + kVec * threadIdx.x; i < end; i += step) {
        float a[kVec];
        if (kU16) {
            const unsigned short *X = v.ux + (size_t)e * v.stride, *Y = v.uy + (size_t)e * v.stride;
            const unsigned short *Z = D == 3 ? v.uz + (size_t)e * v.stride : nullptr;
            const uint4 x = __ldcs(reinterpret_cast<const uint4 *>(X + i));
            const uint4 y = __ldcs(reinterpret_cast<const uint4 *>(Y + i));
            uint4 z = make_uint4(0u, 0u, 0u, 0u);
            if (D == 3) z = __ldcs(reinterpret_cast<const uint4 *>(Z + i));
#define MIRROR_U16(wx, wy, wz, j)                                                                   \
            {                                                                                       \
                const float dx0 = qx - __uint_as_float(__byte_perm(wx, 0x4B00u, 0x5410));           \
                const float dy0 = qy - __uint_as_float(__byte_perm(wy, 0x4B00u, 0x5410));           \
                const float dx1 = qx - __uint_as_float(__byte_perm(wx, 0x4B00u, 0x5432));           \
                const float dy1 = qy - __uint_as_float(__byte_perm(wy, 0x4B00u, 0x5432));           \
                a[j] = fmaf(dy0, dy0, dx0 * dx0); a[(j) + 1] = fmaf(dy1, dy1, dx1 * dx1);           \
                if (D == 3) {                                                                       \
                    const float dz0 = qz - __uint_as_float(__byte_perm(wz, 0x4B00u, 0x5410));       \
                    const float dz1 = qz - __uint_as_float(__byte_perm(wz, 0x4B00u, 0x5432));       \
                    a[j] = fmaf(dz0, dz0, a[j]); a[(j) + 1] = fmaf(dz1, dz1, a[(j) + 1]);           \
                }                                                                                   \
            }
            MIRROR_U16(x.x, y.x, z.x, 0)
            MIRROR_U16(x.y, y.y, z.y, 2)
            MIRROR_U16(x.z, y.z, z.z, 4)
            MIRROR_U16(x.w, y.w, z.w, 6)
#undef MIRROR_U16
        } else {
            const float *X = v.fx + (size_t)e * v.stride, *Y = v.fy + (size_t)e * v.stride;
            const float *Z = D == 3 ? v.fz + (size_t)e * v.stride : nullptr;
            const float4 x = __ldcs(reinterpret_cast<const float4 *>(X + i));
            const float4 y = __ldcs(reinterpret_cast<const float4 *>(Y + i));
            float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            if (D == 3) z = __ldcs(reinterpret_cast<const float4 *>(Z + i));
#define MIRROR_F32(xx, yy, zz, j)                                                                   \
            {                                                                                       \
                const float dx = qx - (xx), dy = qy - (yy);                                         \
                a[j] = fmaf(dy, dy, dx * dx);                                                       \
                if (D == 3) { const float dz = qz - (zz); a[j] = fmaf(dz, dz, a[j]); }              \
            }
            MIRROR_F32(x.x, y.x, z.x, 0)
            MIRROR_F32(x.y, y.y, z.y, 1)
            MIRROR_F32(x.z, y.z, z.z, 2)
            MIRROR_F32(x.w, y.w, z.w, 3)
#undef MIRROR_F32
        }
        if (i + kVec > end) {
#pragma unroll
            for (int j = 0; j < kVec; j++) if (i + j >= end) a[j] = INFINITY;
        }
        visit(a, i);
    }
}
// u16 mirror, software pipelined: the loads of iteration i + 1 are issued before iteration i is evaluated, so a warp
// always has one set of loads in flight while it computes (the plain loop alternates "wait for the loads" and "compute":
// per-SM throughput = bytes in flight / (latency + compute time); pipelined it is bytes in flight / max of the two).
template <int D, typename F>
__device__ __forceinline__ void mirror_scan_u16_pipelined(const View &v, int e, int beg, int end, float qx, float qy, float qz, F &&visit);

template <int D>
__device__ __forceinline__ void mirror_u16_vals(const uint4 &x, const uint4 &y, const uint4 &z, float qx, float qy, float qz, float (&a)[8]);
template <int N> __device__ __forceinline__ float vec_min(const float (&a)[N]) {
    float m = fminf(a[0], a[1]);
#pragma unroll
    for (int j = 2; j < N; j++) m = fminf(m, a[j]);
    return m;
}

template <int D, typename F>
__device__ __forceinline__ void mirror_scan_u16_pipelined(const View &v, int e, int beg, int end, float qx, float qy, float qz, F &&visit) {
    const unsigned short *X = v.ux + (size_t)e * v.stride, *Y = v.uy + (size_t)e * v.stride;
    const unsigned short *Z = D == 3 ? v.uz + (size_t)e * v.stride : nullptr;
    const int step = 8 * blockDim.x;
    int i = beg + 8 * threadIdx.x;
    if (i >= end) return;
    uint4 x = __ldcs(reinterpret_cast<const uint4 *>(X + i)), y = __ldcs(reinterpret_cast<const uint4 *>(Y + i));
    uint4 z = make_uint4(0u, 0u, 0u, 0u);
    if (D == 3) z = __ldcs(reinterpret_cast<const uint4 *>(Z + i));
    for (;;) {
        const int nx = i + step;
        uint4 x2 = x, y2 = y, z2 = z;
        if (nx < end) {
            x2 = __ldcs(reinterpret_cast<const uint4 *>(X + nx)); y2 = __ldcs(reinterpret_cast<const uint4 *>(Y + nx));
            if (D == 3) z2 = __ldcs(reinterpret_cast<const uint4 *>(Z + nx));
        }
        float a[8];
        mirror_u16_vals<D>(x, y, z, qx, qy, qz, a);
        if (i + 8 > end) {
#pragma unroll
            for (int j = 0; j < 8; j++) if (i + j >= end) a[j] = INFINITY;
        }
        visit(a, i);
        if (nx >= end) break;
        x = x2; y = y2; z = z2; i = nx;
    }
}

__device__ __forceinline__ void append_cand(const View &v, EnvCtl *c, int e, int idx) {
    const int slot = atomicAdd(&c->cand_cnt, 1);
    if (slot < v.near_cap) v.cand[(size_t)e * v.near_cap + slot] = idx;
}

// the mirror pass over vertices [beg, end) of problem e by the calling CTA (any block size): Nearest candidates of
// the range into the problem's candidate list, speculative Near members into its cand2 list
template <int D, bool kU16, bool kPipe>
__device__ __forceinline__ void nearest_m_range(const View &v, int e, EnvCtl *c, const ScanHdr &h, int beg, int end) {
    __shared__ unsigned s_min;
    if (threadIdx.x == 0) s_min = 0x7f800000u;
    __syncthreads();
    constexpr int kVec = kU16 ? 8 : 4;
    if (beg >= end) return;
    float a1 = INFINITY, a2 = INFINITY;   // best and second-best mirror value of this thread
    int i1 = INT_MAX;
    const float thr = h.thr;              // speculative Near ball around x_rand (see top_body)
    auto visit = [&](const float (&a)[kVec], int base) {
        const float m = vec_min(a);
        if (m < a2) {                     // rare once the running values have settled
#pragma unroll
            for (int j = 0; j < kVec; j++) {
                a2 = fminf(a2, fmaxf(a[j], a1));
                if (a[j] < a1) { a1 = a[j]; i1 = base + j; }
            }
        }
        if (m <= thr) {                   // ~|Near| of the env's vertices: cost proportional to the hits
            unsigned hit = 0;
#pragma unroll
            for (int j = 0; j < kVec; j++) hit |= (a[j] <= thr ? 1u : 0u) << j;
            while (hit) {
                const int j = __ffs(hit) - 1;
                hit &= hit - 1;
                const int slot = atomicAdd(&IT.spec_cnt, 1) - h.base;
                if (slot < v.near_cap) cand2_of(v, e)[slot] = base + j;
            }
        }
    };
    if constexpr (kPipe) mirror_scan_u16_pipelined<D>(v, e, beg, end, h.qx, h.qy, h.qz, visit);
    else mirror_scan<D, kU16>(v, e, beg, end, h.qx, h.qy, h.qz, visit);
    // a >= 0: the float order is the order of the bit patterns
    const unsigned wmin = __reduce_min_sync(0xffffffffu, __float_as_uint(a1));
    if ((threadIdx.x & 31) == 0) atomicMin(&s_min, wmin);
    __syncthreads();
    const float amin = __uint_as_float(s_min);
    // band in the distance domain, rounded up: sqrt(a) <= sqrt(amin) + 2 * margin
    const float lim = __fadd_ru(__fsqrt_ru(amin), h.band);
    const float band = __fmul_ru(__fmul_ru(lim, lim), 1.000001f);
    if (a1 <= band) {
        if (a2 > band) append_cand(v, c, e, i1);
        else mirror_scan<D, kU16>(v, e, beg, end, h.qx, h.qy, h.qz, [&](const float (&a)[kVec], int base) {
#pragma unroll
            for (int j = 0; j < kVec; j++) if (a[j] <= band) append_cand(v, c, e, base + j);
        });
    }
}

template <int D, bool kU16, bool kForce, bool kPipe = false>
__global__ void __launch_bounds__(256, kPipe ? 5 : 8) k_nearest_m(View v) {
    pdl_wait();
    pdl_launch_dependents();
    const int e = v.env0 + blockIdx.y;
    EnvCtl *c = v.ctl + e;
    const ScanHdr h = load_hdr(&c->hdr0);
    if (!kForce && !h.go) return;
    constexpr int kVec = kU16 ? 8 : 4;
    const int per = (((h.n + (int)gridDim.x - 1) / (int)gridDim.x) + kVec - 1) & ~(kVec - 1);
    const int beg = blockIdx.x * per;
    nearest_m_range<D, kU16, kPipe>(v, e, c, h, beg, min(h.n, beg + per));
}

// ---- TMA-staged u16 mirror scan (NIRRT_TMA=1..3; measured slower than the LDG kernel).  Same filter, same candidate logic, same results as k_nearest_m<D, true>;
// what changes is how the bytes reach the SM.  k_nearest_m issues three LDG.128 per thread and iteration and waits for
// them (a CTA's share is only ~5 iterations: every one of them exposes a full DRAM round trip -- long-scoreboard stalls
// dominate its profile, 0.80 of the copy peak).  Here one thread per CTA issues bulk copies (cp.async.bulk, completion on
// an mbarrier) of whole 2048-vertex tiles of the x / y / z arrays into a ring of shared-memory stages as soon as the
// scan header has arrived, so up to kTileStages * 12 KB per CTA are in flight before the first vertex is looked at and
// the compute only ever waits on shared memory.  Evict-first L2 policy: the mirror streams through once per iteration.
constexpr int kTileVerts = 2048;   // 256 threads x 8 vertices: 4 KB per coordinate array and stage

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
// bounded wait: a mis-programmed pipeline traps (the launch fails) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_addr(bar);
    for (uint32_t spins = 0;; spins++) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return;
        if (spins > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)), "l"(policy) : "memory");
}

// mirror squared distances of 8 consecutive vertices (see mirror_scan: cells carried as 2^23 + cell floats, exact subtraction)
template <int D>
__device__ __forceinline__ void mirror_u16_vals(const uint4 &x, const uint4 &y, const uint4 &z, float qx, float qy, float qz, float (&a)[8]) {
    const unsigned wx[4] = {x.x, x.y, x.z, x.w}, wy[4] = {y.x, y.y, y.z, y.w}, wz[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const float dx0 = qx - __uint_as_float(__byte_perm(wx[j], 0x4B00u, 0x5410));
        const float dy0 = qy - __uint_as_float(__byte_perm(wy[j], 0x4B00u, 0x5410));
        const float dx1 = qx - __uint_as_float(__byte_perm(wx[j], 0x4B00u, 0x5432));
        const float dy1 = qy - __uint_as_float(__byte_perm(wy[j], 0x4B00u, 0x5432));
        a[2 * j] = fmaf(dy0, dy0, dx0 * dx0); a[2 * j + 1] = fmaf(dy1, dy1, dx1 * dx1);
        if (D == 3) {
            const float dz0 = qz - __uint_as_float(__byte_perm(wz[j], 0x4B00u, 0x5410));
            const float dz1 = qz - __uint_as_float(__byte_perm(wz[j], 0x4B00u, 0x5432));
            a[2 * j] = fmaf(dz0, dz0, a[2 * j]); a[2 * j + 1] = fmaf(dz1, dz1, a[2 * j + 1]);
        }
    }
}

template <int D, bool kForce, int kTileStages, int kMinBlocks>
__global__ void __launch_bounds__(256, kMinBlocks) k_nearest_t(View v) {
    pdl_wait();
    pdl_launch_dependents();
    const int e = v.env0 + blockIdx.y;
    EnvCtl *c = v.ctl + e;
    const ScanHdr h = load_hdr(&c->hdr0);
    if (!kForce && !h.go) return;
    __shared__ __align__(128) unsigned short s_t[kTileStages][3][kTileVerts];
    __shared__ __align__(8) uint64_t s_full[kTileStages];
    __shared__ unsigned s_min;
    const int tid = threadIdx.x;
    const int per = (((h.n + (int)gridDim.x - 1) / (int)gridDim.x) + 7) & ~7;
    const int beg = blockIdx.x * per;
    const int end = min(h.n, beg + per);
    if (beg >= end) return;
    const int ntiles = (end - beg + kTileVerts - 1) / kTileVerts;
    const unsigned short *X = v.ux + (size_t)e * v.stride, *Y = v.uy + (size_t)e * v.stride;
    const unsigned short *Z = D == 3 ? v.uz + (size_t)e * v.stride : nullptr;
    uint64_t policy = 0;
    auto issue = [&](int t) {        // thread 0 only
        const int st = t % kTileStages, first = beg + t * kTileVerts;
        const uint32_t bytes = (uint32_t)((min(kTileVerts, end - first) * 2 + 15) & ~15);
        mbar_expect_tx(&s_full[st], D * bytes);
        bulk_load(s_t[st][0], X + first, bytes, &s_full[st], policy);
        bulk_load(s_t[st][1], Y + first, bytes, &s_full[st], policy);
        if (D == 3) bulk_load(s_t[st][2], Z + first, bytes, &s_full[st], policy);
    };
    if (tid == 0) {
        s_min = 0x7f800000u;
#pragma unroll
        for (int st = 0; st < kTileStages; st++) mbar_init(&s_full[st], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        policy = l2_evict_first_policy();
        for (int t = 0; t < min(kTileStages, ntiles); t++) issue(t);
    }
    __syncthreads();
    float a1 = INFINITY, a2 = INFINITY;   // best and second-best mirror value of this thread
    int i1 = INT_MAX;
    const float thr = h.thr;              // speculative Near ball around x_rand (see top_body)
    for (int t = 0; t < ntiles; t++) {
        const int st = t % kTileStages;
        mbar_wait(&s_full[st], (uint32_t)((t / kTileStages) & 1));
        const int base = beg + t * kTileVerts + tid * 8;
        if (base < end) {
            const uint4 x = *reinterpret_cast<const uint4 *>(&s_t[st][0][tid * 8]);
            const uint4 y = *reinterpret_cast<const uint4 *>(&s_t[st][1][tid * 8]);
            uint4 z = make_uint4(0u, 0u, 0u, 0u);
            if (D == 3) z = *reinterpret_cast<const uint4 *>(&s_t[st][2][tid * 8]);
            float a[8];
            mirror_u16_vals<D>(x, y, z, h.qx, h.qy, h.qz, a);
            if (base + 8 > end) {
#pragma unroll
                for (int j = 0; j < 8; j++) if (base + j >= end) a[j] = INFINITY;
            }
            const float m = vec_min(a);
            if (m < a2) {                     // rare once the running values have settled
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    a2 = fminf(a2, fmaxf(a[j], a1));
                    if (a[j] < a1) { a1 = a[j]; i1 = base + j; }
                }
            }
            if (m <= thr) {                   // ~|Near| of the env's vertices: cost proportional to the hits
                unsigned hit = 0;
#pragma unroll
                for (int j = 0; j < 8; j++) hit |= (a[j] <= thr ? 1u : 0u) << j;
                while (hit) {
                    const int j = __ffs(hit) - 1;
                    hit &= hit - 1;
                    const int slot = atomicAdd(&IT.spec_cnt, 1) - h.base;
                    if (slot < v.near_cap) cand2_of(v, e)[slot] = base + j;
                }
            }
        }
        if (t + kTileStages < ntiles) {       // the stage is free once every thread has read its vertices
            __syncthreads();
            if (tid == 0) issue(t + kTileStages);
        }
    }
    // a >= 0: the float order is the order of the bit patterns
    const unsigned wmin = __reduce_min_sync(0xffffffffu, __float_as_uint(a1));
    if ((tid & 31) == 0) atomicMin(&s_min, wmin);
    __syncthreads();
    const float amin = __uint_as_float(s_min);
    // band in the distance domain, rounded up: sqrt(a) <= sqrt(amin) + 2 * margin
    const float lim = __fadd_ru(__fsqrt_ru(amin), h.band);
    const float band = __fmul_ru(__fmul_ru(lim, lim), 1.000001f);
    if (a1 <= band) {
        if (a2 > band) append_cand(v, c, e, i1);
        else {      // rare: several of this thread's vertices inside the band -- re-read them (tiles t with base < end)
            for (int t = 0; t < ntiles; t++) {
                const int base = beg + t * kTileVerts + tid * 8;
                if (base >= end) break;
                const uint4 x = __ldcs(reinterpret_cast<const uint4 *>(X + base));
                const uint4 y = __ldcs(reinterpret_cast<const uint4 *>(Y + base));
                uint4 z = make_uint4(0u, 0u, 0u, 0u);
                if (D == 3) z = __ldcs(reinterpret_cast<const uint4 *>(Z + base));
                float a[8];
                mirror_u16_vals<D>(x, y, z, h.qx, h.qy, h.qz, a);
#pragma unroll
                for (int j = 0; j < 8; j++) if (base + j < end && a[j] <= band) append_cand(v, c, e, base + j);
            }
        }
    }
}

// ---- u8 mirror scan (NIRRT_SCAN=u8): 4 bytes per vertex, two DP4A + one IMAD per vertex.
// A CTA iteration covers 16 vertices per thread: four fully coalesced LDG.128 per thread (load u of lane l of
// warp w reads vertices i0 + w*512 + u*128 + l*4 .. +3).  visit(a[16], base): a[j] = |m|^2 - 2 m.q8 (exact
// integer, |q8|^2 is folded into the thresholds) of vertex base + (j >> 2) * 128 + (j & 3); INT_MAX past `end`.
__device__ __forceinline__ int byte_index(int base, int j) { return base + (j >> 2) * 128 + (j & 3); }
template <typename F>
__device__ __forceinline__ void byte_scan(const View &v, int e, int beg, int end, unsigned q, F &&visit) {
    const unsigned *M = v.m8 + (size_t)e * v.stride;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int step = 16 * blockDim.x;
    for (int i0 = beg; i0 < end; i0 += step) {
        const int base = i0 + warp * 512 + lane * 4;
        int a[16];
        if (i0 + step <= end) {           // the whole CTA tile lies inside the chunk: no per-vertex range checks
            uint4 w[4];
#pragma unroll
            for (int u = 0; u < 4; u++) w[u] = __ldcs(reinterpret_cast<const uint4 *>(M + base + u * 128));   // streaming: evict-first
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const unsigned m[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
#pragma unroll
                for (int k = 0; k < 4; k++) a[4 * u + k] = (int)__dp4a(m[k], m[k], 0u) - 2 * (int)__dp4a(m[k], q, 0u);
            }
        } else {
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = base + u * 128;
                uint4 w = make_uint4(0u, 0u, 0u, 0u);
                if (i < end) w = __ldcs(reinterpret_cast<const uint4 *>(M + i));
                const unsigned m[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                for (int k = 0; k < 4; k++)
                    a[4 * u + k] = (i + k < end) ? (int)__dp4a(m[k], m[k], 0u) - 2 * (int)__dp4a(m[k], q, 0u) : INT_MAX;
            }
        }
        visit(a, base);
    }
}
__device__ __forceinline__ int vec_min_int16(const int (&a)[16]) {
    int m = min(a[0], a[1]);
#pragma unroll
    for (int j = 2; j < 16; j++) m = min(m, a[j]);
    return m;
}

template <int D, bool kForce>
__global__ void __launch_bounds__(256, 5) k_nearest_b(View v) {
    pdl_wait();
    pdl_launch_dependents();
    const int e = v.env0 + blockIdx.y;
    EnvCtl *c = v.ctl + e;
    const ScanHdr h = load_hdr(&c->hdr0);
    if (!kForce && !h.go) return;
    __shared__ int s_min;
    if (threadIdx.x == 0) s_min = INT_MAX;
    __syncthreads();
    const int per = (((h.n + (int)gridDim.x - 1) / (int)gridDim.x) + 3) & ~3;
    const int beg = blockIdx.x * per;
    const int end = min(h.n, beg + per);
    if (beg >= end) return;
    const unsigned q = __float_as_uint(h.qx);
    const int qq = __float_as_int(h.qy), thr = __float_as_int(h.thr);
    int a1 = INT_MAX, a2 = INT_MAX, i1 = INT_MAX;     // best and second-best value of this thread (a' = d8^2 - |q8|^2)
    byte_scan(v, e, beg, end, q, [&](const int (&a)[16], int base) {
        const int m = vec_min_int16(a);
        if (m < a2) {                     // rare once the running values have settled
#pragma unroll
            for (int j = 0; j < 16; j++) {
                a2 = min(a2, max(a[j], a1));
                if (a[j] < a1) { a1 = a[j]; i1 = byte_index(base, j); }
            }
        }
        if (m <= thr) {                   // speculative Near ball around x_rand (see top_body)
            unsigned hit = 0;
#pragma unroll
            for (int j = 0; j < 16; j++) hit |= (a[j] <= thr ? 1u : 0u) << j;
            while (hit) {
                const int j = __ffs(hit) - 1;
                hit &= hit - 1;
                const int slot = atomicAdd(&IT.spec_cnt, 1) - h.base;
                if (slot < v.near_cap) cand2_of(v, e)[slot] = byte_index(base, j);
            }
        }
    });
    const int wmin = __reduce_min_sync(0xffffffffu, a1);
    if ((threadIdx.x & 31) == 0) atomicMin(&s_min, wmin);
    __syncthreads();
    // band in the distance domain, rounded up: sqrt(d8^2) <= sqrt(min d8^2) + 2 * margin
    const float lim = __fadd_ru(__fsqrt_ru((float)(s_min + qq)), h.band);
    const int band = (int)fminf(__fmul_ru(__fmul_ru(lim, lim), 1.000001f), 1.0e9f) + 1 - qq;
    if (a1 <= band) {
        if (a2 > band) append_cand(v, c, e, i1);
        else byte_scan(v, e, beg, end, q, [&](const int (&a)[16], int base) {
#pragma unroll
            for (int j = 0; j < 16; j++) if (a[j] <= band) append_cand(v, c, e, byte_index(base, j));
        });
    }
}

// ---- SAD scan (default): 4 bytes per vertex {x8, y8, z8, 0}, ONE instruction per vertex.
// VABSDIFF4 with accumulation sums |m_b - q_b| over the four bytes of a word: the L1 distance of the vertex to the
// query in cells.  L1 is a conservative stand-in for the L2 test (see sad_bound): the pass keeps every vertex whose L1
// distance is within the bound of the speculative Near ball -- staged in shared memory with the CTA-local counter,
// flushed with ONE global atomic per CTA -- and tracks the minimum; the Nearest candidates are the staged vertices
// within the bound derived from the smallest L1 distance ANY chunk of the problem has published so far (any vertex's
// distance is an upper bound of the nearest one's, so a stale value only widens the band).  Everything that passes is
// re-decided in exact float64 by k_expand, so the results are bit-identical to every other scan layout.
// Against the u16 pass: 4 instead of 6 bytes per vertex and ~3 instead of ~12 issued instructions per vertex -- the
// pass is bound by HBM, not by instruction issue -- at the price of a coarser filter (~2.5x the Near candidates and
// ~30 instead of ~12 Nearest candidates for k_expand's exact tests at 1e5 vertices).
constexpr int kHitCap = 1024;     // staged hits per CTA; more: the problem falls back to its own scans for this iteration
template <int D, bool kForce>
__global__ void __launch_bounds__(256, 6) k_nearest_s(View v) {
    pdl_wait();
    pdl_launch_dependents();
    const int e = v.env0 + blockIdx.y;
    EnvCtl *c = v.ctl + e;
    const ScanHdr h = load_hdr(&c->hdr0);
    if (!kForce && !h.go) return;
    __shared__ int s_cnt, s_min, s_band, s_slot;
    __shared__ int s_hidx[kHitCap];
    __shared__ unsigned short s_hs[kHitCap];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { s_cnt = 0; s_min = 0x3fffffff; }
    __syncthreads();
    const int per = (((h.n + (int)gridDim.x - 1) / (int)gridDim.x) + 3) & ~3;
    const int beg = blockIdx.x * per;
    const int end = min(h.n, beg + per);
    if (beg >= end) return;
    const unsigned q = __float_as_uint(h.qx);
    const int thr = __float_as_int(h.thr);
    const unsigned *M = v.m8 + (size_t)e * v.stride;
    int best = 0x3fffffff;
    auto one = [&](unsigned w, int idx) {
        const int s = (int)__vsadu4(w, q);
        best = min(best, s);
        if (s <= thr) {
            const int p = atomicAdd(&s_cnt, 1);
            if (p < kHitCap) { s_hidx[p] = idx; s_hs[p] = (unsigned short)s; }
        }
    };
    // a CTA iteration covers 16 vertices per thread: four coalesced LDG.128 per thread (warp w, load u, lane l reads
    // vertices i0 + w*512 + u*128 + l*4 .. +3), all in flight before the first is used
    const int step = 16 * 256;
    for (int i0 = beg; i0 < end; i0 += step) {
        const int base = i0 + warp * 512 + lane * 4;
        if (i0 + step <= end) {
            uint4 w[4];
#pragma unroll
            for (int u = 0; u < 4; u++) w[u] = __ldcs(reinterpret_cast<const uint4 *>(M + base + u * 128));
#pragma unroll
            for (int u = 0; u < 4; u++) {
                one(w[u].x, base + u * 128); one(w[u].y, base + u * 128 + 1); one(w[u].z, base + u * 128 + 2); one(w[u].w, base + u * 128 + 3);
            }
        } else {
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = base + u * 128;
                if (i < end) {
                    const uint4 w = __ldcs(reinterpret_cast<const uint4 *>(M + i));
                    one(w.x, i);
                    if (i + 1 < end) one(w.y, i + 1);
                    if (i + 2 < end) one(w.z, i + 2);
                    if (i + 3 < end) one(w.w, i + 3);
                }
            }
        }
    }
    const int wmin = __reduce_min_sync(0xffffffffu, best);
    if (lane == 0) atomicMin(&s_min, wmin);
    __syncthreads();
    const int nh_all = s_cnt;
    if (tid == 0) {
        // publish this chunk's minimum, take the problem-wide one (so far): any vertex's L1 distance bounds the nearest one's
        const int g = min(atomicMin(&c->scan_min, s_min), s_min);
        s_band = sad_bound(v, (float)(g + D));
        if (nh_all > kHitCap) {
            // too many hits to stage: make both candidate lists overflow -- the problem then re-scans on its own
            // (steer_body's exact Nearest fallback, k_expand's Near scan)
            atomicAdd(&IT.spec_cnt, v.near_cap + 1);
            atomicAdd(&c->cand_cnt, v.near_cap + 1);
            s_slot = -1;
        } else s_slot = nh_all ? atomicAdd(&IT.spec_cnt, nh_all) - h.base : 0;
    }
    __syncthreads();
    if (s_slot < 0) return;
    const int band = s_band, slot0 = s_slot;
    int *spec = cand2_of(v, e);
    for (int i = tid; i < nh_all; i += blockDim.x) {
        if (slot0 + i < v.near_cap) spec[slot0 + i] = s_hidx[i];
        if ((int)s_hs[i] <= band) append_cand(v, c, e, s_hidx[i]);
    }
    if (band > thr) {
        // the Nearest band reaches beyond the staged Near ball (sparse tree): collect it from the chunk directly
        for (int i = beg + 4 * tid; i < end; i += 4 * blockDim.x) {
            const uint4 w = __ldcs(reinterpret_cast<const uint4 *>(M + i));
            const unsigned ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int s = (int)__vsadu4(ww[j], q);
                if (i + j < end && s > thr && s <= band) append_cand(v, c, e, i + j);
            }
        }
    }
}

template <int D, bool kU16, bool kForce>
__global__ void __launch_bounds__(256) k_near_m(View v) {
    pdl_wait();
    pdl_launch_dependents();
    const int e = v.env0 + blockIdx.y;
    EnvCtl *c = v.ctl + e;
    const ScanHdr h = load_hdr(&IT.hdr1);
    if (!kForce && !h.go) return;
    constexpr int kVec = kU16 ? 8 : 4;
    const int per = (((h.n + (int)gridDim.x - 1) / (int)gridDim.x) + kVec - 1) & ~(kVec - 1);
    const int beg = blockIdx.x * per;
    const int end = min(h.n, beg + per);
    const float thr = h.thr;
    mirror_scan<D, kU16>(v, e, beg, end, h.qx, h.qy, h.qz, [&](const float (&a)[kVec], int base) {
        if (vec_min(a) <= thr) {
#pragma unroll
            for (int j = 0; j < kVec; j++) if (a[j] <= thr) append_cand(v, c, e, base + j);
        }
    });
}

// ------------------------------------------------------------------------------------------------
// k_expand: collision filter of find_near_neighbors (rrt_star_3d.py:141-144), choose_parent
// (:80-90), rewire (:92-99), InGoalRegion append (irrt_star_3d.py:70-71), search_goal_parent +
// path length record (rrt_star_3d.py:101-117,225-231), iteration accounting.
constexpr int kExpandThreads = 128;
constexpr int kNearSmem = 1024;   // Near candidates of one iteration staged in shared memory (the common case)
constexpr int kNearBig = 8192;    // upper limit with near_capacity > kNearSmem: per-problem HBM staging (View::big) --
                                  // dense informed trees (a thin ellipsoid holding thousands of vertices inside one Near ball)
constexpr int kPosBits = 13;      // bits of a Near-list position in the stamps / ancestor words (2^13 = kNearBig)
constexpr int kNearRow = 44;      // staging bytes per candidate: cand, near, par (int) + d, cost, first (double) + anc (u64)
constexpr int kNearRowAlloc = 45; // + the Rewire decision bits (1 bit per candidate, rounded up)
// ancestor word of a Near member: three positions, [3 * kPosBits, +2): their count, then: more than three, through x_new
constexpr int kAncCnt = 3 * kPosBits, kAncOver = kAncCnt + 2, kAncNew = kAncCnt + 3;

__device__ void bitonic_sort_int(int *a, int n_pow2) {
    for (int k = 2; k <= n_pow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const int x = a[i], y = a[ixj];
                    const bool up = ((i & k) == 0);
                    if ((x > y) == up) { a[i] = y; a[ixj] = x; }
                }
            }
            __syncthreads();
        }
    }
}

// ---- RRT* eval driver (planning_random of RRT* / NRRT*, rrt_star_3d.py:200-270): search_goal_parent (:101-117) +
// extract_path + get_path_len (rrt_base_3d.py:69-91) after EVERY iteration.
// The reference re-scans the tree for vertices within step_len of the goal and re-walks cost() for each of them
// (thousands at 1e5 vertices) every iteration.  Here:
//   * the candidate list (ascending vertex index == np.where order) is maintained incrementally: a vertex enters it
//     when it is inserted (coordinates never change), gc_d = its goal distance or +inf when the goal edge collides;
//   * gc_cost[k] caches cost(vertex_k) + gc_d[k].  cost() of a vertex changes only when the vertex or one of its
//     ancestors is re-parented, i.e. exactly for the subtrees below this iteration's re-wired vertices.  Every
//     vertex carries a child list (Kid: first child + doubly linked siblings, kept current on insert / ChooseParent /
//     Rewire), so those subtrees are enumerated (a few vertices on average) and only the goal candidates found in
//     them are re-walked -- with the reference's own leaf -> root summation, so the cached values ARE what
//     search_goal_parent would recompute;
//   * the arg-min is np.argmin over the list == lexicographic (value, slot) minimum: the previous winner stays the
//     minimum of the untouched candidates, so it is merged with the refreshed and the new ones; if the winner
//     itself was refreshed the cached array is re-reduced (coalesced, no walks);
//   * the path length only changes when the winner or its root path changed; otherwise the last value stands.
struct __align__(16) Kid {
    int head;    // first child, -1: leaf
    int next;    // next sibling
    int prev;    // previous sibling, -1: this vertex is its parent's first child
    int gslot;   // slot in the goal-candidate list if the vertex is a candidate with a free goal edge, else -1
};
__device__ __forceinline__ Kid load_kid(const Kid *p) {
    const int4 r = __ldcg(reinterpret_cast<const int4 *>(p));
    Kid k; k.head = r.x; k.next = r.y; k.prev = r.z; k.gslot = r.w;
    return k;
}
__device__ __forceinline__ void store_kid(Kid *p, int head, int next, int prev, int gslot) {
    __stcg(reinterpret_cast<int4 *>(p), make_int4(head, next, prev, gslot));
}
// thread-serial list surgery (one thread per problem; L2-coherent accesses, the traversal reads after a barrier)
__device__ __forceinline__ void kid_unlink(Kid *kid, int v, int old_parent) {
    const Kid k = load_kid(kid + v);
    if (k.prev >= 0) __stcg(&kid[k.prev].next, k.next); else __stcg(&kid[old_parent].head, k.next);
    if (k.next >= 0) __stcg(&kid[k.next].prev, k.prev);
}
// makes v the first child of `parent`, whose current first child is `head`
__device__ __forceinline__ void kid_link(Kid *kid, int v, int parent, int head) {
    __stcg(&kid[v].next, head); __stcg(&kid[v].prev, -1);
    if (head >= 0) __stcg(&kid[head].prev, v);
    __stcg(&kid[parent].head, v);
}

// The goal-candidate list of a problem: the RRT* eval driver's vertices within step_len of the goal (gc_idx, count
// n_goal), or the informed family's path_solutions (sol, count n_sol, irrt_star_3d.py:70-71 -- it may hold a vertex
// twice, through the duplicate guard; only the first slot of a vertex carries a cached value, the later ones +inf,
// which never changes the first-minimum rule).
struct GoalList { int *idx; double *d; double *cost; int *count; int cap; };
__device__ __forceinline__ GoalList goal_list(const View &v, EnvCtl *c, int e) {
    GoalList L;
    if (fam_informed(v.variant)) { L.idx = v.sol + (size_t)e * v.sol_cap; L.count = &c->n_sol; L.cap = v.sol_cap; }
    else { L.idx = v.gc_idx + (size_t)e * v.cap; L.count = &c->n_goal; L.cap = v.cap; }
    L.d = v.gc_d + (size_t)e * v.gc_stride; L.cost = v.gc_cost + (size_t)e * v.gc_stride;
    return L;
}

// get_path_len(extract_path(gp)) (rrt_base_3d.py:69-91): row norms root -> goal, numpy pairwise summation.  One thread.
template <int D>
__device__ double goal_path_len_from(const View &v, EnvCtl *c, int e, const Node *nodes, int gp) {
    int depth = 0;
    for (int i = gp; i != 0; i = (int)load_node(nodes + i).parent) depth++;
    const int M = depth + 1;
    if (M > v.path_cap) { atomicOr(&c->err, ERR_PATH_DEPTH); return XINF; }
    double *seg = v.pathseg + (size_t)e * v.path_cap;
    Node cur = load_node(nodes + gp);
    seg[M - 1] = row_norm<D>(XSUB(c->goal[0], cur.x), XSUB(c->goal[1], cur.y), XSUB(c->goal[2], cur.z));
    int idx = gp;
    for (int j = 0; idx != 0; j++) {
        const int par = (int)cur.parent;
        const Node p = load_node(nodes + par);
        seg[M - 2 - j] = row_norm<D>(XSUB(cur.x, p.x), XSUB(cur.y, p.y), XSUB(cur.z, p.z));
        idx = par; cur = p;
    }
    return (M == 1) ? seg[0] : XADD(seg[0], pairwise_sum(seg + 1, M - 1));
}

// Full evaluation: walks every candidate, refreshes the cache, arg-min, path length.  Called by all threads (any block
// size); result in thread 0, which also updates the problem's bookkeeping.
template <int D>
__device__ double goal_path_len(const View &v, EnvCtl *c, int e, const Node *nodes, double *sm_s, int *sm_i) {
    const GoalList L = goal_list(v, c, e);
    const int ng = min(*L.count, L.cap);
    if (ng == 0) {
        if (threadIdx.x == 0) { c->last_gp = -1; c->best_k = -1; c->best_val = XINF; c->last_len = XINF; }
        return XINF;
    }
    const int *gi = L.idx;
    const double *gd = L.d;
    double *gcst = L.cost;
    double bs = XINF; int bk = INT_MAX;
    for (int k = threadIdx.x; k < ng; k += blockDim.x) {
        const double d = gd[k];
        const double val = (d < XINF) ? XADD(cost_walk<D>(tree_of(v, e), gi[k]), d) : XINF;
        __stcg(gcst + k, val);
        lexmin(bs, bk, val, k);
    }
    block_lexmin(bs, bk, sm_s, sm_i);
    double len = XINF;
    if (threadIdx.x == 0) {
        const int gp = gi[bk];
        // the informed family records c_best itself (irrt_star_3d.py:80-93), not the numpy path length
        len = fam_informed(v.variant) ? bs : goal_path_len_from<D>(v, c, e, nodes, gp);
        c->last_gp = gp; c->best_k = bk; c->best_val = bs; c->last_len = len;
    }
    return len;
}

constexpr int kFrontMax = 512;   // subtree traversal frontier (two of them live in k_expand's candidate staging)
constexpr int kDirtyMax = 512;   // goal candidates refreshed individually per iteration; more: full evaluation

// One iteration's goal bookkeeping of the RRT* eval driver (see the comment above Kid).  Called by all threads of
// k_expand after ChooseParent / Rewire have been applied.
//   s_near / s_par / s_rew: Near list, each member's parent BEFORE this iteration, and the Rewire decisions (bit k)
//   par_new_old: parent of x_new's vertex before ChooseParent when it re-used an existing vertex (duplicate guard)
//   s_front (2 * kFrontMax ints) and s_dirty (kDirtyMax ints): scratch
template <int D, typename G>
__device__ void goal_track(const View &v, EnvCtl *c, int e, const G &g, const TreeRef &t, const Node *nodes,
                           int new_idx, bool inserted, bool new_moved, int m, const int *s_near, const int *s_par,
                           const unsigned *s_rew, int par_new_old, bool have_cnew, double c_new, const double *xnew,
                           int *s_front, int *s_dirty, double *sm_s, int *sm_i) {
    __shared__ int s_cnt[4];     // [0] next frontier size, [1] dirty candidates, [2] overflow, [3] slot of the new candidate / -1
    Kid *kid = v.kid + (size_t)e * v.stride;
    const GoalList L = goal_list(v, c, e);
    int *gi = L.idx;
    double *gd = L.d;
    double *gcst = L.cost;
    const bool informed = fam_informed(v.variant);
    const int tid = threadIdx.x;
    if (tid == 0) {
        int ns = 0, newk = -1, gslot_new = -1;
        bool ovf = false;
        const double gx = XSUB(c->goal[0], xnew[0]), gy = XSUB(c->goal[1], xnew[1]), gz = XSUB(c->goal[2], xnew[2]);
        if (informed) {
            // InGoalRegion (rrt_base_3d.py:93-95) -> path_solutions.append(node_new_index) (irrt_star_3d.py:70-71), also when
            // the duplicate guard re-used an existing vertex
            const double dg = edge_len<D>(gx, gy, gz);
            if (dg < c->step_len && !seg_collides(g, xnew, c->goal)) {
                const int k = c->n_sol;
                c->n_sol = k + 1;
                if (k < v.sol_cap) {
                    gi[k] = new_idx;
                    gd[k] = dg;
                    if (inserted || __ldcg(&kid[new_idx].gslot) < 0) { newk = k; gslot_new = k; }
                    else __stcg(gcst + k, XINF);      // a second slot of the same vertex: the first one carries the value
                } else atomicOr(&c->err, ERR_SOL_OVERFLOW);
            }
        } else if (inserted) {
            const double s2 = scan_sq<D>(gx, gy, gz);
            // dist_to_goal <= step_len (rrt_star_3d.py:103-104 / rrt_star_2d.py:103-104)
            if (s2 <= c->T_goal && (D == 3 || np_hypot(gx, gy) <= c->step_len)) {
                newk = c->n_goal;
                const bool hit = seg_collides(g, xnew, c->goal);
                gi[newk] = new_idx;
                gd[newk] = hit ? XINF : (D == 3 ? XSQRT(s2) : np_hypot(gx, gy));
                c->n_goal = newk + 1;
                if (!hit) gslot_new = newk;
            }
        }
        if (inserted) {
            // the new vertex under its final parent (after ChooseParent)
            const int pn = load_link(t.links + new_idx).parent;
            const int h = __ldcg(&kid[pn].head);
            store_kid(kid + new_idx, -1, h, -1, gslot_new);
            if (h >= 0) __stcg(&kid[h].prev, new_idx);
            __stcg(&kid[pn].head, new_idx);
        } else if (gslot_new >= 0) __stcg(&kid[new_idx].gslot, gslot_new);      // an existing vertex became a candidate
        if (!inserted && new_moved) {
            // the duplicate guard re-used an existing vertex and ChooseParent moved it: its whole subtree got cheaper
            const int pn = load_link(t.links + new_idx).parent;
            kid_unlink(kid, new_idx, par_new_old);
            kid_link(kid, new_idx, pn, __ldcg(&kid[pn].head));
            s_front[ns++] = new_idx | 0x40000000;
        }
        bool any = false;
        for (int w = 0; w < (m + 31) / 32 && !any; w++) any = s_rew[w] != 0u;
        if (any) {
            int head = -1;       // a vertex inserted in this iteration has no children yet
            for (int k = 0; k < m; k++)
                if ((s_rew[k >> 5] >> (k & 31)) & 1u) {
                    const int q = s_near[k];
                    kid_unlink(kid, q, s_par[k]);
                    // a re-used vertex (duplicate guard) may already be q's parent -- the unlink above then changed its
                    // child list, possibly its head: read it back instead of tracking it
                    if (!inserted) head = __ldcg(&kid[new_idx].head);
                    kid_link(kid, q, new_idx, head);
                    head = q;
                    if (ns < kFrontMax) s_front[ns++] = q | 0x40000000; else ovf = true;
                }
        }
        s_cnt[0] = ns; s_cnt[1] = 0; s_cnt[2] = ovf ? 1 : 0; s_cnt[3] = newk;
        c->work[6] += (unsigned long long)ns;
        __threadfence_block();
    }
    __syncthreads();
    // ---- the subtrees below the re-parented vertices (left-child / right-sibling walk, one level per round)
    int *cur = s_front, *nxt = s_front + kFrontMax;
    int ncur = s_cnt[0];
    for (int round = 0; ncur > 0 && !s_cnt[2]; round++) {
        if (round > v.cap) { if (tid == 0) { s_cnt[2] = 1; atomicOr(&c->err, ERR_CHILD_LISTS); } break; }   // never: a tree has < cap levels
        __syncthreads();
        if (tid == 0) s_cnt[0] = 0;
        __syncthreads();
        for (int i = tid; i < ncur; i += blockDim.x) {
            const int raw = cur[i], u = raw & 0x3fffffff;
            const Kid k = load_kid(kid + u);
            if (k.gslot >= 0) { const int p = atomicAdd(&s_cnt[1], 1); if (p < kDirtyMax) s_dirty[p] = k.gslot; }
            if (k.head >= 0) { const int p = atomicAdd(&s_cnt[0], 1); if (p < kFrontMax) nxt[p] = k.head; else s_cnt[2] = 1; }
            if (!(raw >> 30) && k.next >= 0) { const int p = atomicAdd(&s_cnt[0], 1); if (p < kFrontMax) nxt[p] = k.next; else s_cnt[2] = 1; }
        }
        __syncthreads();
        ncur = min(s_cnt[0], kFrontMax);
        int *tmp = cur; cur = nxt; nxt = tmp;
        if (tid == 0) c->work[3] += 1;
    }
    __syncthreads();
    const int nd = s_cnt[1], newk = s_cnt[3];
    if (tid == 0) c->work[5] += (unsigned long long)min(nd, kDirtyMax);
    if (s_cnt[2] || nd > kDirtyMax) {       // a huge subtree moved: evaluate everything (what the reference does every time)
        if (tid == 0) { c->work[4] += 1; c->work[7] += (unsigned long long)*L.count; }
        goal_path_len<D>(v, c, e, nodes, sm_s, sm_i);
        return;
    }
    const int old_k = c->best_k;
    const double old_val = c->best_val;
    if (nd == 0 && newk < 0) return;        // nothing that search_goal_parent looks at changed
    double bs = XINF; int bk = INT_MAX;
    bool best_dirty = false;
    for (int i = tid; i < nd; i += blockDim.x) {
        const int k = s_dirty[i];
        const double val = XADD(cost_walk<D>(t, gi[k]), gd[k]);
        __stcg(gcst + k, val);
        lexmin(bs, bk, val, k);
        best_dirty = best_dirty || k == old_k;
    }
    if (tid == 0 && newk >= 0) {
        const double dn = gd[newk];
        // c_new == cost(new_idx) as cost() would walk it now (no Near member: nobody walked it yet)
        const double val = dn < XINF ? XADD(have_cnew ? c_new : cost_walk<D>(t, new_idx), dn) : XINF;
        __stcg(gcst + newk, val);
        lexmin(bs, bk, val, newk);
    }
    const int any_best_dirty = __syncthreads_or(best_dirty);
    if (any_best_dirty) {
        __threadfence_block();
        __syncthreads();
        const int ng = min(*L.count, L.cap);
        bs = XINF; bk = INT_MAX;
        for (int k = tid; k < ng; k += blockDim.x) lexmin(bs, bk, __ldcg(gcst + k), k);
    }
    block_lexmin(bs, bk, sm_s, sm_i);
    if (tid == 0) {
        if (!any_best_dirty && old_k >= 0) lexmin(bs, bk, old_val, old_k);
        if (informed) { c->last_gp = gi[bk]; c->last_len = bs; }
        else if (any_best_dirty || bk != old_k) {
            const int gp = gi[bk];
            c->last_gp = gp;
            c->last_len = goal_path_len_from<D>(v, c, e, nodes, gp);
        }
        c->best_k = bk; c->best_val = bs;
    }
}

// g: the CTA's shared-memory copy of the obstacle table (g_ready: already staged by the caller); s_buf: the CTA's
// candidate staging (kNearSmem * kNearRow bytes of shared memory); presorted >= 0: the caller's own scan left that many
// speculative Near candidates in ascending index order at the start of s_buf (persistent kernel), -1: lists are in HBM
template <int D>
__device__ __forceinline__ void expand_iteration(const View &v, const int e, typename GeomOf<D>::type &g, const bool g_ready,
                                                 unsigned char *s_buf, const int presorted) {
    typedef typename GeomOf<D>::type G;
    EnvCtl *c = v.ctl + e;
#ifdef NIRRT_PHASE_TIMING
    const long long t_entry = clock64();
#endif
    if (v.fuse_steer) {
        if (!__ldcg(&c->hdr0.go)) return;
        if (threadIdx.x < 32) steer_body<D>(v, e, g_ready ? &g : nullptr);
        __syncthreads();
    } else if (!IT.go) return;
    // per-candidate staging, kNearRow bytes each: shared memory for up to kNearSmem candidates, else the problem's HBM
    // staging area (same layout, capacity near_cap) after the index sort, which always runs in shared memory
    int stage_cap = kNearSmem;
    unsigned char *stage = s_buf;
#define STAGE_INT(slot) (reinterpret_cast<int *>(stage) + (size_t)(slot) * stage_cap)
#define STAGE_F64(slot) (reinterpret_cast<double *>(stage + (size_t)12 * stage_cap) + (size_t)(slot) * stage_cap)
    int *s_cand = STAGE_INT(0);
    int *s_near = STAGE_INT(1);
    int *s_par = STAGE_INT(2);                          // parent(near_k) before this iteration (RRT* eval driver: child-list surgery)
    double *s_d = STAGE_F64(0);
    double *s_cost = STAGE_F64(1);                      // cost(near_k) before any rewiring of this iteration
    double *s_first = STAGE_F64(2);                     // Line(near_k, x_new) by math.hypot: first term of cost(x_new) via near_k, and the edge length if near_k is re-wired
    unsigned long long *s_anc = reinterpret_cast<unsigned long long *>(STAGE_F64(3));   // Near members on near_k's root path (see below)
    __shared__ unsigned s_rew_small[kNearSmem / 32];
    unsigned *s_rew = s_rew_small;                      // Rewire decisions, one bit per Near position
    int rew_words = kNearSmem / 32;
    __shared__ int s_par_new;                           // parent of the re-used vertex before ChooseParent (duplicate guard)
    __shared__ double s_curr[3];                        // curr_node_new_cost, cost(new) via the steer parent, cost(new) via ChooseParent's winner
    __shared__ Hint s_hnew;                             // ancestor hints of x_new after ChooseParent
    __shared__ double sm_s[4];
    __shared__ int sm_i[4];
    __shared__ int s_m;
    Node *nodes = v.nodes + (size_t)e * v.stride;
    const int tid = threadIdx.x;
#ifdef NIRRT_PHASE_TIMING
    long long t_ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define PHASE_MARK(k) t_ph[k] = clock64();
#else
#define PHASE_MARK(k)
#endif
    PHASE_MARK(0)
    const bool skipped = IT.skip;

    if (!skipped) {
        if (!g_ready) stage_geom<D>(&g, v, e);
        const int *cand = v.cand + (size_t)e * v.near_cap;
        int cnt;
        bool sorted = false;
        if (IT.need_scan) {
            // Near scan by this CTA (x_new != x_rand: sparse tree or the duplicate guard) -- same filter as the
            // stand-alone k_near_m; the matches replace the (useless) speculative list of this iteration
            const ScanHdr h = load_hdr(&IT.hdr1);
            const float thr = h.thr;
            int *list = cand2_of(v, e);
            if (tid == 0) atomicExch(&IT.fb_cnt, 0);
            __syncthreads();
            auto add = [&](int idx) { const int slot = atomicAdd(&IT.fb_cnt, 1); if (slot < v.near_cap) list[slot] = idx; };
            if (v.m8) {
                const int thr8 = __float_as_int(h.thr);
                byte_scan(v, e, 0, h.n, __float_as_uint(h.qx), [&](const int (&a)[16], int base) {
                    if (vec_min_int16(a) <= thr8) {
#pragma unroll
                        for (int j = 0; j < 16; j++) if (a[j] <= thr8) add(byte_index(base, j));
                    }
                });
            } else if (v.ux) mirror_scan<D, true>(v, e, 0, h.n, h.qx, h.qy, h.qz, [&](const float (&a)[8], int base) {
                if (vec_min(a) <= thr) {
#pragma unroll
                    for (int j = 0; j < 8; j++) if (a[j] <= thr) add(base + j);
                }
            });
            else mirror_scan<D, false>(v, e, 0, h.n, h.qx, h.qy, h.qz, [&](const float (&a)[4], int base) {
                if (vec_min(a) <= thr) {
#pragma unroll
                    for (int j = 0; j < 4; j++) if (a[j] <= thr) add(base + j);
                }
            });
            __syncthreads();
            cnt = __ldcg(&IT.fb_cnt);
            cand = list;
        } else if (IT.use_spec) {
            cnt = __ldcg(&IT.spec_cnt) - IT.spec_base;
            cand = cand2_of(v, e);
            if (presorted >= 0) { cnt = presorted; sorted = true; }     // already in s_cand, ascending
        } else cnt = __ldcg(&c->cand_cnt);
        {
            const int limit = v.big ? min(v.near_cap, kNearBig) : kNearSmem;
            if (cnt > limit) {
                if (tid == 0) atomicOr(&c->err, ERR_NEAR_OVERFLOW);
                cnt = limit;
            }
        }
        if (sorted) {
            if (tid == 0) s_m = 0;
            __syncthreads();
        } else if (cnt > kNearSmem) {
            // More candidates than the shared staging holds (dense informed tree): sort the indices in shared memory
            // (the whole staging buffer as keys), then continue with every per-candidate array in HBM.
            int p2 = 1;
            while (p2 < cnt) p2 <<= 1;
            int *keys = reinterpret_cast<int *>(s_buf);        // kNearSmem * kNearRow / 4 = 11264 >= kNearBig ints
            for (int i = tid; i < p2; i += blockDim.x) keys[i] = i < cnt ? __ldcg(cand + i) : INT_MAX;
            if (tid == 0) s_m = 0;
            __syncthreads();
            bitonic_sort_int(keys, p2);
            stage_cap = v.near_cap;
            stage = v.big + (size_t)e * v.near_cap * kNearRowAlloc;
            s_cand = STAGE_INT(0); s_near = STAGE_INT(1); s_par = STAGE_INT(2);
            s_d = STAGE_F64(0); s_cost = STAGE_F64(1); s_first = STAGE_F64(2);
            s_anc = reinterpret_cast<unsigned long long *>(STAGE_F64(3));
            s_rew = reinterpret_cast<unsigned *>(stage + (size_t)kNearRow * stage_cap); rew_words = stage_cap / 32;
            for (int i = tid; i < cnt; i += blockDim.x) s_cand[i] = keys[i];
            __threadfence_block();
            __syncthreads();
        } else if (cnt <= 256) {
            // ascending order by rank counting (the indices are distinct): one barrier instead of a sorting network
            int *s_raw = s_near;                          // scratch until the compaction below fills s_near
            for (int i = tid; i < cnt; i += blockDim.x) s_raw[i] = __ldcg(cand + i);
            if (tid == 0) s_m = 0;
            __syncthreads();
            for (int i = tid; i < cnt; i += blockDim.x) {
                const int mine = s_raw[i];
                int rank = 0;
                for (int j = 0; j < cnt; j++) rank += s_raw[j] < mine;
                s_cand[rank] = mine;
            }
            __syncthreads();
        } else {
            int p2 = 1;
            while (p2 < cnt) p2 <<= 1;
            for (int i = tid; i < p2; i += blockDim.x) s_cand[i] = i < cnt ? __ldcg(cand + i) : INT_MAX;
            if (tid == 0) s_m = 0;
            __syncthreads();
            bitonic_sort_int(s_cand, p2);
        }
        PHASE_MARK(1)

        const double xnew[3] = {IT.x_new[0], IT.x_new[1], IT.x_new[2]};
        const int new_idx = IT.new_idx;
        const double T_near = IT.T_near, r_near = IT.r;
        // collision filter + ordered compaction (tiles of blockDim candidates, ascending)
        for (int base = 0; base < cnt; base += blockDim.x) {
            const int k = base + tid;
            bool keep = false;
            int idx = -1;
            double d = 0.0, first = 0.0;
            if (k < cnt) {
                idx = s_cand[k];
                const Node nd = load_node(nodes + idx);
                const double p1[3] = {nd.x, nd.y, nd.z};
                const double ex = XSUB(xnew[0], p1[0]), ey = XSUB(xnew[1], p1[1]), ez = XSUB(xnew[2], p1[2]);
                d = vec_dist<D>(ex, ey, ez);
                // dist <= r exactly as the reference decides it (the scan may have over-selected)
                const bool within = D == 3 ? (sq3_rows(ex, ey, ez) <= T_near) : (d <= r_near);
                keep = within && (idx != new_idx) && !seg_collides(g, xnew, p1);
                if (keep) first = edge_len<D>(ex, ey, ez);      // Line(near_k, x_new): first term of cost(x_new) via near_k
            }
            // block-ordered positions: warp ballots + warp-count prefix through smem
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            const int w = tid >> 5, l = tid & 31;
            __shared__ int s_wcnt[kExpandThreads / 32];
            if (l == 0) s_wcnt[w] = __popc(bal);
            __syncthreads();
            int off = s_m;
            for (int q = 0; q < w; q++) off += s_wcnt[q];
            if (keep) {
                const int pos = off + __popc(bal & ((1u << l) - 1u));
                s_near[pos] = idx;
                s_d[pos] = d;
                s_first[pos] = first;
            }
            __syncthreads();
            if (tid == 0) { int t = 0; for (int q = 0; q < kExpandThreads / 32; q++) t += s_wcnt[q]; s_m += t; }
            __syncthreads();
        }
        const int m = s_m;
        PHASE_MARK(2)
        int *near_out = v.near_out + (size_t)e * v.near_cap;
        for (int k = tid; k < m; k += blockDim.x) near_out[k] = s_near[k];
        if (tid == 0) { IT.near_cnt = m; c->work[0] += 1; c->work[1] += (unsigned long long)m; c->work[2] += (unsigned long long)cnt; }

        // the informed family needs c_best every iteration (its sampler); the RRT* family only in the eval driver
        const bool track = fam_informed(v.variant) || v.mode == NIRRT_MODE_PLANNING_RANDOM;
        double c_new_final = 0.0;          // cost(new_idx) after ChooseParent (valid when m > 0)
        bool moved_final = false;          // the duplicate guard's vertex was re-parented by ChooseParent
        if (m > 0) {
            // One parallel round of root walks serves ChooseParent, node_new_cost and Rewire.
            // Every walk also notes which OTHER Near members (and x_new itself when it re-used an
            // existing vertex) lie on its path: Rewire is sequential in the reference
            // (rrt_star_3d.py:96-99), a neighbour's cost only changes when one of its ancestors was
            // re-parented earlier in the same loop, and only then is it walked again.
            for (int i = tid; i < rew_words; i += blockDim.x) s_rew[i] = 0;
            // stamp the Near members in their walk records: a walk recognises them from the record it loads anyway
            const TreeRef t = tree_of(v, e);
            const int tag = (int)(c->stamp % ((1u << (31 - kPosBits)) - 1u)) + 1;   // (tag << kPosBits) | k stays a non-negative int
            for (int k = tid; k < m; k += blockDim.x)
                __stcg(reinterpret_cast<int *>(t.links + s_near[k]) + 3, (tag << kPosBits) | k);
            __syncthreads();
            double bs = XINF, via_best = 0.0; int bk = INT_MAX;
            PHASE_MARK(6)
            for (int k = tid; k <= m; k += blockDim.x) {
                if (k == m) {
                    // the steer parent: curr_node_new_cost = cost(nearest) + Line(nearest, new)
                    // (rrt_star_3d.py:46,51) and the cost(new) ChooseParent falls back to
                    const double e0 = IT.cnew_default;
                    double cn, via;
                    if (e0 < 0.0) {
                        int fp = -1;
                        cn = 0.0;
                        walk_to_root(t, IT.nearest, [&](int par, double eg, int) { if (fp < 0) fp = par; cn = XADD(cn, eg); });
                        s_curr[0] = cn; s_curr[1] = cn; s_par_new = fp;
                    }
                    else { cost_walk2<D>(t, IT.nearest, e0, cn, via); s_curr[0] = XADD(cn, e0); s_curr[1] = via; }
                    continue;
                }
                const int idx = s_near[k];
                double cacc = 0.0, vacc = s_first[k];
                unsigned long long anc = 0;   // three kPosBits-bit positions, their count, overflow flag, through-x_new flag (kAnc*)
                int fpar = -1;
                walk_to_root(t, idx, [&](int par, double eg, int pad) {
                    cacc = XADD(cacc, eg); vacc = XADD(vacc, eg);
                    if (fpar < 0) fpar = par;
                    if (par == new_idx) anc |= 1ull << kAncNew;
                    else if ((pad >> kPosBits) == tag) {      // a stale or aliased stamp can only add a (harmless) dependency
                        const int pos = pad & (kNearBig - 1);
                        const int cntp = (int)((anc >> kAncCnt) & 3ull);
                        if (cntp < 3) { anc |= (unsigned long long)pos << (kPosBits * cntp); anc = (anc & ~(3ull << kAncCnt)) | ((unsigned long long)(cntp + 1) << kAncCnt); }
                        else anc |= 1ull << kAncOver;
                    }
                });
                s_cost[k] = cacc; s_anc[k] = anc; s_par[k] = fpar;
                const double via_cost = XADD(cacc, s_d[k]);
                if (via_cost < bs) { bs = via_cost; bk = k; via_best = vacc; }   // k ascends per thread: first minimum
            }
            const int my_bk = bk;
            PHASE_MARK(7)
            block_lexmin(bs, bk, sm_s, sm_i);      // result broadcast to every thread
            if (my_bk == bk) s_curr[2] = via_best; // cost(x_new) if the winner becomes its parent (one walk, two sums)
            __syncthreads();
            PHASE_MARK(3)
            // ---- choose_parent (rrt_star_3d.py:80-90)
            const bool reparent = bs < s_curr[0];
            const double c_new = reparent ? s_curr[2] : s_curr[1];
            const bool new_moved = reparent && !IT.inserted;   // an existing vertex (duplicate guard) changed its parent
            c_new_final = c_new; moved_final = new_moved;
            if (tid == 0) {
                c->stamp++;
                Hint hnew;
                if (reparent) {
                    const int q = s_near[bk];
                    set_parent(t, new_idx, q, s_first[bk]);    // s_first[bk] = Line(near_bk, x_new) == the new edge
                    const Hint hq = load_hint(t.hints + q);
                    hnew.a[0] = q;
#pragma unroll
                    for (int w = 1; w < kHintHops; w++) hnew.a[w] = hq.a[w - 1];
                    store_hint_raw(t.hints + new_idx, hnew);
                    c->tree_changed = 1;
                } else hnew = load_hint(t.hints + new_idx);
                s_hnew = hnew;
            }
            // ---- rewire (rrt_star_3d.py:92-99): ascending index order, later neighbours see earlier
            // re-parentings.  A neighbour's cost can only change inside the loop if one of ITS ancestors
            // is a Near member at an earlier position that gets re-parented (or x_new itself moved), so:
            // (1) every thread decides its neighbours on the pre-loop costs; (2) if no neighbour has such
            // an ancestor with a positive decision, those decisions ARE the sequential result (induction
            // over the position) and the stores are applied in parallel; (3) otherwise thread 0 replays
            // the loop sequentially, re-walking exactly the neighbours the reference would see changed.
            for (int k0 = 0; k0 < m; k0 += blockDim.x) {
                const int k = k0 + tid;
                const bool dec = k < m && s_cost[k] > XADD(c_new, s_d[k]);
                const unsigned bal = __ballot_sync(0xffffffffu, dec);
                if ((tid & 31) == 0 && (k >> 5) < rew_words) s_rew[k >> 5] = bal;
            }
            __syncthreads();
            bool dep = false;
            bool any_dec = false;
            for (int w = 0; w < (m + 31) / 32; w++) any_dec = any_dec || s_rew[w] != 0u;
            for (int k = tid; k < m; k += blockDim.x) {
                const unsigned long long anc = s_anc[k];
                if (anc == 0ull) continue;
                if (new_moved && ((anc >> kAncNew) & 1ull)) dep = true;
                if (any_dec && ((anc >> kAncOver) & 1ull)) dep = true;
                const int cntp = (int)((anc >> kAncCnt) & 3ull);
                for (int q = 0; q < cntp; q++) {
                    const int pos = (int)((anc >> (kPosBits * q)) & (unsigned long long)(kNearBig - 1));
                    if (pos < k && ((s_rew[pos >> 5] >> (pos & 31)) & 1u)) dep = true;
                }
            }
            const int any_dep = __syncthreads_or(dep);
            const Hint hnew = s_hnew;
            if (!any_dep) {
                for (int k = tid; k < m; k += blockDim.x)
                    if ((s_rew[k >> 5] >> (k & 31)) & 1u) {
                        set_parent(t, s_near[k], new_idx, s_first[k]);
                        store_hint(t.hints + s_near[k], new_idx, hnew);
                    }
                if (tid == 0 && any_dec) c->tree_changed = 1;
            } else {
                __syncthreads();
                if (tid == 0) {
                    for (int w = 0; w < rew_words; w++) s_rew[w] = 0;
                    bool any = false;
                    for (int k = 0; k < m; k++) {
                        const unsigned long long anc = s_anc[k];
                        bool dirty = (new_moved && ((anc >> kAncNew) & 1ull)) || (any && ((anc >> kAncOver) & 1ull));
                        const int cntp = (int)((anc >> kAncCnt) & 3ull);
                        for (int q = 0; q < cntp && !dirty; q++) {
                            const int pos = (int)((anc >> (kPosBits * q)) & (unsigned long long)(kNearBig - 1));
                            dirty = (s_rew[pos >> 5] >> (pos & 31)) & 1u;
                        }
                        const double ck = dirty ? cost_walk<D>(t, s_near[k]) : s_cost[k];
                        if (ck > XADD(c_new, s_d[k])) {
                            set_parent(t, s_near[k], new_idx, s_first[k]);
                            store_hint(t.hints + s_near[k], new_idx, hnew);
                            s_rew[k >> 5] |= 1u << (k & 31);
                            any = true;
                            c->tree_changed = 1;
                        }
                    }
                }
            }
            __threadfence_block();
            __syncthreads();
        }
        PHASE_MARK(4)
        // ---- goal bookkeeping: path_solutions / goal candidates, their cached costs, the current best (goal_track)
        if (track) {
            __syncthreads();
            goal_track<D>(v, c, e, g, tree_of(v, e), nodes, new_idx, IT.inserted != 0, moved_final, m, s_near, s_par, s_rew,
                          s_par_new, m > 0, c_new_final, xnew, s_cand, reinterpret_cast<int *>(s_anc), sm_s, sm_i);
        }
        __syncthreads();
    }
    PHASE_MARK(5)

    // ---- per-iteration record + phase machine (RRT* family; the IRRT* family records in k_top)
    if (!fam_informed(v.variant) && v.mode == NIRRT_MODE_PLANNING_RANDOM) {
        if (tid == 0) {
            const double len = c->last_len;         // kept current by goal_track (k_goal_init at begin)
            push_record(v, c, e, len);
            if (c->state == ST_PHASE1) {
                c->p1_done++;
                if (len < c->cfg.stop_below) { c->state = ST_PHASE2; c->left = c->cfg.iter_after; if (c->left <= 0) c->state = ST_DONE; }
                else if (c->p1_done >= c->cfg.iter_max) c->state = ST_DONE;
            } else {
                c->left--;
                if (c->left <= 0) c->state = ST_DONE;
            }
            c->budget--;
        }
    } else if (tid == 0 && !v.pipe) {
        c->budget--;
        if (v.mode == NIRRT_MODE_PLANNING) {
            c->p1_done++;
            if (c->p1_done >= c->cfg.iter_max) c->state = ST_DONE;
        } else {  // IRRT* family planning_random: counters only, phase switches happen in k_top
            if (c->state == ST_PHASE1) c->p1_done++; else c->left--;
        }
    }
#ifdef NIRRT_PHASE_TIMING
    const long long t_rec = clock64();
#endif
    if (v.fuse_top) {   // the next iteration's driver step + sample
        __syncthreads();
        top_body<D>(v, e, g, g_ready || !skipped, sm_s, sm_i);
    }
#ifdef NIRRT_PHASE_TIMING
    if (tid == 0 && !skipped) {
        const long long t_end = clock64();
        atomicAdd(&g_phase[0], (unsigned long long)(t_ph[0] - t_entry));
        atomicAdd(&g_phase[1], (unsigned long long)(t_ph[1] - t_ph[0]));
        atomicAdd(&g_phase[2], (unsigned long long)(t_ph[2] - t_ph[1]));
        atomicAdd(&g_phase[3], (unsigned long long)(t_ph[3] > t_ph[2] ? t_ph[3] - t_ph[2] : 0));
        atomicAdd(&g_phase[4], (unsigned long long)(t_ph[3] > t_ph[2] ? t_ph[4] - t_ph[3] : t_ph[4] - t_ph[2]));
        atomicAdd(&g_phase[5], (unsigned long long)(t_ph[5] - t_ph[4]));
        atomicAdd(&g_phase[6], (unsigned long long)(t_rec - t_ph[5]));
        atomicAdd(&g_phase[7], (unsigned long long)(t_end - t_rec));
        atomicAdd(&g_phase[8], 1ull);
    }
#endif
}

template <int D>
__global__ void __launch_bounds__(kExpandThreads, 4) k_expand(View v) {
    __shared__ typename GeomOf<D>::type g;
    __shared__ __align__(16) unsigned char s_buf[kNearSmem * kNearRow];
    pdl_wait();
    expand_iteration<D>(v, v.env0 + blockIdx.x, g, false, s_buf, -1);
}

// The persistent kernel's own scan: the same mirror filter as nearest_m_range over the whole tree, with the warps
// working on contiguous quarters of the vertex range so that the speculative Near members come out in ascending index
// order without any sort -- each warp compacts its hits with a prefix over its lanes into its own shared-memory
// segment, and the segments are concatenated in warp order at the start of s_buf (where expand_iteration keeps its
// sorted candidate indices).  Returns their number, or -1 when they do not fit into the shared staging (the list is
// then in the problem's HBM list, still ordered).  Nearest candidates go to the HBM candidate list as usual.
constexpr int kSegCap = 1024;
template <int D>
__device__ __forceinline__ int scan_own_sorted(const View &v, int e, EnvCtl *c, const ScanHdr &h, unsigned char *s_buf) {
    __shared__ unsigned s_min;
    __shared__ int s_wcnt[kExpandThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    if (tid == 0) s_min = 0x7f800000u;
    __syncthreads();
    int *s_cand = reinterpret_cast<int *>(s_buf);
    int *seg = s_cand + kNearSmem + warp * kSegCap;
    const int per = ((h.n + nw - 1) / nw + 255) & ~255;
    const int beg = warp * per, end = min(h.n, beg + per);
    const unsigned short *X = v.ux + (size_t)e * v.stride, *Y = v.uy + (size_t)e * v.stride;
    const unsigned short *Z = D == 3 ? v.uz + (size_t)e * v.stride : nullptr;
    float a1 = INFINITY, a2 = INFINITY;
    int i1 = INT_MAX, wc = 0;
    const float thr = h.thr;
    for (int b0 = beg; b0 < end; b0 += 256) {          // warp-uniform trip count
        const int base = b0 + 8 * lane;
        float a[8];
        unsigned hit = 0;
        if (base < end) {
            const uint4 x = __ldcs(reinterpret_cast<const uint4 *>(X + base)), y = __ldcs(reinterpret_cast<const uint4 *>(Y + base));
            uint4 z = make_uint4(0u, 0u, 0u, 0u);
            if (D == 3) z = __ldcs(reinterpret_cast<const uint4 *>(Z + base));
            mirror_u16_vals<D>(x, y, z, h.qx, h.qy, h.qz, a);
            if (base + 8 > end) {
#pragma unroll
                for (int j = 0; j < 8; j++) if (base + j >= end) a[j] = INFINITY;
            }
            const float m = vec_min(a);
            if (m < a2) {
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    a2 = fminf(a2, fmaxf(a[j], a1));
                    if (a[j] < a1) { a1 = a[j]; i1 = base + j; }
                }
            }
            if (m <= thr) {
#pragma unroll
                for (int j = 0; j < 8; j++) hit |= (a[j] <= thr ? 1u : 0u) << j;
            }
        }
        if (__any_sync(0xffffffffu, hit != 0u)) {
            const int mine = __popc(hit);
            int incl = mine;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, off);
                if (lane >= off) incl += t;
            }
            int pos = wc + incl - mine;
            while (hit) {
                const int j = __ffs(hit) - 1;
                hit &= hit - 1;
                if (pos < kSegCap) seg[pos] = base + j;
                pos++;
            }
            wc += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
    const unsigned wmin = __reduce_min_sync(0xffffffffu, __float_as_uint(a1));
    if (lane == 0) { atomicMin(&s_min, wmin); s_wcnt[warp] = wc; }
    __syncthreads();
    // ---- Nearest candidates (as nearest_m_range)
    const float amin = __uint_as_float(s_min);
    const float lim = __fadd_ru(__fsqrt_ru(amin), h.band);
    const float band = __fmul_ru(__fmul_ru(lim, lim), 1.000001f);
    if (a1 <= band) {
        if (a2 > band) append_cand(v, c, e, i1);
        else {
            for (int b0 = beg; b0 < end; b0 += 256) {
                const int base = b0 + 8 * lane;
                if (base >= end) continue;
                const uint4 x = __ldcs(reinterpret_cast<const uint4 *>(X + base)), y = __ldcs(reinterpret_cast<const uint4 *>(Y + base));
                uint4 z = make_uint4(0u, 0u, 0u, 0u);
                if (D == 3) z = __ldcs(reinterpret_cast<const uint4 *>(Z + base));
                float a[8];
                mirror_u16_vals<D>(x, y, z, h.qx, h.qy, h.qz, a);
#pragma unroll
                for (int j = 0; j < 8; j++) if (base + j < end && a[j] <= band) append_cand(v, c, e, base + j);
            }
        }
    }
    // ---- the speculative Near list: warp segments concatenated in warp order == ascending index order
    int off = 0, total = 0;
    bool overflow = false;
    for (int w = 0; w < nw; w++) {
        const int cw = s_wcnt[w];
        if (w < warp) off += cw;
        total += cw;
        overflow = overflow || cw > kSegCap;
    }
    const bool fits = !overflow && total <= kNearSmem;
    if (fits) {
        for (int k = lane; k < wc; k += 32) s_cand[off + k] = seg[k];
    } else if (!overflow && total <= v.near_cap) {
        int *list = cand2_of(v, e);
        for (int k = lane; k < wc; k += 32) list[off + k] = seg[k];
    }
    // the counter arithmetic steer_body / expand_iteration use (list length = spec_cnt - base); an overflow makes the
    // length exceed near_cap, which sends the iteration to its own Near scan
    if (tid == 0) __stcg(&IT.spec_cnt, h.base + ((overflow || total > v.near_cap) ? v.near_cap + 1 : total));
    __threadfence_block();
    __syncthreads();
    return fits ? total : -1;
}

// Small trees (capacity <= kPersistCap): the whole run of `iters` iterations of a problem in ONE launch by ONE CTA --
// the CTA scans its own problem's mirror (a few thousand vertices: L2-resident, a few microseconds), then runs the
// expansion (steer, Near filter, walks, ChooseParent, Rewire, goal bookkeeping, next sample), and loops.  No kernel
// boundary between iterations, and no lock step: with one launch per iteration of a group of problems every iteration
// lasts as long as the slowest problem of the group (measured at 64 x 5000 IRRT* 2D: 242 us per lock-step iteration
// against a mean of 63 us per problem); here a problem's time is the sum of ITS iterations.  Same device functions,
// same results (tests run both paths against each other).  Large trees keep the two-kernel pipeline: their scan needs
// the whole machine.
template <int D>
__global__ void __launch_bounds__(kExpandThreads, 4) k_iterate(View v, int iters) {
    const int e = v.env0 + blockIdx.x;
    EnvCtl *c = v.ctl + e;
    __shared__ typename GeomOf<D>::type g;            // the obstacle table stays in shared memory for the whole run
    __shared__ __align__(16) unsigned char s_buf[kNearSmem * kNearRow];
    stage_geom<D>(&g, v, e);
    __syncthreads();
    for (int it = 0; it < iters; it++) {
        const ScanHdr h = load_hdr(&c->hdr0);          // written by k_top / by this CTA's top_body of the previous iteration
        if (!h.go) break;                              // driver finished, waits for a guidance cloud, or budget exhausted
        const int presorted = scan_own_sorted<D>(v, e, c, h, s_buf);
        expand_iteration<D>(v, e, g, true, s_buf, presorted);
        __threadfence_block();
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// setup / IO kernels

// RRT* eval driver, start of a run (1 CTA / env): the goal-candidate list of the existing tree in ascending index
// order, the child lists of every vertex, the cached candidate costs, the current goal parent and path length
template <int D>
__global__ void __launch_bounds__(256) k_goal_init(View v) {
    typedef typename GeomOf<D>::type G;
    const int e = blockIdx.x;
    EnvCtl *c = v.ctl + e;
    __shared__ G g;
    __shared__ int s_wcnt[8];
    __shared__ int s_total;
    __shared__ double sm_s[8];
    __shared__ int sm_i[8];
    stage_geom<D>(&g, v, e);
    if (threadIdx.x == 0) s_total = 0;
    const int n = c->n;
    const Node *nodes = v.nodes + (size_t)e * v.stride;
    Kid *kid = v.kid + (size_t)e * v.stride;
    for (int i = threadIdx.x; i < n; i += blockDim.x) store_kid(kid + i, -1, -1, -1, -1);
    __threadfence_block();
    __syncthreads();
    // child lists: every field below has exactly one writer (v writes its own next, its successor in the push
    // order writes its prev)
    for (int i = 1 + threadIdx.x; i < n; i += blockDim.x) {
        const int p = (int)load_node(nodes + i).parent;
        const int old = atomicExch(&kid[p].head, i);
        __stcg(&kid[i].next, old);
        if (old >= 0) __stcg(&kid[old].prev, i);
    }
    __syncthreads();
    if (fam_informed(v.variant)) {
        // the candidates are the stored path_solutions; the first slot of a vertex carries its cached value
        const GoalList L = goal_list(v, c, e);
        const int ns = min(*L.count, L.cap);
        for (int k = threadIdx.x; k < ns; k += blockDim.x) {
            const int i = L.idx[k];
            const Node nd = load_node(nodes + i);
            L.d[k] = edge_len<D>(XSUB(c->goal[0], nd.x), XSUB(c->goal[1], nd.y), XSUB(c->goal[2], nd.z));
        }
        __syncthreads();
        // first occurrence of every vertex: slots are visited in ascending order by one thread per 32-slot stripe, the
        // minimum slot wins through atomicMin on a scratch encoding (gslot = -1 means "none": use unsigned compare)
        for (int k = threadIdx.x; k < ns; k += blockDim.x)
            atomicMin(reinterpret_cast<unsigned *>(&kid[L.idx[k]].gslot), (unsigned)k);
        __threadfence_block();
        __syncthreads();
        for (int k = threadIdx.x; k < ns; k += blockDim.x)
            if (__ldcg(&kid[L.idx[k]].gslot) != k) L.d[k] = XINF;     // later slot of the same vertex: never the first minimum
        __threadfence_block();
        __syncthreads();
        goal_path_len<D>(v, c, e, nodes, sm_s, sm_i);
        return;
    }
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + threadIdx.x;
        bool keep = false;
        double d = 0.0;
        if (i < n) {
            const Node nd = load_node(nodes + i);
            const double p[3] = {nd.x, nd.y, nd.z};
            const double gx = XSUB(c->goal[0], p[0]), gy = XSUB(c->goal[1], p[1]), gz = XSUB(c->goal[2], p[2]);
            const double s2 = scan_sq<D>(gx, gy, gz);
            if (s2 <= c->T_goal && (D == 3 || np_hypot(gx, gy) <= c->step_len)) {
                keep = true;
                d = seg_collides(g, p, c->goal) ? XINF : (D == 3 ? XSQRT(s2) : np_hypot(gx, gy));
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
        if (l == 0) s_wcnt[w] = __popc(bal);
        __syncthreads();
        int off = s_total;
        for (int q = 0; q < w; q++) off += s_wcnt[q];
        if (keep) {
            const int pos = off + __popc(bal & ((1u << l) - 1u));
            v.gc_idx[(size_t)e * v.cap + pos] = i;
            v.gc_d[(size_t)e * v.gc_stride + pos] = d;
            if (d < XINF) __stcg(&kid[i].gslot, pos);
        }
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int q = 0; q < 8; q++) t += s_wcnt[q]; s_total += t; }
        __syncthreads();
    }
    if (threadIdx.x == 0) c->n_goal = s_total;
    __threadfence_block();
    __syncthreads();
    goal_path_len<D>(v, c, e, nodes, sm_s, sm_i);      // fills the cost cache, best_k / best_val / last_gp / last_len
}

__global__ void k_begin(View v) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= v.E) return;
    EnvCtl *c = v.ctl + e;
    c->state = ST_PHASE1; c->saved_state = ST_PHASE1;
    c->p1_done = 0; c->left = 0; c->budget = 0; c->n_rec = 0; c->resumed = 0;
    set_idle(c); IT.skip = 0; IT.nearest = 0; IT.new_idx = -1; IT.inserted = 0; c->cand_cnt = 0; IT.near_cnt = 0;
    c->c_best = XINF; c->c_update = XINF;
    c->tree_changed = 1; c->last_len = XINF; c->last_gp = -1;
    c->err = 0;
}

__global__ void k_set_budget(View v, int iters, RunCfg cfg) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < v.E) { v.ctl[e].budget = iters; v.ctl[e].cfg = cfg; }
}

struct ProblemUpload {
    const double *start, *goal, *step_len, *search_radius, *clearance, *range, *balls, *ball_r2, *boxes, *rot_c;
    const int *n_balls, *n_boxes;
};

__global__ void k_set_problems(View v, ProblemUpload u) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= v.E) return;
    EnvCtl *c = v.ctl + e;
    Geom3 *g = v.geom + e;
    for (int i = 0; i < 3; i++) { c->start[i] = u.start[3 * e + i]; c->goal[i] = u.goal[3 * e + i]; }
    c->step_len = u.step_len[e];
    c->search_radius = u.search_radius[e];
    c->T_goal = sqrt_le_threshold(c->step_len);
    c->c_min = hypot3(XSUB(c->goal[0], c->start[0]), XSUB(c->goal[1], c->start[1]), XSUB(c->goal[2], c->start[2]));
    for (int i = 0; i < 3; i++) c->center[i] = XDIV(XADD(c->start[i], c->goal[i]), 2.0);
    for (int i = 0; i < 9; i++) c->C[i] = u.rot_c ? u.rot_c[9 * e + i] : ((i % 4) == 0 ? 1.0 : 0.0);
    g->n_balls = u.n_balls[e]; g->n_boxes = u.n_boxes[e];
    g->clearance = u.clearance[e];
    for (int i = 0; i < 6; i++) g->range[i] = u.range[6 * e + i];
    for (int k = 0; k < kMaxObs; k++) {
        for (int i = 0; i < 4; i++) g->balls[k][i] = u.balls[((size_t)e * kMaxObs + k) * 4 + i];
        g->ball_r2[k] = u.ball_r2[(size_t)e * kMaxObs + k];
        for (int i = 0; i < 6; i++) g->boxes[k][i] = u.boxes[((size_t)e * kMaxObs + k) * 6 + i];
    }
    // tree = {start}, parent[0] = 0 (rrt_base_3d.py:25-28)
    const size_t o = (size_t)e * v.stride;
    v.vx[o] = c->start[0]; v.vy[o] = c->start[1]; v.vz[o] = c->start[2];
    {
        double R = 1.0, ext = 0.0;
        for (int i = 0; i < 6; i++) R = fmax(R, fabs(g->range[i]));
        for (int i = 0; i < 3; i++) { c->qlo[i] = g->range[2 * i]; ext = fmax(ext, g->range[2 * i + 1] - g->range[2 * i]); }
        c->qscale = (v.m8 ? 255.0 : 65535.0) / fmax(ext, 1e-300);
        c->margin = v.m8 ? kMarginU8 : (v.ux ? kMarginU16 : R * 0x1p-19);
        c->fallbacks = 0;
    }
    mirror_store(v, c, o, c->start[0], c->start[1], c->start[2]);
    Node nd; nd.x = c->start[0]; nd.y = c->start[1]; nd.z = c->start[2]; nd.parent = 0;
    v.nodes[o] = nd;
    v.hints[o] = zero_hint();
    { Link l; l.elen = 0.0; l.parent = 0; l.pad = 0; v.links[o] = l; }
    c->n = 1;
    c->n_sol = 0; c->n_goal = 0; c->n_pc = 0;
    for (int k = 0; k < 8; k++) c->work[k] = 0;
    c->state = ST_DONE; c->budget = 0; set_idle(c); c->n_rec = 0; c->err = 0; c->resumed = 0;
    c->c_best = XINF; c->c_update = XINF; c->tree_changed = 1; c->last_len = XINF; c->last_gp = -1;
}

struct ProblemUpload2 {
    const double *start, *goal, *step_len, *search_radius, *clearance, *range, *circles, *rects, *rot_c;
    const int *n_circles, *n_rects;
};

// 2D problems: RRTStar2D.__init__ arguments (rrt_star_2d.py:10-30) + Utils (rrt_utils_2d.py:5-17)
__global__ void k_set_problems_2d(View v, ProblemUpload2 u) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= v.E) return;
    EnvCtl *c = v.ctl + e;
    Geom2 *g = v.geom2 + e;
    for (int i = 0; i < 2; i++) { c->start[i] = u.start[2 * e + i]; c->goal[i] = u.goal[2 * e + i]; }
    c->start[2] = c->goal[2] = 0.0;
    c->step_len = u.step_len[e];
    c->search_radius = u.search_radius[e];
    c->T_goal = hypot_band_sq(c->step_len);
    c->c_min = hypot2(XSUB(c->goal[0], c->start[0]), XSUB(c->goal[1], c->start[1]));
    for (int i = 0; i < 3; i++) c->center[i] = XDIV(XADD(c->start[i], c->goal[i]), 2.0);
    for (int i = 0; i < 9; i++) c->C[i] = u.rot_c ? u.rot_c[9 * e + i] : ((i % 4) == 0 ? 1.0 : 0.0);
    g->n_circles = u.n_circles[e]; g->n_rects = u.n_rects[e];
    g->clearance = u.clearance[e];
    for (int i = 0; i < 4; i++) g->range[i] = u.range[4 * e + i];
    for (int k = 0; k < kMaxObs; k++) {
        for (int i = 0; i < 3; i++) g->circles[k][i] = u.circles[((size_t)e * kMaxObs + k) * 3 + i];
        for (int i = 0; i < 4; i++) g->rects[k][i] = u.rects[((size_t)e * kMaxObs + k) * 4 + i];
    }
    const size_t o = (size_t)e * v.stride;
    v.vx[o] = c->start[0]; v.vy[o] = c->start[1];
    {
        double R = 1.0, ext = 0.0;
        for (int i = 0; i < 4; i++) R = fmax(R, fabs(g->range[i]));
        for (int i = 0; i < 2; i++) { c->qlo[i] = g->range[2 * i]; ext = fmax(ext, g->range[2 * i + 1] - g->range[2 * i]); }
        c->qlo[2] = 0.0;
        c->qscale = (v.m8 ? 255.0 : 65535.0) / fmax(ext, 1e-300);
        c->margin = v.m8 ? kMarginU8 : (v.ux ? kMarginU16 : R * 0x1p-19);
        c->fallbacks = 0;
    }
    mirror_store(v, c, o, c->start[0], c->start[1], 0.0);
    Node nd; nd.x = c->start[0]; nd.y = c->start[1]; nd.z = 0.0; nd.parent = 0;
    v.nodes[o] = nd;
    v.hints[o] = zero_hint();
    { Link l; l.elen = 0.0; l.parent = 0; l.pad = 0; v.links[o] = l; }
    c->n = 1;
    c->n_sol = 0; c->n_goal = 0; c->n_pc = 0;
    for (int k = 0; k < 8; k++) c->work[k] = 0;
    c->state = ST_DONE; c->budget = 0; set_idle(c); c->n_rec = 0; c->err = 0; c->resumed = 0;
    c->c_best = XINF; c->c_update = XINF; c->tree_changed = 1; c->last_len = XINF; c->last_gp = -1;
}

// stand-alone predicates ---------------------------------------------------------------------------
template <int D>
__global__ void k_collide_edges(View v, int env, const double *edges, long long m, uint8_t *out) {
    __shared__ typename GeomOf<D>::type g;
    stage_geom<D>(&g, v, env);
    __syncthreads();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x) {
        double p0[3] = {0.0, 0.0, 0.0}, p1[3] = {0.0, 0.0, 0.0};
        for (int k = 0; k < D; k++) { p0[k] = edges[2 * D * i + k]; p1[k] = edges[2 * D * i + D + k]; }
        out[i] = seg_collides(g, p0, p1) ? 1 : 0;
    }
}
template <int D>
__global__ void k_points_check(View v, int env, int kind, const double *pts, long long m, uint8_t *out) {
    __shared__ typename GeomOf<D>::type g;
    stage_geom<D>(&g, v, env);
    __syncthreads();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x) {
        double p[3] = {0.0, 0.0, 0.0};
        for (int k = 0; k < D; k++) p[k] = pts[D * i + k];
        out[i] = (kind == 0 ? point_inside_obs(g, p) : point_valid(g, p)) ? 1 : 0;
    }
}
template <int D>
__global__ void k_costs(View v, int env, const long long *idx, long long m, double *out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < m) out[i] = cost_walk<D>(tree_of(v, env), (int)idx[i]);
}
// goal parent of every env for the final search (rrt_star_3d.py:58; irrt_star_3d.py:74-76)
template <int D>
__global__ void __launch_bounds__(kExpandThreads) k_goal_parent(View v, int use_solutions, long long *gp_out, double *cost_out) {
    const int e = blockIdx.x;
    EnvCtl *c = v.ctl + e;
    __shared__ double sm_s[4];
    __shared__ int sm_i[4];
    const Node *nodes = v.nodes + (size_t)e * v.stride;
    if (use_solutions) {
        const int n_sol = min(c->n_sol, v.sol_cap);
        const int *sol = v.sol + (size_t)e * v.sol_cap;
        double bs = XINF; int bk = INT_MAX;
        for (int k = threadIdx.x; k < n_sol; k += blockDim.x) {
            const int idx = sol[k];
            const Node nd = load_node(nodes + idx);
            lexmin(bs, bk, XADD(cost_walk<D>(tree_of(v, e), idx), edge_len<D>(XSUB(c->goal[0], nd.x), XSUB(c->goal[1], nd.y), XSUB(c->goal[2], nd.z))), k);
        }
        block_lexmin(bs, bk, sm_s, sm_i);
        if (threadIdx.x == 0) { gp_out[e] = n_sol > 0 ? sol[bk] : -1; cost_out[e] = n_sol > 0 ? bs : XINF; }
    } else {
        const double len = goal_path_len<D>(v, c, e, nodes, sm_s, sm_i);
        if (threadIdx.x == 0) { gp_out[e] = c->last_gp; cost_out[e] = len; }
    }
}

// ------------------------------------------------------------------------------------------------
// host side: batch object

// launches kernel template KERNEL<D, ...> for the batch's dimension
#define LAUNCH_D(dim, KERNEL, grid, block, smem, stream, ...)                          \
    do {                                                                               \
        if ((dim) == 3) KERNEL<3><<<grid, block, smem, stream>>>(__VA_ARGS__);         \
        else KERNEL<2><<<grid, block, smem, stream>>>(__VA_ARGS__);                    \
    } while (0)
#define LAUNCH_DB(dim, KERNEL, FORCE, grid, block, smem, stream, ...)                  \
    do {                                                                               \
        if ((dim) == 3) KERNEL<3, FORCE><<<grid, block, smem, stream>>>(__VA_ARGS__);  \
        else KERNEL<2, FORCE><<<grid, block, smem, stream>>>(__VA_ARGS__);             \
    } while (0)

// launch of a kernel(View); pdl: allow it to become resident while its predecessor in the stream drains
static void launch_view(void (*kernel)(View), dim3 grid, dim3 block, cudaStream_t s, const View &v, bool pdl, size_t smem = 0) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, v);
}

// the two scans of one iteration in the batch's scan layout (u16 / f32 mirror or the f64 arrays)
template <bool kForce>
static void launch_scan(const View &v, int which, int count, cudaStream_t s, bool pdl = false) {
    const dim3 grid(v.chunks, count);
    void (*k)(View);
    if (v.m8 && which == 0 && v.sad) {
        k = v.dim == 3 ? k_nearest_s<3, kForce> : k_nearest_s<2, kForce>;
    } else if (v.m8 && which == 0) {
        k = v.dim == 3 ? k_nearest_b<3, kForce> : k_nearest_b<2, kForce>;
    } else if (v.ux) {
        if (which == 0 && v.tma == 1) k = v.dim == 3 ? k_nearest_t<3, kForce, 3, 4> : k_nearest_t<2, kForce, 3, 4>;
        else if (which == 0 && v.tma == 2) k = v.dim == 3 ? k_nearest_t<3, kForce, 2, 6> : k_nearest_t<2, kForce, 2, 6>;
        else if (which == 0 && v.tma == 3) k = v.dim == 3 ? k_nearest_t<3, kForce, 2, 8> : k_nearest_t<2, kForce, 2, 8>;
        else if (which == 0 && v.tma == 4) k = v.dim == 3 ? k_nearest_m<3, true, kForce, true> : k_nearest_m<2, true, kForce, true>;
        else if (which == 0) k = v.dim == 3 ? k_nearest_m<3, true, kForce> : k_nearest_m<2, true, kForce>;
        else k = v.dim == 3 ? k_near_m<3, true, kForce> : k_near_m<2, true, kForce>;
    } else if (v.fx) {
        if (which == 0) k = v.dim == 3 ? k_nearest_m<3, false, kForce> : k_nearest_m<2, false, kForce>;
        else k = v.dim == 3 ? k_near_m<3, false, kForce> : k_near_m<2, false, kForce>;
    } else {
        if (which == 0) k = v.dim == 3 ? k_nearest<3, kForce> : k_nearest<2, kForce>;
        else k = v.dim == 3 ? k_near<3, kForce> : k_near<2, kForce>;
    }
    // (Capping the scan's residency with a dummy shared-memory reservation so that k_expand CTAs always find room
    // was measured and is worse: 86-94 us per step at 6 / 5 / 4 scan CTAs per SM against 72 us uncapped.)
    launch_view(k, grid, 256, s, v, pdl);
}

constexpr int kPersistCap = 32768;       // capacity up to which a problem runs as one persistent CTA (k_iterate)
constexpr int kMaxGroups = 32;
constexpr int kGraphItersDefault = 16;   // steady-state iterations per CUDA graph replay (see ensure_graph)
struct nirrt_batch {
    View v;
    bool pdl;        // iteration kernels use programmatic dependent launch (NIRRT_PDL=0 disables)
    bool use_graph;  // steady-state iterations are replayed from a CUDA graph (NIRRT_GRAPH=0 disables)
    // graph cache: one executable per distinct (View, pipelined) -- i.e. per (variant, mode) of this batch; the
    // run parameters live in EnvCtl::cfg, so begin() with a known variant/mode re-uses its graph
    struct GraphEntry { View view; bool pipelined; cudaGraphExec_t exec; int64_t launches; int64_t last_use; };
    std::vector<GraphEntry> graphs;
    int64_t graph_builds, graph_replays, graph_fallbacks, graph_clock;
    int graph_iters; // iterations per graph replay
    RunCfg cfg;      // run parameters, handed to the device by k_set_budget at the start of every run
    bool persistent;            // small trees: one k_iterate launch per run instead of two kernels per iteration
    struct CloudWs *cloud_ws;   // device workspace of the guidance-cloud generator (allocated on first use)
    int cloud_count;            // problems in the last nirrt_batch_sample_clouds_sync call
    cudaStream_t cs; // capture origin
    int device;
    size_t stride_bytes;
    std::vector<void *> allocs;
    int64_t launches;
    bool goal_lists;   // gc_idx/gc_d allocated
    // group pipeline (nirrt_batch_run): the problems are split in G groups (16 for >= 512 problems,
    // NIRRT_GROUPS overrides) that run the same kernel sequence on G internal streams, so the
    // latency-bound k_expand of one group (dependent pointer chasing) overlaps the HBM-bound scans of
    // the others.  Measured at 512 x 100k vertices (late builds of round 1): 0.108 ms/step with one group,
    // 0.106 (2), 0.091 (4), 0.073 (8), 0.072 (16 groups, CUDA-graph replay).
    int groups;
    cudaStream_t gs[kMaxGroups];
    cudaEvent_t ev_fork, ev_join[kMaxGroups];
    // pipelined RRT* driver: second stream per group for k_expand, events per IterScratch copy
    bool pipeline;   // NIRRT_PIPELINE=0 disables
    cudaStream_t gs2[kMaxGroups];
    cudaEvent_t ev_a[kMaxGroups][2], ev_b[kMaxGroups][2];
    // pinned scratch for small synchronous reads
    EnvCtl *h_ctl;
    // tree transfers (load_trees / read_trees): two persistent device staging buffers on two internal
    // streams, so the layout kernel of one chunk of problems overlaps the PCIe copy of the next
    void *stage[2];
    int *stage_n;
    int stage_envs;          // problems per staging buffer
    cudaStream_t xs[2];
    cudaEvent_t xe[2], xfork;
};

static int ensure_staging(nirrt_batch *b) {
    if (b->stage[0]) return NIRRT_OK;
    const View &v = b->v;
    const size_t per_env = (size_t)v.cap * (v.dim * sizeof(double) + sizeof(long long));
    size_t envs = ((size_t)128 << 20) / (per_env ? per_env : 1);   // ~128 MB per buffer
    if (envs < 1) envs = 1;
    if (envs > (size_t)v.E) envs = v.E;
    b->stage_envs = (int)envs;
    for (int i = 0; i < 2; i++) {
        CUDA_TRY(nirrt_dev_malloc(&b->stage[i], envs * per_env));
        b->allocs.push_back(b->stage[i]);
        CUDA_TRY(cudaStreamCreateWithFlags(&b->xs[i], cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&b->xe[i], cudaEventDisableTiming));
    }
    CUDA_TRY(cudaEventCreateWithFlags(&b->xfork, cudaEventDisableTiming));
    CUDA_TRY(nirrt_dev_malloc((void **)&b->stage_n, sizeof(int) * v.E));
    b->allocs.push_back(b->stage_n);
    return NIRRT_OK;
}

static int dalloc(nirrt_batch *b, void **p, size_t bytes) {
    cudaError_t e = nirrt_dev_malloc(p, bytes ? bytes : 16);
    if (e != cudaSuccess) return fail(NIRRT_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    b->allocs.push_back(*p);
    return NIRRT_OK;
}
#define DALLOC(ptr, type, count)                                                          \
    do { void *_p = nullptr; int _r = dalloc(b, &_p, sizeof(type) * (size_t)(count));     \
         if (_r) { nirrt_batch_destroy(b); return _r; } ptr = (type *)_p; } while (0)

static int pick_chunks(int E) {
    const char *env = getenv("NIRRT_CHUNKS");
    if (env && atoi(env) > 0) return atoi(env) > 64 ? 64 : atoi(env);
    // aim for ~4 resident 256-thread CTAs of one group's scan on each of the 148 SMs (the other groups'
    // kernels fill the rest; measured on 512 x 100k: 10 chunks at 64 problems per group beat 5 and 20)
    int c = (148 * 4 + E - 1) / E;
    if (E >= 32 && c > 10) c = 10;        // 16 groups of 32 problems: 10 chunks (72.3 us / step) beat 19 (76.9)
    if (c < 1) c = 1;
    if (c > 64) c = 64;
    return c;
}

extern "C" int nirrt_batch_destroy(nirrt_batch *b) {
    if (!b) return NIRRT_OK;
    cudaSetDevice(b->device);
    for (void *p : b->allocs) nirrt_dev_free(p);
    if (b->h_ctl) cudaFreeHost(b->h_ctl);
    free(b->cloud_ws);
    for (int g = 0; g < kMaxGroups; g++) {
        if (b->gs[g]) cudaStreamDestroy(b->gs[g]);
        if (b->ev_join[g]) cudaEventDestroy(b->ev_join[g]);
        if (b->gs2[g]) cudaStreamDestroy(b->gs2[g]);
        for (int q = 0; q < 2; q++) {
            if (b->ev_a[g][q]) cudaEventDestroy(b->ev_a[g][q]);
            if (b->ev_b[g][q]) cudaEventDestroy(b->ev_b[g][q]);
        }
    }
    if (b->ev_fork) cudaEventDestroy(b->ev_fork);
    for (auto &g : b->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    if (b->cs) cudaStreamDestroy(b->cs);
    for (int i = 0; i < 2; i++) {
        if (b->xs[i]) cudaStreamDestroy(b->xs[i]);
        if (b->xe[i]) cudaEventDestroy(b->xe[i]);
    }
    if (b->xfork) cudaEventDestroy(b->xfork);
    delete b;
    return NIRRT_OK;
}

extern "C" int nirrt_batch_create(const nirrt_batch_desc *d, nirrt_batch **out) {
    if (!d || !out) return fail(NIRRT_ERR_INVALID, "null argument");
    if (d->dim != 3 && d->dim != 2) return fail(NIRRT_ERR_INVALID, "nirrt_batch_create: dim must be 2 or 3");
    if (d->n_envs < 1 || d->capacity < 2) return fail(NIRRT_ERR_INVALID, "n_envs >= 1 and capacity >= 2 required");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(NIRRT_ERR_NO_DEVICE, "no CUDA device visible");
    if (d->device < 0 || d->device >= ndev) return fail(NIRRT_ERR_INVALID, "bad device ordinal");
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, d->device));
    if (prop.major != 10) return fail(NIRRT_ERR_NO_DEVICE, "libnirrt_b200 is built for sm_100a only");
    CUDA_TRY(cudaSetDevice(d->device));
    nirrt_batch *b = new nirrt_batch();
    memset(&b->v, 0, sizeof(View));
    b->device = d->device; b->launches = 0; b->goal_lists = false; b->h_ctl = nullptr;
    b->cloud_ws = nullptr; b->cloud_count = 0; b->persistent = false;
    b->groups = 1; b->ev_fork = nullptr;
    for (int g = 0; g < kMaxGroups; g++) {
        b->gs[g] = nullptr; b->ev_join[g] = nullptr; b->gs2[g] = nullptr;
        for (int q = 0; q < 2; q++) { b->ev_a[g][q] = nullptr; b->ev_b[g][q] = nullptr; }
    }
    // off by default: measured 77-79 us per step against 72 us unpipelined at 512 x 100k (the step is bound by the
    // SM time of the scan, not by the dependency chain); NIRRT_PIPELINE=1 enables it
    { const char *pl = getenv("NIRRT_PIPELINE"); b->pipeline = pl && atoi(pl) == 1; }
    for (int i = 0; i < 2; i++) { b->stage[i] = nullptr; b->xs[i] = nullptr; b->xe[i] = nullptr; }
    b->stage_n = nullptr; b->stage_envs = 0; b->xfork = nullptr;
    View &v = b->v;
    v.E = d->n_envs; v.cap = d->capacity; v.dim = d->dim;
    v.stride = (d->capacity + 63) & ~63;
    {
        const char *pdl = getenv("NIRRT_PDL");
        b->pdl = !(pdl && atoi(pdl) == 0);
        const char *gr = getenv("NIRRT_GRAPH");
        b->graph_iters = gr ? atoi(gr) : kGraphItersDefault;     // NIRRT_GRAPH=<iterations per graph>, 0 disables
        if (b->graph_iters > 256) b->graph_iters = 256;
        b->graph_iters &= ~1;      // even: a pipelined block ends on IterScratch copy 0
        b->use_graph = b->graph_iters > 0;
        b->graph_builds = b->graph_replays = b->graph_fallbacks = b->graph_clock = 0; b->cs = nullptr;
        const char *g = getenv("NIRRT_GROUPS");
        b->groups = g ? atoi(g) : (d->n_envs >= 512 ? 16 : (d->n_envs >= 256 ? 8 : (d->n_envs >= 128 ? 4 : (d->n_envs >= 64 ? 2 : 1))));
        if (b->groups < 1) b->groups = 1;
        if (b->groups > kMaxGroups) b->groups = kMaxGroups;
        if (b->groups > d->n_envs) b->groups = d->n_envs;
    }
    v.chunks = pick_chunks((v.E + b->groups - 1) / b->groups);
    {
        // NIRRT_PERSIST: largest capacity (vertices per problem) that runs as one persistent CTA per problem (0 disables)
        const char *pe = getenv("NIRRT_PERSIST");
        const int persist_cap = pe ? atoi(pe) : kPersistCap;
        const char *mode = getenv("NIRRT_SCAN");
        b->persistent = persist_cap > 0 && d->capacity <= persist_cap && (!mode || strcmp(mode, "u16") == 0 || strcmp(mode, "u16ldg") == 0);
    }
    v.near_cap = d->near_capacity > 0 ? ((d->near_capacity + 63) & ~63) : kNearSmem;
    if (v.near_cap > kNearBig) { delete b; return fail(NIRRT_ERR_INVALID, "near_capacity: at most 8192"); }
    v.rec_cap = d->record_capacity > 0 ? d->record_capacity : d->capacity + 8;
    v.sol_cap = v.rec_cap > v.cap + 8 ? v.rec_cap : v.cap + 8;   // at most one append per iteration
    v.pc_cap = 4096; v.path_cap = 4096;
    memset(&b->cfg, 0, sizeof(RunCfg));
    b->cfg.pc_rate = 0.5; b->cfg.pc_ratio = 0.9; b->cfg.stop_below = (double)INFINITY;
    const size_t EV = (size_t)v.E * v.stride;
    DALLOC(v.vx, double, EV); DALLOC(v.vy, double, EV);
    if (v.dim == 3) DALLOC(v.vz, double, EV);
    {
        // NIRRT_SCAN: "u16" (default) fixed-point mirror, "f32" float mirror, "f64" scan the f64 arrays
        const char *mode = getenv("NIRRT_SCAN");
        if (mode && strcmp(mode, "f64") == 0) {
        } else if (mode && strcmp(mode, "f32") == 0) {
            DALLOC(v.fx, float, EV); DALLOC(v.fy, float, EV);
            if (v.dim == 3) DALLOC(v.fz, float, EV);
        } else if (mode && (strcmp(mode, "u8") == 0 || strcmp(mode, "s8") == 0)) {
            DALLOC(v.m8, unsigned, EV);
            v.sad = strcmp(mode, "s8") == 0;
        } else {
            DALLOC(v.ux, unsigned short, EV); DALLOC(v.uy, unsigned short, EV);
            if (v.dim == 3) DALLOC(v.uz, unsigned short, EV);
            v.tma = 0;      // 0: LDG kernel (measured best); NIRRT_TMA = 1..3: TMA-staged variants, 4: software-pipelined LDG
            if (const char *tv = getenv("NIRRT_TMA")) v.tma = atoi(tv);
        }
    }
    DALLOC(v.nodes, Node, EV);
    DALLOC(v.hints, Hint, EV);
    DALLOC(v.links, Link, EV);
    if (v.dim == 3) { DALLOC(v.geom, Geom3, v.E); }
    else {
        DALLOC(v.geom2, Geom2, v.E); DALLOC(v.mt_py, MtState, v.E);
        CUDA_TRY(cudaMemset(v.mt_py, 0, sizeof(MtState) * v.E));
    }
    DALLOC(v.mt, MtState, v.E); DALLOC(v.ctl, EnvCtl, v.E);
    DALLOC(v.part_s, double, (size_t)v.E * v.chunks); DALLOC(v.part_i, int, (size_t)v.E * v.chunks);
    DALLOC(v.cand, int, (size_t)v.E * v.near_cap); DALLOC(v.near_out, int, (size_t)v.E * v.near_cap);
    DALLOC(v.cand2, int, (size_t)2 * v.E * v.near_cap);     // two halves: IterScratch copies 0 / 1
    if (v.near_cap > kNearSmem) DALLOC(v.big, unsigned char, (size_t)v.E * v.near_cap * kNearRowAlloc);
    DALLOC(v.sol, int, (size_t)v.E * v.sol_cap);
    DALLOC(v.records, double, (size_t)v.E * v.rec_cap);
    DALLOC(v.pathseg, double, (size_t)v.E * v.path_cap);
    { double *t; DALLOC(t, double, (size_t)v.cap + 2); v.near_table = t; }
    CUDA_TRY(cudaMemset(v.ctl, 0, sizeof(EnvCtl) * v.E));
    CUDA_TRY(cudaMemset(v.mt, 0, sizeof(MtState) * v.E));
    CUDA_TRY(cudaMallocHost((void **)&b->h_ctl, sizeof(EnvCtl) * v.E));
    if (b->groups >= 2) {
        for (int g = 0; g < b->groups; g++) {
            CUDA_TRY(cudaStreamCreateWithFlags(&b->gs[g], cudaStreamNonBlocking));
            CUDA_TRY(cudaEventCreateWithFlags(&b->ev_join[g], cudaEventDisableTiming));
            CUDA_TRY(cudaStreamCreateWithFlags(&b->gs2[g], cudaStreamNonBlocking));
            for (int q = 0; q < 2; q++) {
                CUDA_TRY(cudaEventCreateWithFlags(&b->ev_a[g][q], cudaEventDisableTiming));
                CUDA_TRY(cudaEventCreateWithFlags(&b->ev_b[g][q], cudaEventDisableTiming));
            }
        }
        CUDA_TRY(cudaEventCreateWithFlags(&b->ev_fork, cudaEventDisableTiming));
        CUDA_TRY(cudaStreamCreateWithFlags(&b->cs, cudaStreamNonBlocking));
    }
    *out = b;
    return NIRRT_OK;
}

static int ensure_goal_lists(nirrt_batch *b) {
    if (b->goal_lists) return NIRRT_OK;
    View &v = b->v;
    void *p = nullptr;
    int r = dalloc(b, &p, sizeof(int) * (size_t)v.E * v.cap); if (r) return r; v.gc_idx = (int *)p;
    v.gc_stride = v.sol_cap > v.cap ? v.sol_cap : v.cap;
    r = dalloc(b, &p, sizeof(double) * (size_t)v.E * v.gc_stride); if (r) return r; v.gc_d = (double *)p;
    r = dalloc(b, &p, sizeof(double) * (size_t)v.E * v.gc_stride); if (r) return r; v.gc_cost = (double *)p;
    r = dalloc(b, &p, sizeof(Kid) * (size_t)v.E * v.stride); if (r) return r; v.kid = (Kid *)p;
    b->goal_lists = true;
    return NIRRT_OK;
}

// copies a host array to a temporary device buffer on `s` (freed after the stream is drained by the caller)
struct TempBufs {
    std::vector<void *> ptrs;
    ~TempBufs() { for (void *p : ptrs) nirrt_dev_free(p); }
    template <typename T> int up(const T *host, size_t count, cudaStream_t s, T **dev) {
        void *p = nullptr;
        cudaError_t e = nirrt_dev_malloc(&p, sizeof(T) * (count ? count : 1));
        if (e != cudaSuccess) return fail(NIRRT_ERR_CUDA, std::string("cudaMalloc(temp): ") + cudaGetErrorString(e));
        ptrs.push_back(p);
        if (count) {
            e = cudaMemcpyAsync(p, host, sizeof(T) * count, cudaMemcpyHostToDevice, s);
            if (e != cudaSuccess) return fail(NIRRT_ERR_CUDA, std::string("cudaMemcpyAsync(H2D): ") + cudaGetErrorString(e));
        }
        *dev = (T *)p;
        return NIRRT_OK;
    }
    template <typename T> int make(size_t count, T **dev) {
        void *p = nullptr;
        cudaError_t e = nirrt_dev_malloc(&p, sizeof(T) * (count ? count : 1));
        if (e != cudaSuccess) return fail(NIRRT_ERR_CUDA, std::string("cudaMalloc(temp): ") + cudaGetErrorString(e));
        ptrs.push_back(p);
        *dev = (T *)p;
        return NIRRT_OK;
    }
};
#define TRY(expr) do { int _r = (expr); if (_r) return _r; } while (0)
#define CHECK_LAUNCH() CUDA_TRY(cudaGetLastError())

extern "C" int nirrt_batch_set_problems(nirrt_batch *b, const double *start, const double *goal,
                                        const double *step_len, const double *search_radius, const double *clearance,
                                        const double *range, const int *n_balls, const double *balls,
                                        const double *ball_r2, const int *n_boxes, const double *boxes,
                                        const double *near_table, const double *rot_c, void *stream) {
    if (!b || !start || !goal || !step_len || !search_radius || !clearance || !range || !n_balls || !balls || !ball_r2 ||
        !n_boxes || !boxes || !near_table)
        return fail(NIRRT_ERR_INVALID, "nirrt_batch_set_problems: null argument");
    View &v = b->v;
    if (v.dim != 3) return fail(NIRRT_ERR_INVALID, "nirrt_batch_set_problems: batch is 2D, use nirrt_batch_set_problems_2d");
    for (int e = 0; e < v.E; e++)
        if (n_balls[e] < 0 || n_balls[e] > kMaxObs || n_boxes[e] < 0 || n_boxes[e] > kMaxObs)
            return fail(NIRRT_ERR_INVALID, "more than NIRRT_MAX_OBSTACLES obstacles of one type");
    CUDA_TRY(cudaSetDevice(b->device));
    cudaStream_t s = (cudaStream_t)stream;
    TempBufs t;
    ProblemUpload u;
    double *dp; int *ip;
    const size_t E = v.E;
    TRY(t.up(start, 3 * E, s, &dp)); u.start = dp;
    TRY(t.up(goal, 3 * E, s, &dp)); u.goal = dp;
    TRY(t.up(step_len, E, s, &dp)); u.step_len = dp;
    TRY(t.up(search_radius, E, s, &dp)); u.search_radius = dp;
    TRY(t.up(clearance, E, s, &dp)); u.clearance = dp;
    TRY(t.up(range, 6 * E, s, &dp)); u.range = dp;
    TRY(t.up(balls, 4 * E * kMaxObs, s, &dp)); u.balls = dp;
    TRY(t.up(ball_r2, E * kMaxObs, s, &dp)); u.ball_r2 = dp;
    TRY(t.up(boxes, 6 * E * kMaxObs, s, &dp)); u.boxes = dp;
    u.rot_c = nullptr;
    if (rot_c) { TRY(t.up(rot_c, 9 * E, s, &dp)); u.rot_c = dp; }
    TRY(t.up(n_balls, E, s, &ip)); u.n_balls = ip;
    TRY(t.up(n_boxes, E, s, &ip)); u.n_boxes = ip;
    CUDA_TRY(cudaMemcpyAsync((void *)v.near_table, near_table, sizeof(double) * ((size_t)v.cap + 2), cudaMemcpyHostToDevice, s));
    k_set_problems<<<(v.E + 127) / 128, 128, 0, s>>>(v, u);
    CHECK_LAUNCH();
    CUDA_TRY(cudaStreamSynchronize(s));   // temporaries are freed on return
    return NIRRT_OK;
}

extern "C" int nirrt_batch_set_problems_2d(nirrt_batch *b, const double *start, const double *goal,
                                           const double *step_len, const double *search_radius, const double *clearance,
                                           const double *range, const int *n_circles, const double *circles,
                                           const int *n_rects, const double *rects, const double *near_table,
                                           const double *rot_c, void *stream) {
    if (!b || !start || !goal || !step_len || !search_radius || !clearance || !range || !n_circles || !circles ||
        !n_rects || !rects || !near_table)
        return fail(NIRRT_ERR_INVALID, "nirrt_batch_set_problems_2d: null argument");
    View &v = b->v;
    if (v.dim != 2) return fail(NIRRT_ERR_INVALID, "nirrt_batch_set_problems_2d: batch is 3D");
    for (int e = 0; e < v.E; e++)
        if (n_circles[e] < 0 || n_circles[e] > kMaxObs || n_rects[e] < 0 || n_rects[e] > kMaxObs)
            return fail(NIRRT_ERR_INVALID, "more than NIRRT_MAX_OBSTACLES obstacles of one type");
    CUDA_TRY(cudaSetDevice(b->device));
    cudaStream_t s = (cudaStream_t)stream;
    TempBufs t;
    ProblemUpload2 u;
    double *dp; int *ip;
    const size_t E = v.E;
    TRY(t.up(start, 2 * E, s, &dp)); u.start = dp;
    TRY(t.up(goal, 2 * E, s, &dp)); u.goal = dp;
    TRY(t.up(step_len, E, s, &dp)); u.step_len = dp;
    TRY(t.up(search_radius, E, s, &dp)); u.search_radius = dp;
    TRY(t.up(clearance, E, s, &dp)); u.clearance = dp;
    TRY(t.up(range, 4 * E, s, &dp)); u.range = dp;
    TRY(t.up(circles, 3 * E * kMaxObs, s, &dp)); u.circles = dp;
    TRY(t.up(rects, 4 * E * kMaxObs, s, &dp)); u.rects = dp;
    u.rot_c = nullptr;
    if (rot_c) { TRY(t.up(rot_c, 9 * E, s, &dp)); u.rot_c = dp; }
    TRY(t.up(n_circles, E, s, &ip)); u.n_circles = ip;
    TRY(t.up(n_rects, E, s, &ip)); u.n_rects = ip;
    CUDA_TRY(cudaMemcpyAsync((void *)v.near_table, near_table, sizeof(double) * ((size_t)v.cap + 2), cudaMemcpyHostToDevice, s));
    k_set_problems_2d<<<(v.E + 127) / 128, 128, 0, s>>>(v, u);
    CHECK_LAUNCH();
    CUDA_TRY(cudaStreamSynchronize(s));
    return NIRRT_OK;
}

static int upload_mt(nirrt_batch *b, MtState *dst, const uint32_t *key, const int *pos, cudaStream_t s) {
    View &v = b->v;
    std::vector<MtState> h(v.E);
    for (int e = 0; e < v.E; e++) {
        memcpy(h[e].key[0], key + (size_t)e * 624, 624 * sizeof(uint32_t));
        memset(h[e].key[1], 0, 624 * sizeof(uint32_t));
        if (pos[e] < 0 || pos[e] > 624) return fail(NIRRT_ERR_INVALID, "rng pos out of range");
        h[e].pos = pos[e]; h[e].cur = 0; h[e].has_next = 0; h[e].pad = 0;
    }
    CUDA_TRY(cudaMemcpyAsync(dst, h.data(), sizeof(MtState) * v.E, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return NIRRT_OK;
}
static int download_mt(nirrt_batch *b, const MtState *src, uint32_t *key, int *pos, cudaStream_t s) {
    View &v = b->v;
    std::vector<MtState> h(v.E);
    CUDA_TRY(cudaMemcpyAsync(h.data(), src, sizeof(MtState) * v.E, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    for (int e = 0; e < v.E; e++) {
        memcpy(key + (size_t)e * 624, h[e].key[h[e].cur], 624 * sizeof(uint32_t));
        pos[e] = h[e].pos;
    }
    return NIRRT_OK;
}
// random.getstate()[1] of each problem (CPython MT19937: 624 key words + position), 2D batches only
extern "C" int nirrt_batch_set_py_rng(nirrt_batch *b, const uint32_t *key, const int *pos, void *stream) {
    if (!b || !key || !pos) return fail(NIRRT_ERR_INVALID, "nirrt_batch_set_py_rng: null argument");
    if (b->v.dim != 2) return fail(NIRRT_ERR_INVALID, "nirrt_batch_set_py_rng: only 2D planners consume the CPython stream");
    CUDA_TRY(cudaSetDevice(b->device));
    return upload_mt(b, b->v.mt_py, key, pos, (cudaStream_t)stream);
}
extern "C" int nirrt_batch_get_py_rng_sync(nirrt_batch *b, uint32_t *key, int *pos, void *stream) {
    if (!b || !key || !pos) return fail(NIRRT_ERR_INVALID, "nirrt_batch_get_py_rng_sync: null argument");
    if (b->v.dim != 2) return fail(NIRRT_ERR_INVALID, "nirrt_batch_get_py_rng_sync: only 2D planners consume the CPython stream");
    CUDA_TRY(cudaSetDevice(b->device));
    return download_mt(b, b->v.mt_py, key, pos, (cudaStream_t)stream);
}

extern "C" int nirrt_batch_set_rng(nirrt_batch *b, const uint32_t *key, const int *pos, void *stream) {
    if (!b || !key || !pos) return fail(NIRRT_ERR_INVALID, "nirrt_batch_set_rng: null argument");
    View &v = b->v;
    CUDA_TRY(cudaSetDevice(b->device));
    cudaStream_t s = (cudaStream_t)stream;
    std::vector<MtState> h(v.E);
    for (int e = 0; e < v.E; e++) {
        memcpy(h[e].key[0], key + (size_t)e * 624, 624 * sizeof(uint32_t));
        memset(h[e].key[1], 0, 624 * sizeof(uint32_t));
        if (pos[e] < 0 || pos[e] > 624) return fail(NIRRT_ERR_INVALID, "rng pos out of range");
        h[e].pos = pos[e]; h[e].cur = 0; h[e].has_next = 0; h[e].pad = 0;
    }
    CUDA_TRY(cudaMemcpyAsync(v.mt, h.data(), sizeof(MtState) * v.E, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return NIRRT_OK;
}

extern "C" int nirrt_batch_get_rng_sync(nirrt_batch *b, uint32_t *key, int *pos, void *stream) {
    if (!b || !key || !pos) return fail(NIRRT_ERR_INVALID, "nirrt_batch_get_rng_sync: null argument");
    View &v = b->v;
    CUDA_TRY(cudaSetDevice(b->device));
    cudaStream_t s = (cudaStream_t)stream;
    std::vector<MtState> h(v.E);
    CUDA_TRY(cudaMemcpyAsync(h.data(), v.mt, sizeof(MtState) * v.E, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    for (int e = 0; e < v.E; e++) {
        memcpy(key + (size_t)e * 624, h[e].key[h[e].cur], 624 * sizeof(uint32_t));
        pos[e] = h[e].pos;
    }
    return NIRRT_OK;
}

extern "C" int nirrt_batch_set_guidance(nirrt_batch *b, double pc_sample_rate, double pc_update_cost_ratio) {
    if (!b) return fail(NIRRT_ERR_INVALID, "null batch");
    b->cfg.pc_rate = pc_sample_rate; b->cfg.pc_ratio = pc_update_cost_ratio;
    return NIRRT_OK;
}

__global__ void k_set_cloud_meta(View v, int env, int n) {
    EnvCtl *c = v.ctl + env;
    c->n_pc = n;
    if (c->state == ST_WAIT_CLOUD) c->state = c->saved_state;
}

extern "C" int nirrt_batch_set_cloud(nirrt_batch *b, int env, const double *points, int n, void *stream) {
    if (!b || env < 0 || env >= b->v.E || n < 0 || (n > 0 && !points)) return fail(NIRRT_ERR_INVALID, "nirrt_batch_set_cloud: bad argument");
    View &v = b->v;
    if (n > v.pc_cap) return fail(NIRRT_ERR_CAPACITY, "guidance cloud larger than 4096 points");
    CUDA_TRY(cudaSetDevice(b->device));
    cudaStream_t s = (cudaStream_t)stream;
    if (!v.pc) {
        void *p = nullptr;
        TRY(dalloc(b, &p, sizeof(double) * 3 * (size_t)v.E * v.pc_cap));
        v.pc = (double *)p;
    }
    if (n) {
        if (v.dim == 3) {
            CUDA_TRY(cudaMemcpyAsync(v.pc + (size_t)env * v.pc_cap * 3, points, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, s));
        } else {   // 2D clouds [n][2] are stored as [n][3] with z = 0
            std::vector<double> tmp((size_t)n * 3, 0.0);
            for (int i = 0; i < n; i++) { tmp[3 * (size_t)i] = points[2 * (size_t)i]; tmp[3 * (size_t)i + 1] = points[2 * (size_t)i + 1]; }
            CUDA_TRY(cudaMemcpyAsync(v.pc + (size_t)env * v.pc_cap * 3, tmp.data(), sizeof(double) * 3 * n, cudaMemcpyHostToDevice, s));
            CUDA_TRY(cudaStreamSynchronize(s));
        }
    }
    k_set_cloud_meta<<<1, 1, 0, s>>>(v, env, n);
    CHECK_LAUNCH();
    CUDA_TRY(cudaStreamSynchronize(s));
    return NIRRT_OK;
}

// Tree transfers move whole groups of problems per copy (staging <= 2 GB): two large PCIe
// transfers and one layout kernel per group instead of per-problem copies.
__global__ void k_scatter_trees(View v, int env_begin, const int *n, const double *verts, const long long *parents) {
    const int k = blockIdx.y;
    const int nk = n[k];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nk; i += gridDim.x * blockDim.x) {
        const int D = v.dim, env = env_begin + k;
        const double *vs = verts + ((size_t)k * v.cap + i) * D;
        const size_t o = (size_t)env * v.stride + i;
        Node nd; nd.x = vs[0]; nd.y = vs[1]; nd.z = D == 3 ? vs[2] : 0.0; nd.parent = parents[(size_t)k * v.cap + i];
        v.vx[o] = nd.x; v.vy[o] = nd.y;
        if (D == 3) v.vz[o] = nd.z;
        if (has_mirror(v)) {
            mirror_store(v, v.ctl + env, o, nd.x, nd.y, nd.z);
            const double *rg = D == 3 ? v.geom[env].range : v.geom2[env].range;
            bool out = nd.x < rg[0] || nd.x > rg[1] || nd.y < rg[2] || nd.y > rg[3];
            if (D == 3) out = out || nd.z < rg[4] || nd.z > rg[5];
            if (out) atomicOr(&v.ctl[env].err, ERR_OUT_OF_RANGE);
        }
        v.nodes[o] = nd;
        if (i == 0) { v.ctl[env].n = nk; v.ctl[env].tree_changed = 1; }
    }
}
// walk records of freshly loaded trees: edge length to the parent and the true ancestors 1..8 hops up
template <int D>
__global__ void k_build_links(View v, int env_begin, const int *n) {
    const int k = blockIdx.y, env = env_begin + k;
    const int nk = n[k];
    const Node *nodes = v.nodes + (size_t)env * v.stride;
    Hint *hints = v.hints + (size_t)env * v.stride;
    Link *links = v.links + (size_t)env * v.stride;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nk; i += gridDim.x * blockDim.x) {
        Hint hh;
        int cur = i;
        for (int q = 0; q < kHintHops; q++) {
            const long long par = nodes[cur].parent;
            cur = (par >= 0 && par < nk) ? (int)par : 0;     // malformed parents only make a useless hint
            hh.a[q] = cur;
        }
        hints[i] = hh;
        const Node me = nodes[i], pa = nodes[hh.a[0]];
        Link l;
        l.elen = i == 0 ? 0.0 : edge_len<D>(XSUB(me.x, pa.x), XSUB(me.y, pa.y), XSUB(me.z, pa.z));
        l.parent = hh.a[0]; l.pad = 0;
        links[i] = l;
    }
}
__global__ void k_gather_trees(View v, int env_begin, double *verts, long long *parents) {
    const int k = blockIdx.y, env = env_begin + k, D = v.dim;
    const int n = v.ctl[env].n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < v.cap; i += gridDim.x * blockDim.x) {
        Node nd; nd.x = nd.y = nd.z = 0.0; nd.parent = 0;
        if (i < n) nd = v.nodes[(size_t)env * v.stride + i];
        double *vs = verts + ((size_t)k * v.cap + i) * D;
        vs[0] = nd.x; vs[1] = nd.y;
        if (D == 3) vs[2] = nd.z;
        parents[(size_t)k * v.cap + i] = nd.parent;
    }
}

extern "C" int nirrt_batch_load_trees(nirrt_batch *b, int env_begin, int count, const int *n,
                                      const double *vertices, const int64_t *parents, void *stream) {
    if (!b || !n || !vertices || !parents) return fail(NIRRT_ERR_INVALID, "nirrt_batch_load_trees: null argument");
    View &v = b->v;
    if (env_begin < 0 || count < 0 || env_begin + count > v.E) return fail(NIRRT_ERR_INVALID, "env range out of bounds");
    CUDA_TRY(cudaSetDevice(b->device));
    cudaStream_t s = (cudaStream_t)stream;
    for (int k = 0; k < count; k++)
        if (n[k] < 1 || n[k] > v.cap) return fail(NIRRT_ERR_INVALID, "tree size out of range");
    if (count == 0) return NIRRT_OK;
    TRY(ensure_staging(b));
    const int chunk = b->stage_envs;
    int *dn = b->stage_n + env_begin;
    CUDA_TRY(cudaMemcpyAsync(dn, n, sizeof(int) * count, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaEventRecord(b->xfork, s));
    for (int i = 0; i < 2; i++) CUDA_TRY(cudaStreamWaitEvent(b->xs[i], b->xfork, 0));
    int ci = 0;
    for (int k0 = 0; k0 < count; k0 += chunk, ci++) {
        const int m = count - k0 < chunk ? count - k0 : chunk;
        cudaStream_t xs = b->xs[ci & 1];
        double *dv = (double *)b->stage[ci & 1];
        long long *dp = (long long *)(dv + (size_t)m * v.cap * v.dim);
        CUDA_TRY(cudaMemcpyAsync(dv, vertices + (size_t)k0 * v.cap * v.dim, sizeof(double) * v.dim * v.cap * m, cudaMemcpyHostToDevice, xs));
        CUDA_TRY(cudaMemcpyAsync(dp, parents + (size_t)k0 * v.cap, sizeof(long long) * v.cap * m, cudaMemcpyHostToDevice, xs));
        const int gx = (v.cap + 255) / 256 < 256 ? (v.cap + 255) / 256 : 256;
        k_scatter_trees<<<dim3(gx, m), 256, 0, xs>>>(v, env_begin + k0, dn + k0, dv, dp);
        LAUNCH_D(v.dim, k_build_links, dim3(gx, m), 256, 0, xs, v, env_begin + k0, dn + k0);
        CHECK_LAUNCH();
    }
    for (int i = 0; i < 2; i++) {
        CUDA_TRY(cudaEventRecord(b->xe[i], b->xs[i]));
        CUDA_TRY(cudaStreamWaitEvent(s, b->xe[i], 0));
    }
    CUDA_TRY(cudaStreamSynchronize(s));
    return NIRRT_OK;
}

extern "C" int nirrt_batch_read_trees_sync(nirrt_batch *b, int env_begin, int count, int *n,
                                           double *vertices, int64_t *parents, void *stream) {
    if (!b || !n || !vertices || !parents) return fail(NIRRT_ERR_INVALID, "nirrt_batch_read_trees_sync: null argument");
    View &v = b->v;
    if (env_begin < 0 || count < 0 || env_begin + count > v.E) return fail(NIRRT_ERR_INVALID, "env range out of bounds");
    CUDA_TRY(cudaSetDevice(b->device));
    cudaStream_t s = (cudaStream_t)stream;
    CUDA_TRY(cudaMemcpyAsync(b->h_ctl, v.ctl, sizeof(EnvCtl) * v.E, cudaMemcpyDeviceToHost, s));
    if (count == 0) { CUDA_TRY(cudaStreamSynchronize(s)); return NIRRT_OK; }
    TRY(ensure_staging(b));
    const int chunk = b->stage_envs;
    CUDA_TRY(cudaEventRecord(b->xfork, s));
    for (int i = 0; i < 2; i++) CUDA_TRY(cudaStreamWaitEvent(b->xs[i], b->xfork, 0));
    int ci = 0;
    for (int k0 = 0; k0 < count; k0 += chunk, ci++) {
        const int m = count - k0 < chunk ? count - k0 : chunk;
        cudaStream_t xs = b->xs[ci & 1];
        double *dv = (double *)b->stage[ci & 1];
        long long *dp = (long long *)(dv + (size_t)m * v.cap * v.dim);
        const int gx = (v.cap + 255) / 256 < 256 ? (v.cap + 255) / 256 : 256;
        k_gather_trees<<<dim3(gx, m), 256, 0, xs>>>(v, env_begin + k0, dv, dp);
        CHECK_LAUNCH();
        CUDA_TRY(cudaMemcpyAsync(vertices + (size_t)k0 * v.cap * v.dim, dv, sizeof(double) * v.dim * v.cap * m, cudaMemcpyDeviceToHost, xs));
        CUDA_TRY(cudaMemcpyAsync(parents + (size_t)k0 * v.cap, dp, sizeof(long long) * v.cap * m, cudaMemcpyDeviceToHost, xs));
    }
    for (int i = 0; i < 2; i++) {
        CUDA_TRY(cudaEventRecord(b->xe[i], b->xs[i]));
        CUDA_TRY(cudaStreamWaitEvent(s, b->xe[i], 0));
    }
    CUDA_TRY(cudaStreamSynchronize(s));
    for (int k = 0; k < count; k++) n[k] = b->h_ctl[env_begin + k].n;
    return NIRRT_OK;
}

static bool can_pipeline(const nirrt_batch *b);
static nirrt_batch::GraphEntry *ensure_graph(nirrt_batch *b, bool pipelined);

extern "C" int nirrt_batch_begin(nirrt_batch *b, int variant, int mode, int iter_max, int iter_after_initial, void *stream) {
    if (!b) return fail(NIRRT_ERR_INVALID, "null batch");
    if (variant < 0 || variant > 3 || mode < 0 || mode > 1 || iter_max < 0 || iter_after_initial < 0)
        return fail(NIRRT_ERR_INVALID, "nirrt_batch_begin: bad variant/mode/iteration counts");
    View &v = b->v;
    CUDA_TRY(cudaSetDevice(b->device));
    cudaStream_t s = (cudaStream_t)stream;
    v.variant = variant; v.mode = mode; b->cfg.iter_max = iter_max; b->cfg.iter_after = iter_after_initial;
    b->cfg.stop_below = (double)INFINITY;
    k_begin<<<(v.E + 127) / 128, 128, 0, s>>>(v);
    CHECK_LAUNCH();
    if (fam_informed(variant) || mode == NIRRT_MODE_PLANNING_RANDOM) {      // drivers that track the goal candidates
        TRY(ensure_goal_lists(b));
        LAUNCH_D(v.dim, k_goal_init, v.E, 256, 0, s, v);
        CHECK_LAUNCH();
    }
    // one-time cost of this (variant, mode): capture + instantiate + upload its iteration graph HERE, so that no
    // nirrt_batch_run ever builds one in its steady state (a later begin() with the same variant/mode finds it)
    if (!b->persistent) ensure_graph(b, can_pipeline(b));
    return NIRRT_OK;
}

// One lock-step iteration of problems [env0, env0 + count) on stream s.  Every iteration is the same two launches
// (scan + k_expand, whose tail draws the next sample): the run's first sample comes from one k_top launch in
// nirrt_batch_run, and the tail of the run's last k_expand finds the iteration budget exhausted and idles the
// problem (top_body) -- so a run of any length is k_top + uniform iterations, and whole blocks of them replay
// from one CUDA graph.
static int launch_iteration(nirrt_batch *b, cudaStream_t s, int env0, int count) {
    View v = b->v;
    v.env0 = env0;
    v.fuse_top = 1;
    const bool pdl = b->pdl;
    const bool mirror = has_mirror(v);   // mirror scans collect Near speculatively during the Nearest pass:
    v.fuse_steer = mirror ? 1 : 0;       // the iteration is two kernels, the scan and everything else
    launch_scan<false>(v, 0, count, s, pdl);
    if (!mirror) {
        launch_view(v.dim == 3 ? k_steer<3> : k_steer<2>, count, 32, s, v, pdl);
        launch_scan<false>(v, 1, count, s, pdl);
    }
    launch_view(v.dim == 3 ? k_expand<3> : k_expand<2>, count, kExpandThreads, s, v, pdl);
    b->launches += mirror ? 2 : 4;
    return NIRRT_OK;
}

// the sample of a run's first iteration, all problems in one launch
static void launch_top(nirrt_batch *b, cudaStream_t s) {
    View v = b->v;
    v.env0 = 0;
    launch_view(v.dim == 3 ? k_top<3> : k_top<2>, v.E, 128, s, v, false);
    b->launches += 1;
}

// `n` iterations of every group: fork from `s` to the group streams, launch, join back into `s`
static int run_groups(nirrt_batch *b, cudaStream_t s, int n) {
    const View &v = b->v;
    if (n <= 0) return NIRRT_OK;
    if (b->groups == 1) {
        for (int it = 0; it < n; it++) launch_iteration(b, s, 0, v.E);
        return NIRRT_OK;
    }
    const int G = b->groups;
    CUDA_TRY(cudaEventRecord(b->ev_fork, s));
    for (int g = 0; g < G; g++) CUDA_TRY(cudaStreamWaitEvent(b->gs[g], b->ev_fork, 0));
    for (int it = 0; it < n; it++)
        for (int g = 0; g < G; g++) {
            const int e0 = (int)((long long)v.E * g / G), e1 = (int)((long long)v.E * (g + 1) / G);
            if (e1 > e0) launch_iteration(b, b->gs[g], e0, e1 - e0);
        }
    for (int g = 0; g < G; g++) {
        CUDA_TRY(cudaEventRecord(b->ev_join[g], b->gs[g]));
        CUDA_TRY(cudaStreamWaitEvent(s, b->ev_join[g], 0));
    }
    return NIRRT_OK;
}

// The pipelined RRT* driver applies to the plain planning() loop body of RRT* with a mirror scan: nothing the
// next sample / scan needs depends on ChooseParent / Rewire there (see k_front).
static bool can_pipeline(const nirrt_batch *b) {
    const View &v = b->v;
    return b->pipeline && b->groups >= 2 && has_mirror(v) && v.variant == 0 && v.mode == NIRRT_MODE_PLANNING;
}

// `m` (even) pipelined iterations of every group.  Per group, stream gs: scan(j) -> k_front(j) -> scan(j+1) ...,
// stream gs2: k_expand(j) after k_front(j); scan(j+2) waits for k_expand(j), whose IterScratch copy and
// candidate-list half it re-uses.  Entered and left with everything joined into `s` and the sample of the
// next iteration ready on copy 0.
static int run_pipelined(nirrt_batch *b, cudaStream_t s, int m) {
    if (m <= 0) return NIRRT_OK;
    const int G = b->groups;
    CUDA_TRY(cudaEventRecord(b->ev_fork, s));
    for (int g = 0; g < G; g++) CUDA_TRY(cudaStreamWaitEvent(b->gs[g], b->ev_fork, 0));
    for (int j = 0; j < m; j++)
        for (int g = 0; g < G; g++) {
            View v = b->v;
            const int e0 = (int)((long long)v.E * g / G), e1 = (int)((long long)v.E * (g + 1) / G);
            if (e1 <= e0) continue;
            const int count = e1 - e0, q = j & 1;
            v.env0 = e0; v.par = q; v.pipe = 1; v.fuse_steer = 0; v.fuse_top = 0;
            if (j >= 2) CUDA_TRY(cudaStreamWaitEvent(b->gs[g], b->ev_b[g][q], 0));
            launch_scan<false>(v, 0, count, b->gs[g], false);               // follows event operations: plain stream order
            launch_view(v.dim == 3 ? k_front<3> : k_front<2>, count, 128, b->gs[g], v, b->pdl);
            CUDA_TRY(cudaEventRecord(b->ev_a[g][q], b->gs[g]));
            CUDA_TRY(cudaStreamWaitEvent(b->gs2[g], b->ev_a[g][q], 0));
            launch_view(v.dim == 3 ? k_expand<3> : k_expand<2>, count, kExpandThreads, b->gs2[g], v, false);
            CUDA_TRY(cudaEventRecord(b->ev_b[g][q], b->gs2[g]));
            b->launches += 3;
        }
    for (int g = 0; g < G; g++) {
        if ((long long)b->v.E * (g + 1) / G <= (long long)b->v.E * g / G) continue;
        CUDA_TRY(cudaEventRecord(b->ev_join[g], b->gs[g]));
        CUDA_TRY(cudaStreamWaitEvent(s, b->ev_join[g], 0));
        CUDA_TRY(cudaStreamWaitEvent(s, b->ev_b[g][(m - 1) & 1], 0));        // the last k_expand on gs2 (stream ordered)
    }
    return NIRRT_OK;
}

// CUDA graph of b->graph_iters iterations of all groups: one graph launch replaces 2 x groups x graph_iters
// kernel launches, which keeps the host far ahead of the device even with many small groups (and, for a single
// small group, removes the per-launch cost from the latency-bound small-tree regime).  Kernel arguments are the
// View by value, so there is one executable per distinct View -- in practice per (variant, mode): everything a
// run changes (iteration counts, thresholds, vertex limit, guidance knobs) lives in EnvCtl::cfg.  Graphs are
// built by nirrt_batch_begin (never inside a run that finds its graph) and kept for the life of the batch.
constexpr size_t kMaxGraphs = 8;
// what the captured kernels read of the View: the goal-candidate buffers (allocated lazily, possibly after other
// graphs were built) only matter to the RRT* eval driver
static void graph_key(const View &v, View *k) {
    memcpy(k, &v, sizeof(View));     // byte copies throughout: the lookup is a memcmp
    k->occ = nullptr; k->occ_h = 0; k->occ_w = 0;     // the 2D free-space masks are only read by the guidance-cloud kernels
    if (!(v.mode == NIRRT_MODE_PLANNING_RANDOM || fam_informed(v.variant))) { k->gc_idx = nullptr; k->gc_d = nullptr; k->gc_cost = nullptr; k->kid = nullptr; k->gc_stride = 0; }
}
static nirrt_batch::GraphEntry *ensure_graph(nirrt_batch *b, bool pipelined) {
    if (!b->use_graph || b->graph_iters < 2) return nullptr;
    if (!b->cs) {
        if (cudaStreamCreateWithFlags(&b->cs, cudaStreamNonBlocking) != cudaSuccess) { b->use_graph = false; b->graph_fallbacks++; return nullptr; }
    }
    View key;
    graph_key(b->v, &key);
    for (auto &g : b->graphs)
        if (g.pipelined == pipelined && memcmp(&g.view, &key, sizeof(View)) == 0) { g.last_use = ++b->graph_clock; return &g; }
    if (b->graphs.size() >= kMaxGraphs) {       // evict the least recently used executable
        size_t lru = 0;
        for (size_t i = 1; i < b->graphs.size(); i++) if (b->graphs[i].last_use < b->graphs[lru].last_use) lru = i;
        cudaGraphExecDestroy(b->graphs[lru].exec);
        b->graphs.erase(b->graphs.begin() + lru);
    }
    const int64_t launches0 = b->launches;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    bool ok = cudaStreamBeginCapture(b->cs, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    if (ok) {
        const int rc = pipelined ? run_pipelined(b, b->cs, b->graph_iters) : run_groups(b, b->cs, b->graph_iters);
        const cudaError_t e = cudaStreamEndCapture(b->cs, &graph);
        ok = rc == NIRRT_OK && e == cudaSuccess && graph != nullptr;
    }
    if (ok) ok = cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
    if (ok) ok = cudaGraphUpload(exec, b->cs) == cudaSuccess && cudaStreamSynchronize(b->cs) == cudaSuccess;
    if (graph) cudaGraphDestroy(graph);
    const int64_t per_replay = b->launches - launches0;     // kernels per replay
    b->launches = launches0;
    if (!ok) {              // capture is an optimisation only: plain launches from here on, and the counter says so
        cudaGetLastError();
        if (exec) cudaGraphExecDestroy(exec);
        b->use_graph = false;
        b->graph_fallbacks++;
        fprintf(stderr, "libnirrt_b200: CUDA graph capture failed; this batch falls back to plain kernel launches\n");
        return nullptr;
    }
    nirrt_batch::GraphEntry g;
    memcpy(&g.view, &key, sizeof(View));      // byte copy: the lookup above is a memcmp
    g.pipelined = pipelined; g.exec = exec; g.launches = per_replay; g.last_use = ++b->graph_clock;
    b->graphs.push_back(g);
    b->graph_builds++;
    return &b->graphs.back();
}

extern "C" int nirrt_batch_run(nirrt_batch *b, int iters, void *stream) {
    if (!b || iters < 0) return fail(NIRRT_ERR_INVALID, "nirrt_batch_run: bad argument");
    View &v = b->v;
    CUDA_TRY(cudaSetDevice(b->device));
    cudaStream_t s = (cudaStream_t)stream;
    k_set_budget<<<(v.E + 127) / 128, 128, 0, s>>>(v, iters, b->cfg);
    if (iters == 0) { CHECK_LAUNCH(); return NIRRT_OK; }
    const int GI = b->graph_iters;
    launch_top(b, s);
    if (b->persistent && v.ux) {
        View w = v;
        w.env0 = 0; w.fuse_top = 1; w.fuse_steer = 1;
        if (v.dim == 3) k_iterate<3><<<v.E, kExpandThreads, 0, s>>>(w, iters);
        else k_iterate<2><<<v.E, kExpandThreads, 0, s>>>(w, iters);
        b->launches += 1;
        CHECK_LAUNCH();
        return NIRRT_OK;
    }
    if (can_pipeline(b) && iters >= 6) {
        // first and last iteration unpipelined (the first consumes k_top's sample, the last k_expand finds the budget
        // exhausted), an even number of pipelined iterations in between
        TRY(run_groups(b, s, 1));
        const int mid = iters - 2, m = mid & ~1;
        int done = 0;
        if (GI >= 2 && m >= GI) {
            if (nirrt_batch::GraphEntry *g = ensure_graph(b, true))
                for (; done + GI <= m; done += GI) { CUDA_TRY(cudaGraphLaunch(g->exec, s)); b->launches += g->launches; b->graph_replays++; }
        }
        TRY(run_pipelined(b, s, m - done));
        TRY(run_groups(b, s, mid - m + 1));
    } else {
        int done = 0;
        if (GI >= 2 && iters >= GI) {
            if (nirrt_batch::GraphEntry *g = ensure_graph(b, false))
                for (; done + GI <= iters; done += GI) { CUDA_TRY(cudaGraphLaunch(g->exec, s)); b->launches += g->launches; b->graph_replays++; }
        }
        TRY(run_groups(b, s, iters - done));
    }
    CHECK_LAUNCH();
    return NIRRT_OK;
}

static int fetch_ctl(nirrt_batch *b, cudaStream_t s);
extern "C" int nirrt_batch_work_stats_sync(nirrt_batch *b, int64_t *out8, void *stream) {
    if (!b || !out8) return fail(NIRRT_ERR_INVALID, "nirrt_batch_work_stats_sync: null argument");
    TRY(fetch_ctl(b, (cudaStream_t)stream));
    for (int k = 0; k < 8; k++) out8[k] = 0;
    for (int e = 0; e < b->v.E; e++)
        for (int k = 0; k < 8; k++) out8[k] += (int64_t)b->h_ctl[e].work[k];
    return NIRRT_OK;
}

#ifdef NIRRT_PHASE_TIMING
// development build only: mean SM clocks per k_expand phase since the last call (see g_phase)
extern "C" int nirrt_debug_phase_clocks(double *out9) {
    unsigned long long h[16];
    CUDA_TRY(cudaMemcpyFromSymbol(h, g_phase, sizeof(h)));
    for (int k = 0; k < 8; k++) out9[k] = h[8] ? (double)h[k] / (double)h[8] : 0.0;
    out9[8] = (double)h[8];
    memset(h, 0, sizeof(h));
    CUDA_TRY(cudaMemcpyToSymbol(g_phase, h, sizeof(h)));
    return NIRRT_OK;
}
#endif

extern "C" int nirrt_batch_graph_stats(nirrt_batch *b, int64_t *builds, int64_t *replays, int64_t *fallbacks) {
    if (!b) return fail(NIRRT_ERR_INVALID, "null batch");
    if (builds) *builds = b->graph_builds;
    if (replays) *replays = b->graph_replays;
    if (fallbacks) *fallbacks = b->graph_fallbacks;
    return NIRRT_OK;
}

extern "C" int nirrt_batch_set_stop_threshold(nirrt_batch *b, double stop_below) {
    if (!b) return fail(NIRRT_ERR_INVALID, "null batch");
    b->cfg.stop_below = stop_below;
    return NIRRT_OK;
}

extern "C" int nirrt_batch_set_vertex_limit(nirrt_batch *b, int limit) {
    if (!b || limit < 0) return fail(NIRRT_ERR_INVALID, "nirrt_batch_set_vertex_limit: bad argument");
    b->cfg.n_limit = limit;
    return NIRRT_OK;
}

// Same work as nirrt_batch_run, with a CUDA-event bracket around every launch: returns the summed
// device time of each of the five kernels over `iters` iterations (roofline attribution).
extern "C" int nirrt_batch_run_profiled_sync(nirrt_batch *b, int iters, float *ms5, void *stream) {
    if (!b || iters < 1 || !ms5) return fail(NIRRT_ERR_INVALID, "nirrt_batch_run_profiled_sync: bad argument");
    View &v = b->v;
    CUDA_TRY(cudaSetDevice(b->device));
    cudaStream_t s = (cudaStream_t)stream;
    std::vector<cudaEvent_t> ev((size_t)iters * 6);
    for (auto &e : ev) CUDA_TRY(cudaEventCreate(&e));
    k_set_budget<<<(v.E + 127) / 128, 128, 0, s>>>(v, iters, b->cfg);
    for (int it = 0; it < iters; it++) {
        cudaEvent_t *e = ev.data() + (size_t)it * 6;
        cudaEventRecord(e[0], s);
        LAUNCH_D(v.dim, k_top, v.E, 128, 0, s, v);
        cudaEventRecord(e[1], s);
        launch_scan<false>(v, 0, v.E, s);
        cudaEventRecord(e[2], s);
        LAUNCH_D(v.dim, k_steer, v.E, 32, 0, s, v);
        cudaEventRecord(e[3], s);
        if (!has_mirror(v)) launch_scan<false>(v, 1, v.E, s);   // mirror scans: Near is collected by the Nearest pass
        cudaEventRecord(e[4], s);
        LAUNCH_D(v.dim, k_expand, v.E, kExpandThreads, 0, s, v);
        cudaEventRecord(e[5], s);
    }
    b->launches += 5LL * iters;
    CHECK_LAUNCH();
    CUDA_TRY(cudaStreamSynchronize(s));
    for (int k = 0; k < 5; k++) ms5[k] = 0.f;
    for (int it = 0; it < iters; it++)
        for (int k = 0; k < 5; k++) {
            float t = 0.f;
            CUDA_TRY(cudaEventElapsedTime(&t, ev[(size_t)it * 6 + k], ev[(size_t)it * 6 + k + 1]));
            ms5[k] += t;
        }
    for (auto &e : ev) cudaEventDestroy(e);
    return NIRRT_OK;
}

static int fetch_ctl(nirrt_batch *b, cudaStream_t s) {
    CUDA_TRY(cudaSetDevice(b->device));
    CUDA_TRY(cudaMemcpyAsync(b->h_ctl, b->v.ctl, sizeof(EnvCtl) * b->v.E, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return NIRRT_OK;
}

static std::string err_bits(int err) {
    std::string m;
    if (err & ERR_NEAR_OVERFLOW) m += " near-candidate buffer overflow;";
    if (err & ERR_SOL_OVERFLOW) m += " path_solutions overflow;";
    if (err & ERR_VERTEX_OVERFLOW) m += " vertex capacity exceeded;";
    if (err & ERR_EMPTY_CLOUD) m += " SamplePointCloud on an empty predicted cloud (reference raises ValueError);";
    if (err & ERR_PATH_DEPTH) m += " path deeper than 4096 edges;";
    if (err & ERR_RECORD_OVERFLOW) m += " record buffer overflow;";
    if (err & ERR_GOAL_OVERFLOW) m += " goal-candidate overflow;";
    if (err & ERR_CHILD_LISTS) m += " internal: child lists inconsistent (goal tracking fell back to full evaluation);";
    if (err & ERR_OUT_OF_RANGE) m += " a loaded vertex lies outside the world range (the mirror scan margin assumes vertices inside it);";
    return m;
}

extern "C" int nirrt_batch_status_sync(nirrt_batch *b, int *running, int *need_cloud, void *stream) {
    if (!b) return fail(NIRRT_ERR_INVALID, "null batch");
    TRY(fetch_ctl(b, (cudaStream_t)stream));
    int run = 0, need = 0, err = 0, err_env = -1;
    for (int e = 0; e < b->v.E; e++) {
        const EnvCtl &c = b->h_ctl[e];
        if (c.state == ST_PHASE1 || c.state == ST_PHASE2) run++;
        if (c.state == ST_WAIT_CLOUD) need++;
        if (c.err && !err) { err = c.err; err_env = e; }
    }
    if (running) *running = run;
    if (need_cloud) *need_cloud = need;
    if (err) return fail(NIRRT_ERR_CAPACITY, "problem " + std::to_string(err_env) + ":" + err_bits(err));
    return NIRRT_OK;
}

extern "C" int nirrt_batch_env_state_sync(nirrt_batch *b, int *state, int *n_records, int *n_vertices, void *stream) {
    if (!b) return fail(NIRRT_ERR_INVALID, "null batch");
    TRY(fetch_ctl(b, (cudaStream_t)stream));
    for (int e = 0; e < b->v.E; e++) {
        if (state) state[e] = b->h_ctl[e].state;
        if (n_records) n_records[e] = b->h_ctl[e].n_rec;
        if (n_vertices) n_vertices[e] = b->h_ctl[e].n;
    }
    return NIRRT_OK;
}

extern "C" int nirrt_batch_read_cbest_sync(nirrt_batch *b, double *c_best, double *c_min, void *stream) {
    if (!b) return fail(NIRRT_ERR_INVALID, "null batch");
    TRY(fetch_ctl(b, (cudaStream_t)stream));
    for (int e = 0; e < b->v.E; e++) {
        if (c_best) c_best[e] = b->h_ctl[e].c_best;
        if (c_min) c_min[e] = b->h_ctl[e].c_min;
    }
    return NIRRT_OK;
}

extern "C" int nirrt_batch_read_records_sync(nirrt_batch *b, int env_begin, int count, double *records, int *n_records, void *stream) {
    if (!b || !records || !n_records) return fail(NIRRT_ERR_INVALID, "nirrt_batch_read_records_sync: null argument");
    View &v = b->v;
    if (env_begin < 0 || count < 0 || env_begin + count > v.E) return fail(NIRRT_ERR_INVALID, "env range out of bounds");
    cudaStream_t s = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(b->device));
    CUDA_TRY(cudaMemcpyAsync(records, v.records + (size_t)env_begin * v.rec_cap, sizeof(double) * (size_t)count * v.rec_cap, cudaMemcpyDeviceToHost, s));
    TRY(fetch_ctl(b, s));
    for (int k = 0; k < count; k++) n_records[k] = b->h_ctl[env_begin + k].n_rec < v.rec_cap ? b->h_ctl[env_begin + k].n_rec : v.rec_cap;
    return NIRRT_OK;
}

extern "C" int nirrt_batch_read_solutions_sync(nirrt_batch *b, int env, int64_t *out, int cap, void *stream) {
    if (!b || env < 0 || env >= b->v.E) return fail(NIRRT_ERR_INVALID, "nirrt_batch_read_solutions_sync: bad argument");
    View &v = b->v;
    cudaStream_t s = (cudaStream_t)stream;
    TRY(fetch_ctl(b, s));
    int n = b->h_ctl[env].n_sol;
    if (n > v.sol_cap) n = v.sol_cap;
    if (out && cap > 0 && n > 0) {
        const int m = n < cap ? n : cap;
        std::vector<int> h(m);
        CUDA_TRY(cudaMemcpyAsync(h.data(), v.sol + (size_t)env * v.sol_cap, sizeof(int) * m, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        for (int i = 0; i < m; i++) out[i] = h[i];
    }
    return n;
}

extern "C" int nirrt_batch_goal_parent_sync(nirrt_batch *b, int64_t *goal_parent, double *cost, void *stream) {
    if (!b || !goal_parent || !cost) return fail(NIRRT_ERR_INVALID, "nirrt_batch_goal_parent_sync: null argument");
    View &v = b->v;
    cudaStream_t s = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(b->device));
    const int use_solutions = fam_informed(v.variant) ? 1 : 0;
    if (!use_solutions) {
        TRY(ensure_goal_lists(b));
        LAUNCH_D(v.dim, k_goal_init, v.E, 256, 0, s, v);
        CHECK_LAUNCH();
    }
    TempBufs t;
    long long *dgp; double *dc;
    TRY(t.make<long long>(v.E, &dgp)); TRY(t.make<double>(v.E, &dc));
    LAUNCH_D(v.dim, k_goal_parent, v.E, kExpandThreads, 0, s, v, use_solutions, dgp, dc);
    CHECK_LAUNCH();
    CUDA_TRY(cudaMemcpyAsync(goal_parent, dgp, sizeof(long long) * v.E, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(cost, dc, sizeof(double) * v.E, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return NIRRT_OK;
}

extern "C" int nirrt_batch_read_trace_sync(nirrt_batch *b, int *nearest, int *new_index, int *near_count,
                                           int *near, int near_stride, double *x_rand, void *stream) {
    if (!b) return fail(NIRRT_ERR_INVALID, "null batch");
    View &v = b->v;
    cudaStream_t s = (cudaStream_t)stream;
    TRY(fetch_ctl(b, s));
    std::vector<int> h;
    if (near && near_stride > 0) {
        h.resize((size_t)v.E * v.near_cap);
        CUDA_TRY(cudaMemcpyAsync(h.data(), v.near_out, sizeof(int) * h.size(), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
    }
    for (int e = 0; e < v.E; e++) {
        const EnvCtl &c = b->h_ctl[e];
        const IterScratch &it = c.s[0];         // runs end with a non-pipelined iteration on copy 0
        if (nearest) nearest[e] = it.nearest;
        if (new_index) new_index[e] = it.new_idx;
        if (near_count) near_count[e] = it.near_cnt;
        if (x_rand) for (int i = 0; i < v.dim; i++) x_rand[v.dim * e + i] = c.x_rand[i];
        if (near && near_stride > 0) {
            const int m = it.near_cnt < near_stride ? it.near_cnt : near_stride;
            for (int k = 0; k < m; k++) near[(size_t)e * near_stride + k] = h[(size_t)e * v.near_cap + k];
        }
    }
    return NIRRT_OK;
}

// ---- stand-alone predicates --------------------------------------------------------------------
extern "C" int nirrt_collide_edges_sync(nirrt_batch *b, int env, const double *edges, int64_t m, uint8_t *out, void *stream) {
    if (!b || env < 0 || env >= b->v.E || m < 0 || (m > 0 && (!edges || !out))) return fail(NIRRT_ERR_INVALID, "nirrt_collide_edges_sync: bad argument");
    if (m == 0) return NIRRT_OK;
    cudaStream_t s = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(b->device));
    TempBufs t;
    double *de; uint8_t *dout;
    TRY(t.up(edges, 2 * b->v.dim * (size_t)m, s, &de)); TRY(t.make<uint8_t>((size_t)m, &dout));
    const int blocks = (int)((m + 255) / 256 < 148 * 8 ? (m + 255) / 256 : 148 * 8);
    LAUNCH_D(b->v.dim, k_collide_edges, blocks, 256, 0, s, b->v, env, de, m, dout);
    CHECK_LAUNCH();
    CUDA_TRY(cudaMemcpyAsync(out, dout, (size_t)m, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return NIRRT_OK;
}

extern "C" int nirrt_points_check_sync(nirrt_batch *b, int env, int kind, const double *points, int64_t m, uint8_t *out, void *stream) {
    if (!b || env < 0 || env >= b->v.E || m < 0 || kind < 0 || kind > 1 || (m > 0 && (!points || !out)))
        return fail(NIRRT_ERR_INVALID, "nirrt_points_check_sync: bad argument");
    if (m == 0) return NIRRT_OK;
    cudaStream_t s = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(b->device));
    TempBufs t;
    double *dp; uint8_t *dout;
    TRY(t.up(points, b->v.dim * (size_t)m, s, &dp)); TRY(t.make<uint8_t>((size_t)m, &dout));
    const int blocks = (int)((m + 255) / 256 < 148 * 8 ? (m + 255) / 256 : 148 * 8);
    LAUNCH_D(b->v.dim, k_points_check, blocks, 256, 0, s, b->v, env, kind, dp, m, dout);
    CHECK_LAUNCH();
    CUDA_TRY(cudaMemcpyAsync(out, dout, (size_t)m, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return NIRRT_OK;
}

// single-env views for the stand-alone scans: reuse the batch kernels on a 1-env grid
__global__ void k_set_query(View v, int env, const double *q, int which, double r) {
    EnvCtl *c = v.ctl + env;
    const double qz = v.dim == 3 ? q[2] : 0.0;
    if (which == 0) { c->x_rand[0] = q[0]; c->x_rand[1] = q[1]; c->x_rand[2] = qz; }
    else {
        IT.x_new[0] = q[0]; IT.x_new[1] = q[1]; IT.x_new[2] = qz;
        IT.r = r;
        IT.T_near = v.dim == 3 ? sqrt_le_threshold(r) : hypot_band_sq(r);
        c->cand_cnt = 0;
    }
}
__global__ void k_finish_nearest(View v, int env, long long *out) {
    double bs = XINF; int bi = INT_MAX;
    for (int k = threadIdx.x; k < v.chunks; k += 32) lexmin(bs, bi, v.part_s[(size_t)env * v.chunks + k], v.part_i[(size_t)env * v.chunks + k]);
    warp_lexmin(bs, bi);
    if (threadIdx.x == 0) *out = bi;
}

static View single_env_view(const View &v, int env) {
    View w = v;   // shift every per-env array so that blockIdx.y == 0 addresses `env`
    w.vx += (size_t)env * v.stride; w.vy += (size_t)env * v.stride;
    if (v.vz) w.vz += (size_t)env * v.stride;
    w.nodes += (size_t)env * v.stride; w.hints += (size_t)env * v.stride; w.links += (size_t)env * v.stride;
    if (v.geom) w.geom += env;
    if (v.geom2) w.geom2 += env;
    if (v.mt_py) w.mt_py += env;
    w.mt += env; w.ctl += env;
    w.part_s += (size_t)env * v.chunks; w.part_i += (size_t)env * v.chunks;
    w.cand += (size_t)env * v.near_cap; w.near_out += (size_t)env * v.near_cap; w.cand2 += (size_t)env * v.near_cap;   /* single-env views always use copy 0 */
    w.E = 1;
    return w;
}

extern "C" int nirrt_nearest_sync(nirrt_batch *b, int env, const double *queries, int64_t m, int64_t *out, void *stream) {
    if (!b || env < 0 || env >= b->v.E || m < 0 || (m > 0 && (!queries || !out))) return fail(NIRRT_ERR_INVALID, "nirrt_nearest_sync: bad argument");
    if (m == 0) return NIRRT_OK;
    cudaStream_t s = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(b->device));
    TempBufs t;
    double *dq; long long *dout;
    TRY(t.up(queries, b->v.dim * (size_t)m, s, &dq)); TRY(t.make<long long>((size_t)m, &dout));
    View w = single_env_view(b->v, env);
    for (int64_t k = 0; k < m; k++) {
        k_set_query<<<1, 1, 0, s>>>(w, 0, dq + w.dim * k, 0, 0.0);
        LAUNCH_DB(w.dim, k_nearest, true, dim3(w.chunks, 1), 256, 0, s, w);
        k_finish_nearest<<<1, 32, 0, s>>>(w, 0, dout + k);
    }
    CHECK_LAUNCH();
    CUDA_TRY(cudaMemcpyAsync(out, dout, sizeof(long long) * m, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return NIRRT_OK;
}

extern "C" int64_t nirrt_within_sync(nirrt_batch *b, int env, const double *q, double r, int64_t *out, int64_t cap, void *stream) {
    if (!b || env < 0 || env >= b->v.E || !q || r < 0 || (cap > 0 && !out)) return fail(NIRRT_ERR_INVALID, "nirrt_within_sync: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(b->device));
    TempBufs t;
    double *dq;
    TRY(t.up(q, (size_t)b->v.dim, s, &dq));
    View w = single_env_view(b->v, env);
    k_set_query<<<1, 1, 0, s>>>(w, 0, dq, 1, r);
    LAUNCH_DB(w.dim, k_near, true, dim3(w.chunks, 1), 256, 0, s, w);
    CHECK_LAUNCH();
    TRY(fetch_ctl(b, s));
    const int cnt = b->h_ctl[env].cand_cnt;
    if (cnt > b->v.near_cap) return fail(NIRRT_ERR_CAPACITY, "nirrt_within_sync: more matches than near_capacity");
    std::vector<int> h(cnt > 0 ? cnt : 1);
    if (cnt > 0) {
        CUDA_TRY(cudaMemcpyAsync(h.data(), b->v.cand + (size_t)env * b->v.near_cap, sizeof(int) * cnt, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        std::vector<int> sorted(h.begin(), h.begin() + cnt);
        // marshalling only: the device appends matches unordered, callers expect ascending order
        for (int i = 1; i < cnt; i++) { int x = sorted[i], j = i - 1; while (j >= 0 && sorted[j] > x) { sorted[j + 1] = sorted[j]; j--; } sorted[j + 1] = x; }
        for (int i = 0; i < cnt && i < cap; i++) out[i] = sorted[i];
    }
    return cnt;
}

extern "C" int nirrt_costs_sync(nirrt_batch *b, int env, const int64_t *idx, int64_t m, double *out, void *stream) {
    if (!b || env < 0 || env >= b->v.E || m < 0 || (m > 0 && (!idx || !out))) return fail(NIRRT_ERR_INVALID, "nirrt_costs_sync: bad argument");
    if (m == 0) return NIRRT_OK;
    cudaStream_t s = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(b->device));
    TRY(fetch_ctl(b, s));
    for (int64_t i = 0; i < m; i++)
        if (idx[i] < 0 || idx[i] >= b->h_ctl[env].n) return fail(NIRRT_ERR_INVALID, "vertex index out of range");
    TempBufs t;
    long long *di; double *dout;
    TRY(t.up((const long long *)idx, (size_t)m, s, &di)); TRY(t.make<double>((size_t)m, &dout));
    LAUNCH_D(b->v.dim, k_costs, (int)((m + 127) / 128), 128, 0, s, b->v, env, di, m, dout);
    CHECK_LAUNCH();
    CUDA_TRY(cudaMemcpyAsync(out, dout, sizeof(double) * m, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return NIRRT_OK;
}

// ------------------------------------------------------------------------------------------------
// Guidance-cloud generation on the device (3D worlds):
//   generate_rectangle_point_cloud_3d     datasets_3d/point_cloud_mask_utils_3d.py:83-113   (kind 0)
//   ellipsoid_point_cloud_sampling_3d     datasets_3d/point_cloud_mask_utils_3d.py:132-200  (kind 1)
//   get_point_cloud_mask_around_points    datasets/point_cloud_mask_utils.py:20-31
// for a LIST of problems at once, one CTA per problem, consuming each problem's own numpy MT19937 stream word for
// word: kind 0 draws np.random.uniform(low, high, size=(n_raw, 3)) (row-major fill), kind 1 three vectors of n_raw
// (radius, theta, phi).  The candidates are filtered with the reference's predicates at clearance 0
// (points_in_balls_boxes / points_validity_3d), compacted in order, and -- if more than n_points survive --
// down-sampled by farthest point sampling with open3d's semantics (start 0, f64 squared distances, first arg-max;
// third-party arithmetic, parity unpinned).  The cloud never leaves HBM: its float32 copy and the start / goal
// neighbourhood masks go straight into the PointNet++ engine's input buffers, and k_cloud_commit turns the
// network's prediction into the problem's path_point_cloud_pred (nirrt_star_png_3d.py:172).
// kind 1 parameters (host, a handful of scalars per update, evaluated with numpy exactly as the reference does):
// M = C @ L (3x3, row major) and the ellipsoid centre; the device evaluates np.dot(M, samples.T).T + centre as the
// BLAS kernel does, fma(M2, z, fma(M1, y, M0 * x)) + c (probed on 10240 columns), with glibc's sin / cos.
struct CloudWs {
    uint32_t *words;   // [E][6 * n_raw]   raw MT19937 words of this update
    double *cand;      // [E][n_raw][3]    valid candidates, in draw order
    double *cloud;     // [E][n_points][3] the sampled cloud (after down-sampling)
    int *cand_cnt;     // [E]
    int *cloud_cnt;    // [E]
    int *envs;         // [E]              device copy of the env list of the current update
    int *kind;         // [E]
    double *params;    // [E][12]
    float *pc32, *smask, *gmask;   // network inputs when the caller does not supply its own buffers
    int n_raw, n_points;
};

// 2D (datasets/point_cloud_mask_utils.py:35-73,104-174): kind 0 draws np.random.uniform([0, 0], [W, H], (n_raw, 2)); kind 1
// np.random.uniform(-1, 1, (n_raw, 2)), keeps the unit disc (np.linalg.norm(axis=1) <= 1), maps it with np.dot(C @ L, .) +
// centre.  A point is free iff the four pixels around it (astype(int) truncation, clipped to the image) are free in the
// problem's binary_mask; kind 1 additionally requires 0 <= x <= W, 0 <= y <= H (points_in_range, inclusive).
template <int D>
__global__ void __launch_bounds__(256) k_cloud_draw(View v, CloudWs w) {
    const int k = blockIdx.x, e = w.envs[k], tid = threadIdx.x;
    MtState *st = v.mt + e;
    uint32_t *wb = w.words + (size_t)k * 6 * w.n_raw;
    const int total = 2 * D * w.n_raw;
    __shared__ int s_pos, s_cur, s_has;
    if (tid == 0) { s_pos = st->pos; s_cur = st->cur; s_has = st->has_next; }
    __syncthreads();
    // ---- the next `total` words of the stream (mt19937_gen block by block, three barrier-separated phases)
    for (int produced = 0; produced < total;) {
        int pos = s_pos, cur = s_cur;
        if (pos == 624) {
            const uint32_t *o = st->key[cur];
            uint32_t *nx = st->key[cur ^ 1];
            if (!s_has) {
                for (int i = tid; i < 227; i += blockDim.x) nx[i] = o[i + 397] ^ mt_twist(o[i], o[i + 1]);
                __syncthreads();
                for (int i = 227 + tid; i < 454; i += blockDim.x) nx[i] = nx[i - 227] ^ mt_twist(o[i], o[i + 1]);
                __syncthreads();
                for (int i = 454 + tid; i < 623; i += blockDim.x) nx[i] = nx[i - 227] ^ mt_twist(o[i], o[i + 1]);
                __syncthreads();
                if (tid == 0) nx[623] = nx[396] ^ mt_twist(o[623], nx[0]);
            }
            __syncthreads();
            if (tid == 0) { s_cur = cur ^ 1; s_pos = 0; s_has = 0; }
            __syncthreads();
            pos = 0; cur ^= 1;
        }
        const int take = min(624 - pos, total - produced);
        const uint32_t *kk = st->key[cur];
        for (int i = tid; i < take; i += blockDim.x) wb[produced + i] = kk[pos + i];
        __syncthreads();
        if (tid == 0) s_pos = pos + take;
        produced += take;
        __syncthreads();
    }
    if (tid == 0) { st->pos = s_pos; st->cur = s_cur; st->has_next = 0; }
    // ---- candidates, validity, ordered compaction
    typedef typename GeomOf<D>::type G;
    __shared__ G g;
    stage_geom<D>(&g, v, e);
    __syncthreads();
    if (tid == 0) g.clearance = 0.0;                 // the samplers are called with clearance = 0
    __shared__ int s_wcnt[8];
    __shared__ int s_total;
    if (tid == 0) s_total = 0;
    __syncthreads();
    const int kind = w.kind[k];
    const double *P = w.params + (size_t)k * 12;
    double lo[3] = {0.0, 0.0, 0.0}, sc[3] = {0.0, 0.0, 0.0};
    for (int d = 0; d < D; d++) { lo[d] = XADD(g.range[2 * d], 0.0); sc[d] = XSUB(XSUB(g.range[2 * d + 1], 0.0), lo[d]); }
    double *cand = w.cand + (size_t)k * w.n_raw * 3;
    const int n_raw = w.n_raw;
    const unsigned char *occ = D == 2 ? v.occ + (size_t)e * v.occ_h * v.occ_w : nullptr;
    auto free_pixels = [&](const double *p) {        // np.prod(binary_mask[4 neighbours]) != 0
        const int px = (int)p[0], py = (int)p[1];    // astype(int): truncation toward zero
        bool ok = true;
        for (int a = 0; a < 2; a++)
            for (int b = 0; b < 2; b++) {
                const int x = min(max(px + a, 0), v.occ_w - 1), y = min(max(py + b, 0), v.occ_h - 1);
                ok = ok && occ[(size_t)y * v.occ_w + x] != 0;
            }
        return ok;
    };
    for (int base = 0; base < n_raw; base += blockDim.x) {
        const int i = base + tid;
        bool keep = false;
        double p[3] = {0.0, 0.0, 0.0};
        if (i < n_raw) {
            if (D == 2) {
                const double u0 = mt_double(wb[4 * i], wb[4 * i + 1]), u1 = mt_double(wb[4 * i + 2], wb[4 * i + 3]);
                if (kind == 0) {
                    p[0] = XADD(lo[0], XMUL(sc[0], u0)); p[1] = XADD(lo[1], XMUL(sc[1], u1));
                    keep = free_pixels(p);
                } else {
                    const double x = XADD(-1.0, XMUL(2.0, u0)), y = XADD(-1.0, XMUL(2.0, u1));
                    if (XSQRT(XADD(XMUL(x, x), XMUL(y, y))) <= 1.0) {
                        for (int d = 0; d < 2; d++) p[d] = XADD(XFMA(P[3 * d + 1], y, XMUL(P[3 * d], x)), P[9 + d]);
                        keep = free_pixels(p) && XSUB(g.range[0], 0.0) <= p[0] && p[0] <= XADD(XADD(g.range[0], XSUB(g.range[1], g.range[0])), 0.0) &&
                               XSUB(g.range[2], 0.0) <= p[1] && p[1] <= XADD(XADD(g.range[2], XSUB(g.range[3], g.range[2])), 0.0);
                    }
                }
            } else if (kind == 0) {
                for (int d = 0; d < 3; d++) {
                    const int q = 2 * (3 * i + d);
                    p[d] = XADD(lo[d], XMUL(sc[d], mt_double(wb[q], wb[q + 1])));
                }
                keep = !point_inside_obs(g, p);
            } else {
                const double r = mt_double(wb[2 * i], wb[2 * i + 1]);                                  // 0.0 + 1.0 * u == u
                const double th = XMUL(3.141592653589793, mt_double(wb[2 * (n_raw + i)], wb[2 * (n_raw + i) + 1]));
                const double ph = XMUL(6.283185307179586, mt_double(wb[2 * (2 * n_raw + i)], wb[2 * (2 * n_raw + i) + 1]));
                const double st_ = glibc_sin(th), ct = glibc_cos(th), sp = glibc_sin(ph), cp = glibc_cos(ph);
                const double rs = XMUL(r, st_);
                const double x = XMUL(rs, cp), y = XMUL(rs, sp), z = XMUL(r, ct);
                for (int d = 0; d < 3; d++)
                    p[d] = XADD(XFMA(P[3 * d + 2], z, XFMA(P[3 * d + 1], y, XMUL(P[3 * d], x))), P[9 + d]);
                keep = point_valid(g, p);      // (2D never gets here: point_valid(Geom2) would also test the obstacles)
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        const int wi = tid >> 5, l = tid & 31;
        if (l == 0) s_wcnt[wi] = __popc(bal);
        __syncthreads();
        int off = s_total;
        for (int q = 0; q < wi; q++) off += s_wcnt[q];
        if (keep) {
            double *o = cand + (size_t)(off + __popc(bal & ((1u << l) - 1u))) * 3;
            o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
        }
        __syncthreads();
        if (tid == 0) { int t = 0; for (int q = 0; q < 8; q++) t += s_wcnt[q]; s_total += t; }
        __syncthreads();
    }
    if (tid == 0) w.cand_cnt[k] = s_total;
}

constexpr int kFpsThreads = 1024, kFpsPPT = 16;
// farthest point down-sampling of pts[0..n) to npoint points starting at `start` (open3d semantics); all threads of a
// 1024-thread CTA; visit(it, index) is called by thread 0 for every selected point
template <typename F>
__device__ __forceinline__ void fps_f64_body(const double *pts, int n, int npoint, int start, F &&visit) {
    __shared__ double s_d[2][32];
    __shared__ int s_i[2][32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double px[kFpsPPT], py[kFpsPPT], pz[kFpsPPT], dist[kFpsPPT];
#pragma unroll
    for (int j = 0; j < kFpsPPT; j++) {
        const int i = tid + j * kFpsThreads;
        const bool in = i < n;
        px[j] = in ? pts[3 * (size_t)i] : 0.0; py[j] = in ? pts[3 * (size_t)i + 1] : 0.0; pz[j] = in ? pts[3 * (size_t)i + 2] : 0.0;
        dist[j] = in ? XINF : -1.0;
    }
    int far = start;
    for (int it = 0; it < npoint; it++) {
        if (tid == 0) visit(it, far);
        const double cx = __ldg(pts + 3 * (size_t)far), cy = __ldg(pts + 3 * (size_t)far + 1), cz = __ldg(pts + 3 * (size_t)far + 2);
        double bd = -2.0; int bi = INT_MAX;
#pragma unroll
        for (int j = 0; j < kFpsPPT; j++) {
            const double d = sq3_rows(XSUB(px[j], cx), XSUB(py[j], cy), XSUB(pz[j], cz));
            if (d < dist[j]) dist[j] = d;
            if (dist[j] > bd) { bd = dist[j]; bi = tid + j * kFpsThreads; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double od = __shfl_xor_sync(0xffffffffu, bd, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (od > bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
        }
        const int slot = it & 1;
        if (lane == 0) { s_d[slot][warp] = bd; s_i[slot][warp] = bi; }
        __syncthreads();
        bd = s_d[slot][lane]; bi = s_i[slot][lane];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double od = __shfl_xor_sync(0xffffffffu, bd, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (od > bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
        }
        far = bi;
    }
}

// The same selection with the coordinates in shared memory (n <= kFpsSmemMax points: 24 B each) and only the running
// minimum distances in registers: no register spills at 1024 threads and no global load on the critical path of a
// selection step (the chosen point's coordinates are a shared-memory broadcast).  Identical arithmetic, identical result.
constexpr int kFpsSmemMax = 9600, kFpsSmemPPT = 10;      // 9600 x 24 B = 230 400 B of dynamic shared memory
// sp: where the coordinates are read from during the selection -- the shared-memory copy (kStage: filled here), or the
// candidates themselves in L2 for the rare set larger than the staging (slower, but also without register spills)
template <int kPPT, bool kStage, typename F>
__device__ __forceinline__ void fps_f64_smem(const double *pts, int n, int npoint, int start, const double *sp_in, double *sp_stage, F &&visit) {
    __shared__ double s_d[2][32];
    __shared__ int s_i[2][32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (kStage) {
        for (int i = tid; i < 3 * n; i += blockDim.x) sp_stage[i] = pts[i];
        __syncthreads();
    }
    const double *sp = kStage ? sp_stage : sp_in;
    double dist[kPPT];
#pragma unroll
    for (int j = 0; j < kPPT; j++) dist[j] = (tid + j * kFpsThreads < n) ? XINF : -1.0;
    int far = start;
    for (int it = 0; it < npoint; it++) {
        if (tid == 0) visit(it, far);
        const double cx = sp[3 * far], cy = sp[3 * far + 1], cz = sp[3 * far + 2];
        double bd = -2.0; int bi = INT_MAX;
#pragma unroll
        for (int j = 0; j < kPPT; j++) {
            const int i = tid + j * kFpsThreads;
            if (i < n) {
                const double d = sq3_rows(XSUB(sp[3 * i], cx), XSUB(sp[3 * i + 1], cy), XSUB(sp[3 * i + 2], cz));
                if (d < dist[j]) dist[j] = d;
            }
            if (dist[j] > bd) { bd = dist[j]; bi = i; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double od = __shfl_xor_sync(0xffffffffu, bd, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (od > bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
        }
        const int slot = it & 1;
        if (lane == 0) { s_d[slot][warp] = bd; s_i[slot][warp] = bi; }
        __syncthreads();
        bd = s_d[slot][lane]; bi = s_i[slot][lane];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double od = __shfl_xor_sync(0xffffffffu, bd, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (od > bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
        }
        far = bi;
    }
}

// down-sampling + the network's inputs: pc32 [count][n_points][3], start / goal masks [count][n_points]
template <int D>
__global__ void __launch_bounds__(kFpsThreads) k_cloud_fps(View v, CloudWs w, double radius, float *pc32, float *smask, float *gmask) {
    const int k = blockIdx.x, e = w.envs[k], tid = threadIdx.x;
    const int n = w.cand_cnt[k], np_ = w.n_points;
    const double *cand = w.cand + (size_t)k * w.n_raw * 3;
    double *cloud = w.cloud + (size_t)k * np_ * 3;
    int m;
    extern __shared__ double s_fps_pts[];
    auto take = [&](int it, int idx) {
        cloud[3 * (size_t)it] = cand[3 * (size_t)idx]; cloud[3 * (size_t)it + 1] = cand[3 * (size_t)idx + 1];
        cloud[3 * (size_t)it + 2] = cand[3 * (size_t)idx + 2];
    };
    if (n > np_) {
        m = np_;
        if (n <= kFpsSmemMax) fps_f64_smem<kFpsSmemPPT, true>(cand, n, np_, 0, nullptr, s_fps_pts, take);
        else fps_f64_smem<kFpsPPT, false>(cand, n, np_, 0, cand, nullptr, take);
    } else {
        m = n;
        for (int i = tid; i < 3 * n; i += blockDim.x) cloud[i] = cand[i];
    }
    __threadfence_block();
    __syncthreads();
    if (tid == 0) w.cloud_cnt[k] = m;
    const EnvCtl *c = v.ctl + e;
    for (int i = tid; i < np_; i += blockDim.x) {
        const size_t o = (size_t)k * np_ + i;
        if (i < m) {
            const double x = cloud[3 * (size_t)i], y = cloud[3 * (size_t)i + 1], z = cloud[3 * (size_t)i + 2];
            pc32[D * o] = (float)x; pc32[D * o + 1] = (float)y;
            if (D == 3) pc32[D * o + 2] = (float)z;
            // np.linalg.norm(pc[:, None] - p, axis=2) < radius  (strict)
            smask[o] = row_norm<D>(XSUB(x, c->start[0]), XSUB(y, c->start[1]), XSUB(z, c->start[2])) < radius ? 1.f : 0.f;
            gmask[o] = row_norm<D>(XSUB(x, c->goal[0]), XSUB(y, c->goal[1]), XSUB(z, c->goal[2])) < radius ? 1.f : 0.f;
        } else {
            for (int d = 0; d < D; d++) pc32[D * o + d] = 0.f;
            smask[o] = gmask[o] = 0.f;
        }
    }
}

// path_point_cloud_pred = pc[path_pred.nonzero()[0]] (nirrt_star_png_3d.py:172) and the resume of the paused problem
__global__ void __launch_bounds__(256) k_cloud_commit(View v, CloudWs w, const long long *pred, const int *sel) {
    const int k = sel ? sel[blockIdx.x] : blockIdx.x, e = w.envs[k], tid = threadIdx.x;
    const int m = w.cloud_cnt[k], np_ = w.n_points;
    const double *cloud = w.cloud + (size_t)k * np_ * 3;
    const long long *pr = pred + (size_t)k * np_;
    double *out = v.pc + (size_t)e * v.pc_cap * 3;
    __shared__ int s_wcnt[8];
    __shared__ int s_total;
    if (tid == 0) s_total = 0;
    __syncthreads();
    for (int base = 0; base < m; base += blockDim.x) {
        const int i = base + tid;
        const bool keep = i < m && pr[i] != 0;
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        const int wi = tid >> 5, l = tid & 31;
        if (l == 0) s_wcnt[wi] = __popc(bal);
        __syncthreads();
        int off = s_total;
        for (int q = 0; q < wi; q++) off += s_wcnt[q];
        if (keep) {
            double *o = out + (size_t)(off + __popc(bal & ((1u << l) - 1u))) * 3;
            o[0] = cloud[3 * (size_t)i]; o[1] = cloud[3 * (size_t)i + 1]; o[2] = cloud[3 * (size_t)i + 2];
        }
        __syncthreads();
        if (tid == 0) { int t = 0; for (int q = 0; q < 8; q++) t += s_wcnt[q]; s_total += t; }
        __syncthreads();
    }
    if (tid == 0) {
        EnvCtl *c = v.ctl + e;
        c->n_pc = s_total;
        if (c->state == ST_WAIT_CLOUD) c->state = c->saved_state;
    }
}

static int ensure_cloud_ws(nirrt_batch *b, int n_points, int n_raw) {
    View &v = b->v;
    if (b->cloud_ws && b->cloud_ws->n_raw == n_raw && b->cloud_ws->n_points == n_points) return NIRRT_OK;
    if (b->cloud_ws) return fail(NIRRT_ERR_INVALID, "guidance-cloud workspace was created for other n_points / n_raw");
    CloudWs *w = (CloudWs *)calloc(1, sizeof(CloudWs));
    if (!w) return fail(NIRRT_ERR_CUDA, "out of host memory");
    w->n_raw = n_raw; w->n_points = n_points;
    void *p = nullptr;
    const size_t E = (size_t)v.E;
#define WS_ALLOC(field, type, count) do { int _r = dalloc(b, &p, sizeof(type) * (count)); if (_r) { free(w); return _r; } w->field = (type *)p; } while (0)
    WS_ALLOC(words, uint32_t, E * 6 * n_raw);
    WS_ALLOC(cand, double, E * 3 * n_raw);
    WS_ALLOC(cloud, double, E * 3 * n_points);
    WS_ALLOC(cand_cnt, int, E); WS_ALLOC(cloud_cnt, int, E); WS_ALLOC(envs, int, E); WS_ALLOC(kind, int, E);
    WS_ALLOC(params, double, E * 12);
    WS_ALLOC(pc32, float, E * 3 * n_points); WS_ALLOC(smask, float, E * n_points); WS_ALLOC(gmask, float, E * n_points);
#undef WS_ALLOC
    if (!v.pc) {
        int r = dalloc(b, &p, sizeof(double) * 3 * E * v.pc_cap);
        if (r) { free(w); return r; }
        v.pc = (double *)p;
    }
    b->cloud_ws = w;
    return NIRRT_OK;
}

extern "C" int nirrt_batch_set_free_masks(nirrt_batch *b, const uint8_t *masks, int height, int width, void *stream) {
    if (!b || !masks || height < 1 || width < 1) return fail(NIRRT_ERR_INVALID, "nirrt_batch_set_free_masks: bad argument");
    View &v = b->v;
    if (v.dim != 2) return fail(NIRRT_ERR_INVALID, "nirrt_batch_set_free_masks: 2D batches only");
    CUDA_TRY(cudaSetDevice(b->device));
    cudaStream_t s = (cudaStream_t)stream;
    if (v.occ && (v.occ_h != height || v.occ_w != width)) return fail(NIRRT_ERR_INVALID, "nirrt_batch_set_free_masks: mask size changed");
    if (!v.occ) {
        void *p = nullptr;
        TRY(dalloc(b, &p, (size_t)v.E * height * width));
        v.occ = (const unsigned char *)p; v.occ_h = height; v.occ_w = width;
    }
    CUDA_TRY(cudaMemcpyAsync(const_cast<unsigned char *>(v.occ), masks, (size_t)v.E * height * width, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return NIRRT_OK;
}

extern "C" int nirrt_batch_sample_clouds_sync(nirrt_batch *b, const int *envs, int count, const int *kind, const double *params,
                                              int n_points, int n_raw, double neighbor_radius, float *d_pc32, float *d_start_mask,
                                              float *d_goal_mask, int *counts, void *stream) {
    if (!b || !envs || !kind || !params || !counts) return fail(NIRRT_ERR_INVALID, "nirrt_batch_sample_clouds_sync: null argument");
    if ((d_pc32 == nullptr) != (d_start_mask == nullptr) || (d_pc32 == nullptr) != (d_goal_mask == nullptr))
        return fail(NIRRT_ERR_INVALID, "nirrt_batch_sample_clouds_sync: pass all three device buffers or none");
    View &v = b->v;
    if (v.dim == 2 && !v.occ) return fail(NIRRT_ERR_INVALID, "nirrt_batch_sample_clouds_sync: 2D batches need nirrt_batch_set_free_masks first");
    if (count < 1 || count > v.E) return fail(NIRRT_ERR_INVALID, "nirrt_batch_sample_clouds_sync: bad problem count");
    if (n_points < 1 || n_points > v.pc_cap || n_raw < n_points || n_raw > kFpsThreads * kFpsPPT)
        return fail(NIRRT_ERR_INVALID, "nirrt_batch_sample_clouds_sync: need 1 <= n_points <= 4096 and n_points <= n_raw <= 16384");
    for (int k = 0; k < count; k++)
        if (envs[k] < 0 || envs[k] >= v.E || kind[k] < 0 || kind[k] > 1) return fail(NIRRT_ERR_INVALID, "nirrt_batch_sample_clouds_sync: bad env / kind");
    CUDA_TRY(cudaSetDevice(b->device));
    cudaStream_t s = (cudaStream_t)stream;
    TRY(ensure_cloud_ws(b, n_points, n_raw));
    CloudWs &w = *b->cloud_ws;
    CUDA_TRY(cudaMemcpyAsync(w.envs, envs, sizeof(int) * count, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(w.kind, kind, sizeof(int) * count, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(w.params, params, sizeof(double) * 12 * count, cudaMemcpyHostToDevice, s));
    if (v.dim == 3) k_cloud_draw<3><<<count, 256, 0, s>>>(v, w);
    else k_cloud_draw<2><<<count, 256, 0, s>>>(v, w);
    if (!d_pc32) { d_pc32 = w.pc32; d_start_mask = w.smask; d_goal_mask = w.gmask; }
    {
        static bool attr_set = false;
        if (!attr_set) {
            CUDA_TRY(cudaFuncSetAttribute(k_cloud_fps<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFpsSmemMax * 24));
            CUDA_TRY(cudaFuncSetAttribute(k_cloud_fps<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFpsSmemMax * 24));
            attr_set = true;
        }
    }
    if (v.dim == 3) k_cloud_fps<3><<<count, kFpsThreads, kFpsSmemMax * 24, s>>>(v, w, neighbor_radius, d_pc32, d_start_mask, d_goal_mask);
    else k_cloud_fps<2><<<count, kFpsThreads, kFpsSmemMax * 24, s>>>(v, w, neighbor_radius, d_pc32, d_start_mask, d_goal_mask);
    CHECK_LAUNCH();
    b->launches += 2;
    CUDA_TRY(cudaMemcpyAsync(counts, w.cloud_cnt, sizeof(int) * count, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    b->cloud_count = count;
    return NIRRT_OK;
}

extern "C" int nirrt_batch_read_sampled_clouds_sync(nirrt_batch *b, int first, int count, double *points, void *stream) {
    if (!b || !b->cloud_ws || !points || first < 0 || count < 1 || first + count > b->cloud_count)
        return fail(NIRRT_ERR_INVALID, "nirrt_batch_read_sampled_clouds_sync: bad argument");
    const CloudWs &w = *b->cloud_ws;
    cudaStream_t s = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(b->device));
    CUDA_TRY(cudaMemcpyAsync(points, w.cloud + (size_t)first * w.n_points * 3, sizeof(double) * 3 * (size_t)count * w.n_points,
                             cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return NIRRT_OK;
}

extern "C" int nirrt_batch_commit_clouds(nirrt_batch *b, const int64_t *d_pred, const int *sel, int n_sel, void *stream) {
    if (!b || !b->cloud_ws || !d_pred || n_sel < 0 || n_sel > b->cloud_count) return fail(NIRRT_ERR_INVALID, "nirrt_batch_commit_clouds: bad argument");
    const CloudWs &w = *b->cloud_ws;
    cudaStream_t s = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(b->device));
    if (sel) {
        if (n_sel == 0) return NIRRT_OK;
        for (int i = 0; i < n_sel; i++) if (sel[i] < 0 || sel[i] >= b->cloud_count) return fail(NIRRT_ERR_INVALID, "nirrt_batch_commit_clouds: bad selection");
        CUDA_TRY(cudaMemcpyAsync(w.kind, sel, sizeof(int) * n_sel, cudaMemcpyHostToDevice, s));     // kind[] is free again: selection list
        k_cloud_commit<<<n_sel, 256, 0, s>>>(b->v, w, (const long long *)d_pred, w.kind);
    } else {
        k_cloud_commit<<<b->cloud_count, 256, 0, s>>>(b->v, w, (const long long *)d_pred, nullptr);
    }
    CHECK_LAUNCH();
    b->launches += 1;
    return NIRRT_OK;
}

// ---- math.sin / math.cos of the reference's runtime (glibc 2.39 kernels restated, glibc_trig.cuh), element-wise
__global__ void k_sincos(const double *x, long long n, double *s, double *c) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) { s[i] = glibc_sin(x[i]); c[i] = glibc_cos(x[i]); }
}
__global__ void k_atan2(const double *y, const double *x, long long n, double *out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) out[i] = glibc_atan2(y[i], x[i]);
}
extern "C" int nirrt_atan2_sync(const double *y, const double *x, int64_t n, double *out, void *stream) {
    if (n < 0 || (n > 0 && (!x || !y || !out))) return fail(NIRRT_ERR_INVALID, "nirrt_atan2_sync: bad argument");
    if (n == 0) return NIRRT_OK;
    cudaStream_t s = (cudaStream_t)stream;
    TempBufs t;
    double *dy, *dx, *dout;
    TRY(t.up(y, (size_t)n, s, &dy)); TRY(t.up(x, (size_t)n, s, &dx)); TRY(t.make<double>((size_t)n, &dout));
    k_atan2<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(dy, dx, n, dout);
    CHECK_LAUNCH();
    CUDA_TRY(cudaMemcpyAsync(out, dout, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return NIRRT_OK;
}
extern "C" int nirrt_sincos_sync(const double *x, int64_t n, double *out_sin, double *out_cos, void *stream) {
    if (n < 0 || (n > 0 && (!x || !out_sin || !out_cos))) return fail(NIRRT_ERR_INVALID, "nirrt_sincos_sync: bad argument");
    if (n == 0) return NIRRT_OK;
    cudaStream_t s = (cudaStream_t)stream;
    TempBufs t;
    double *dx, *ds, *dc;
    TRY(t.up(x, (size_t)n, s, &dx)); TRY(t.make<double>((size_t)n, &ds)); TRY(t.make<double>((size_t)n, &dc));
    k_sincos<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(dx, n, ds, dc);
    CHECK_LAUNCH();
    CUDA_TRY(cudaMemcpyAsync(out_sin, ds, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(out_cos, dc, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return NIRRT_OK;
}

extern "C" int nirrt_batch_counters(nirrt_batch *b, int64_t *kernel_launches, int64_t *reserved) {
    if (!b) return fail(NIRRT_ERR_INVALID, "null batch");
    if (kernel_launches) *kernel_launches = b->launches;
    if (reserved) *reserved = scan_bytes_per_vertex(b->v);   // bytes per vertex one scan pass reads
    return NIRRT_OK;
}

__global__ void k_reset_cand(View v) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < v.E) { EnvCtl *c = v.ctl + e; c->cand_cnt = 0; c->hdr0.n = IT.hdr1.n = c->n; }
}

extern "C" int nirrt_batch_time_scan_sync(nirrt_batch *b, int which, int reps, float *ms, int64_t *bytes, void *stream) {
    if (!b || reps < 1 || !ms || which < 0 || which > 1) return fail(NIRRT_ERR_INVALID, "nirrt_batch_time_scan_sync: bad argument");
    View &v = b->v;
    cudaStream_t s = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(b->device));
    TRY(fetch_ctl(b, s));
    int64_t total = 0;
    for (int e = 0; e < v.E; e++) total += (int64_t)b->h_ctl[e].n * scan_bytes_per_vertex(v);
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
    float acc = 0.f;
    for (int r = 0; r < reps; r++) {
        k_reset_cand<<<(v.E + 127) / 128, 128, 0, s>>>(v);
        CUDA_TRY(cudaEventRecord(e0, s));
        launch_scan<true>(v, which, v.E, s);
        CUDA_TRY(cudaEventRecord(e1, s));
        CUDA_TRY(cudaEventSynchronize(e1));
        float t = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&t, e0, e1));
        acc += t;
    }
    k_reset_cand<<<(v.E + 127) / 128, 128, 0, s>>>(v);
    CHECK_LAUNCH();
    CUDA_TRY(cudaStreamSynchronize(s));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *ms = acc / reps;
    if (bytes) *bytes = total;
    return NIRRT_OK;
}

// ------------------------------------------------------------------------------------------------
// Farthest-point down-sampling of a float64 point set (guidance-cloud generation, SURVEY row f1:
// open3d PointCloud.farthest_point_down_sample as called by point_cloud_mask_utils_3d.py:49-52,
// 196-199): start at `start`, running minimum of the squared distance (dx*dx + dy*dy) + dz*dz in
// f64, next = first argmax.  One CTA; distances live in registers, the selected point's
// coordinates are re-read from L2 each step.
__global__ void __launch_bounds__(kFpsThreads) k_fps_f64(const double *pts, int n, int npoint, int start, long long *out) {
    fps_f64_body(pts, n, npoint, start, [&](int it, int idx) { out[it] = idx; });
}

extern "C" int nirrt_fps_f64_sync(const double *points, int64_t n, int npoint, int start, int64_t *out_idx, void *stream) {
    if (!points || !out_idx || n < 1 || npoint < 1 || npoint > n || start < 0 || start >= n)
        return fail(NIRRT_ERR_INVALID, "nirrt_fps_f64_sync: bad argument");
    if (n > (int64_t)kFpsThreads * kFpsPPT) return fail(NIRRT_ERR_CAPACITY, "nirrt_fps_f64_sync: at most 16384 points");
    if (nirrt_device_count() <= 0) return fail(NIRRT_ERR_NO_DEVICE, "no sm_100 device");
    cudaStream_t s = (cudaStream_t)stream;
    TempBufs t;
    double *dp; long long *dout;
    TRY(t.up(points, 3 * (size_t)n, s, &dp)); TRY(t.make<long long>((size_t)npoint, &dout));
    k_fps_f64<<<1, kFpsThreads, 0, s>>>(dp, (int)n, npoint, start, dout);
    CHECK_LAUNCH();
    CUDA_TRY(cudaMemcpyAsync(out_idx, dout, sizeof(long long) * npoint, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return NIRRT_OK;
}
