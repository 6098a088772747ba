// pointnet2.cu -- PointNet++ (MSG semantic segmentation) guidance-state inference for batches of
// point clouds on sm_100a, and the C ABI over it (include/nirrt_pointnet2.h).
//
// Stages of one forward (reference: pointnet_pointnet2/models/pointnet2.py:24-42):
//   k_prep         pc_normalize + [xyz, start, goal, free] feature build   (pointnet2_wrapper.py:46-59)
//   per SA level:  k_fps          farthest point sampling, 1 CTA / cloud    (pointnet2_utils.py:65-86)
//                  k_ball_query   both radii: slab-pruned candidates, 1 thread / centroid (levels 1-3 of large batches),
//                                 k_ball_query_bf / _warp exhaustive                       (pointnet2_utils.py:89-109)
//     sa1, sa2:    safused::k_sa_fused   gather + 3 x (conv1x1 + BN + ReLU) + max over K in one tcgen05 kernel
//     sa3:         safused::k_sa_fused (two-layer mode) + umma::k_gemm with the pooling epilogue
//     sa4:         k_group (gather -> fp16 operand, pointnet2_utils.py:246-253) + 3 x umma::k_gemm
//   fp4..fp2:      k_knn_pruned / k_interp<1>  3-NN search (geometry stream), k_interp<2> interpolation + skip concat
//                  (:298-311), umma::k_gemm conv1d + BN + ReLU chain
//   fp1 + head:    safused::k_fp1_fused   interpolation + 3 layers + conv1/bn1/relu + conv2, log_softmax, argmax,
//                  softmax[:,1] in one kernel (layer by layer with NIRRT_PN2_FP_FUSED=0: umma::k_gemm + k_head)
//
// Everything between the input cloud and the per-point outputs stays in HBM/L2; activations and
// weights are fp16 (K-major rows), accumulation fp32 in TMEM, coordinates enter the first
// convolution of each SA level as an fp16 hi+lo pair so that no position information is lost.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/nirrt_b200.h"
#include "../../include/nirrt_pointnet2.h"
#include "errors.h"
#include "devmem.h"
#include "umma_gemm.cuh"
#include "sa_fused.cuh"

static int pfail(int code, const std::string &msg) { return nirrt_set_error(code, msg); }
#define PCUDA(expr)                                                                                   \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess)                                                                        \
            return pfail(NIRRT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));         \
    } while (0)
#define PTRY(expr) do { int _r = (expr); if (_r) return _r; } while (0)

// ------------------------------------------------------------------------------------------------
// k_prep: pc_normalize (pointnet2_utils.py:13-18) in numpy's float32 operation order -- column sums
// accumulate row by row (np.mean over axis 0), norm = sqrt((x*x + y*y) + z*z), divide by the max --
// and the 6-channel network input [x, y, z, start, goal, free] (pointnet2_wrapper.py:51-59).
__global__ void __launch_bounds__(256) k_prep(const float *pc, int dim, const float *sm, const float *gm, int N,
                                              float *xyz0, float *in6) {
    const int b = blockIdx.x;
    // the cloud is staged in shared memory once: the column sums are 3 serial chains of N dependent additions (numpy's
    // order), which must not wait for a global load per term
    extern __shared__ float s_pc[];           // [N][dim]
    {
        const float *G = pc + (size_t)b * N * dim;
        for (int i = threadIdx.x; i < N * dim; i += blockDim.x) s_pc[i] = G[i];
    }
    const float *P = s_pc;
    __shared__ float s_c[3];
    __shared__ float s_red[8];
    __syncthreads();
    if (threadIdx.x < 3) {
        float s = 0.f;
        if ((int)threadIdx.x < dim) {
#pragma unroll 8
            for (int i = 0; i < N; i++) s = __fadd_rn(s, P[(size_t)i * dim + threadIdx.x]);
        }
        s_c[threadIdx.x] = __fdiv_rn(s, (float)N);
    }
    __syncthreads();
    const float cx = s_c[0], cy = s_c[1], cz = s_c[2];
    float mx = 0.f;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const float x = __fsub_rn(P[(size_t)i * dim], cx), y = __fsub_rn(P[(size_t)i * dim + 1], cy);
        const float z = dim > 2 ? __fsub_rn(P[(size_t)i * dim + 2], cz) : 0.f;
        const float n2 = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
        mx = fmaxf(mx, __fsqrt_rn(n2));
    }
    for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = s_red[0];
    for (int w = 1; w < 8; w++) mx = fmaxf(mx, s_red[w]);
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const float x = __fdiv_rn(__fsub_rn(P[(size_t)i * dim], cx), mx);
        const float y = __fdiv_rn(__fsub_rn(P[(size_t)i * dim + 1], cy), mx);
        const float z = dim > 2 ? __fdiv_rn(__fsub_rn(P[(size_t)i * dim + 2], cz), mx) : __fdiv_rn(0.f, mx);
        const size_t o = (size_t)b * N + i;
        xyz0[o * 3] = x; xyz0[o * 3 + 1] = y; xyz0[o * 3 + 2] = z;
        const float s = sm[o], g = gm[o];
        float *f = in6 + o * 6;
        f[0] = x; f[1] = y; f[2] = z; f[3] = s; f[4] = g;
        f[5] = (__fadd_rn(s, g) != 0.f) ? 0.f : 1.f;      // 1 - (start+goal).astype(bool)
    }
}

// ------------------------------------------------------------------------------------------------
// k_fps: farthest_point_sample (pointnet2_utils.py:65-86).  One CTA per cloud, the cloud's running
// min-distance lives in registers (PPT points per thread), one barrier per selected point.
// dist = (dx*dx + dy*dy) + dz*dz in fp32 without contraction; argmax ties -> lowest index.
// A selection step is one dependent chain (centroid -> distances -> arg-max -> next centroid), so what counts is its
// length, not the instruction count: the per-thread maximum is an order-free max over the PPT values (the compiler
// builds a tree), the arg-max is a max-reduction on the bit patterns (distances are >= 0) followed by a min-reduction
// on the index, both REDUX, and every thread combines the per-warp results itself after the one barrier.
// Packed fp32 pairs (sm_100 FADD2 / FMUL2: two IEEE round-to-nearest operations per instruction, the same results as two
// scalar instructions at half the fma-pipe slots).  ptxas contracts a packed multiply feeding a packed add into FFMA2 even
// for .rn operands and with --fmad=false (it even folds fma(a, b, -0) back into a multiply first), so wherever the reference
// rounds the product the sum is formed with SCALAR adds on the unpacked halves: the SASS of those kernels must show
// FMUL2 / FADD2 / FADD and no FFMA2 (checked by tests/test_build_sass.py).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

template <int T, int PPT>
__global__ void __launch_bounds__(T) k_fps(const float *xyz, int N, int npoint, const int *start, int level,
                                           int *fidx, float *new_xyz, int it0, int it1, float *carry_dist, int *carry_far) {
    extern __shared__ float s_xyz[];          // [N][3]
    constexpr int NW = T / 32;
    __shared__ __align__(16) unsigned s_d[2][NW];
    __shared__ __align__(16) int s_i[2][NW];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *X = xyz + (size_t)b * N * 3;
    for (int i = tid; i < N * 3; i += T) s_xyz[i] = X[i];
    __syncthreads();
    static_assert(PPT % 2 == 0, "points are processed in packed pairs");
    f32x2 px[PPT / 2], py[PPT / 2], pz[PPT / 2];
    float dist[PPT];
    // [it0, it1) of the npoint selections: a later chunk picks the running distances up where the previous launch left
    // them (the level-1 selections are issued in chunks so that grouping and the first MLPs start on the early centroids)
    float *cd = carry_dist + (size_t)b * N;
#pragma unroll
    for (int j = 0; j < PPT; j += 2) {
        float c[2][3];
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int i = tid + (j + k) * T;
            const bool in = i < N;
            c[k][0] = in ? s_xyz[i * 3] : 0.f; c[k][1] = in ? s_xyz[i * 3 + 1] : 0.f; c[k][2] = in ? s_xyz[i * 3 + 2] : 0.f;
            dist[j + k] = in ? (it0 > 0 ? cd[i] : 1e10f) : 0.f;
        }
        px[j / 2] = pack2(c[0][0], c[1][0]); py[j / 2] = pack2(c[0][1], c[1][1]); pz[j / 2] = pack2(c[0][2], c[1][2]);
    }
    int far = it0 > 0 ? carry_far[b] : start[b * 4 + level];
    for (int it = it0; it < it1; it++) {
        const float cx = s_xyz[far * 3], cy = s_xyz[far * 3 + 1], cz = s_xyz[far * 3 + 2];
        if (tid == 0) {
            fidx[(size_t)b * npoint + it] = far;
            float *o = new_xyz + ((size_t)b * npoint + it) * 3;
            o[0] = cx; o[1] = cy; o[2] = cz;
        }
        unsigned u[PPT], m = 0u;
        const f32x2 cx2 = pack2(cx, cx), cy2 = pack2(cy, cy), cz2 = pack2(cz, cz);
#pragma unroll
        for (int j = 0; j < PPT; j += 2) {
            const f32x2 dx = sub2(px[j / 2], cx2), dy = sub2(py[j / 2], cy2), dz = sub2(pz[j / 2], cz2);
            float x0, x1, y0, y1, z0, z1;
            unpack2(mul2(dx, dx), x0, x1); unpack2(mul2(dy, dy), y0, y1); unpack2(mul2(dz, dz), z0, z1);
            const float d0 = __fadd_rn(__fadd_rn(x0, y0), z0), d1 = __fadd_rn(__fadd_rn(x1, y1), z1);
            dist[j] = fminf(dist[j], d0); dist[j + 1] = fminf(dist[j + 1], d1);      // no NaNs: the same as `if (d < dist)`
            u[j] = __float_as_uint(dist[j]); u[j + 1] = __float_as_uint(dist[j + 1]);          // out-of-range slots hold 0
            m = max(m, max(u[j], u[j + 1]));
        }
        const unsigned wm = __reduce_max_sync(0xffffffffu, m);
        // lowest index of this thread at the warp's maximum (slot j <-> index tid + j*T, ascending; an out-of-range slot
        // can only match when every distance is 0, and then only after all of the thread's in-range slots)
        unsigned eq = 0u;
#pragma unroll
        for (int j = 0; j < PPT; j++) eq |= (u[j] == wm ? 1u : 0u) << j;
        int bi = eq ? tid + (__ffs(eq) - 1) * T : 0x7fffffff;
        if (bi >= N) bi = 0x7fffffff;
        const int wi = __reduce_min_sync(0xffffffffu, bi);
        if (NW == 1) { far = wi; continue; }
        const int slot = it & 1;
        if (lane == 0) { s_d[slot][warp] = wm; s_i[slot][warp] = wi; }
        __syncthreads();
        unsigned gm = 0u;
#pragma unroll
        for (int w = 0; w < NW; w++) gm = max(gm, s_d[slot][w]);
        int gi = 0x7fffffff;
#pragma unroll
        for (int w = 0; w < NW; w++) gi = min(gi, s_d[slot][w] == gm ? s_i[slot][w] : 0x7fffffff);
        far = gi;
    }
    if (it1 < npoint) {
#pragma unroll
        for (int j = 0; j < PPT; j++) {
            const int i = tid + j * T;
            if (i < N) cd[i] = dist[j];
        }
        if (tid == 0) carry_far[b] = far;
    }
}

// k_fps_bucket: the same selection with spatial pruning.  The cloud is counting-sorted by 9-bit Morton cell and cut into
// buckets of 32 consecutive points (one point per lane; 16 buckets per warp, interleaved over the warps); every bucket keeps
// its bounding box, the largest running distance of its points and the lowest index at that distance.  A selection step
// evaluates, per bucket, the distance formula on the GAPS between the new centroid and the box: every float32 operation of the
// formula is monotone, so that value is a lower bound of the computed distance of every point in the bucket, and a bucket whose
// bound is not below its largest running distance cannot change -- it is skipped.  Late in the selection (most of the steps)
// a new centroid only touches the few buckets around it.  Same distances, same arg-max with lowest-index ties: the selected
// indices are those of k_fps whatever the bucket composition (the scatter order inside a cell is arbitrary).
// Measured (256 clouds x 2048 points, B200): SLOWER than k_fps -- 0.76 vs 0.59 ms for the four levels, 0.70 vs 0.53 ms for one
// cloud: the selection step stays one dependent chain, and bound + ballot + per-bucket REDUX pairs lengthen it by more than the
// skipped distance updates shorten it.  Kept as an opt-in (NIRRT_PN2_FPS_BUCKET=1) with its parity test.
__device__ __forceinline__ int morton_cell(float x, float y, float z) {
    const int cx = min(7, max(0, (int)((x + 1.f) * 4.f))), cy = min(7, max(0, (int)((y + 1.f) * 4.f))), cz = min(7, max(0, (int)((z + 1.f) * 4.f)));
    int m = 0;
#pragma unroll
    for (int b = 0; b < 3; b++) m |= (((cx >> b) & 1) << (3 * b)) | (((cy >> b) & 1) << (3 * b + 1)) | (((cz >> b) & 1) << (3 * b + 2));
    return m;
}
template <int NW>
__global__ void __launch_bounds__(32 * NW) k_fps_bucket(const float *xyz, int N, int npoint, const int *start, int level,
                                                        int *fidx, float *new_xyz) {
    constexpr int T = 32 * NW, NBW = 16, CAP = T * NBW;
    extern __shared__ float s_xyz[];          // [N][3] in the caller's order, then perm[CAP] (u16), hist[513]
    unsigned short *s_perm = reinterpret_cast<unsigned short *>(s_xyz + 3 * N);
    int *s_hist = reinterpret_cast<int *>(s_perm + CAP + (CAP & 1));
    __shared__ __align__(16) unsigned s_d[2][NW];
    __shared__ __align__(16) int s_i[2][NW];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *X = xyz + (size_t)b * N * 3;
    for (int i = tid; i < N * 3; i += T) s_xyz[i] = X[i];
    for (int i = tid; i < 513; i += T) s_hist[i] = 0;
    __syncthreads();
    for (int i = tid; i < N; i += T) atomicAdd(&s_hist[morton_cell(s_xyz[i * 3], s_xyz[i * 3 + 1], s_xyz[i * 3 + 2])], 1);
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int c = 0; c < 512; c++) { const int t = s_hist[c]; s_hist[c] = run; run += t; }
    }
    __syncthreads();
    for (int i = tid; i < N; i += T)
        s_perm[atomicAdd(&s_hist[morton_cell(s_xyz[i * 3], s_xyz[i * 3 + 1], s_xyz[i * 3 + 2])], 1)] = (unsigned short)i;
    __syncthreads();
    float px[NBW], py[NBW], pz[NBW], dist[NBW];
    int idx[NBW];
    float lox = 0.f, loy = 0.f, loz = 0.f, hix = 0.f, hiy = 0.f, hiz = 0.f;      // lane j < NBW: bounding box of bucket j of this warp
    unsigned bmax = 0u;                       // ... its largest running distance (bit pattern)
    int bidx = 0x7fffffff;                    // ... and the lowest index at that distance
#pragma unroll
    for (int j = 0; j < NBW; j++) {
        const int pos = (j * NW + warp) * 32 + lane;
        const bool in = pos < N;
        const int i = in ? (int)s_perm[pos] : 0;
        px[j] = in ? s_xyz[i * 3] : 0.f; py[j] = in ? s_xyz[i * 3 + 1] : 0.f; pz[j] = in ? s_xyz[i * 3 + 2] : 0.f;
        dist[j] = in ? 1e10f : 0.f;
        idx[j] = in ? i : 0x7fffffff;
        float a0 = in ? px[j] : INFINITY, a1 = in ? py[j] : INFINITY, a2 = in ? pz[j] : INFINITY;
        float b0 = in ? px[j] : -INFINITY, b1 = in ? py[j] : -INFINITY, b2 = in ? pz[j] : -INFINITY;
        int mi = idx[j];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            a0 = fminf(a0, __shfl_xor_sync(0xffffffffu, a0, off)); a1 = fminf(a1, __shfl_xor_sync(0xffffffffu, a1, off));
            a2 = fminf(a2, __shfl_xor_sync(0xffffffffu, a2, off));
            b0 = fmaxf(b0, __shfl_xor_sync(0xffffffffu, b0, off)); b1 = fmaxf(b1, __shfl_xor_sync(0xffffffffu, b1, off));
            b2 = fmaxf(b2, __shfl_xor_sync(0xffffffffu, b2, off));
            mi = min(mi, __shfl_xor_sync(0xffffffffu, mi, off));
        }
        if (lane == j) {
            lox = a0; loy = a1; loz = a2; hix = b0; hiy = b1; hiz = b2;
            bmax = mi != 0x7fffffff ? __float_as_uint(1e10f) : 0u;
            bidx = mi;
        }
    }
    int far = start[b * 4 + level];
    for (int it = 0; it < npoint; it++) {
        const float cx = s_xyz[far * 3], cy = s_xyz[far * 3 + 1], cz = s_xyz[far * 3 + 2];
        if (tid == 0) {
            fidx[(size_t)b * npoint + it] = far;
            float *o = new_xyz + ((size_t)b * npoint + it) * 3;
            o[0] = cx; o[1] = cy; o[2] = cz;
        }
        // lower bound of the bucket's distances: the formula on the gaps to the box (monotone operations; an empty bucket has
        // an infinite bound)
        const float gx = fmaxf(fmaxf(__fsub_rn(lox, cx), __fsub_rn(cx, hix)), 0.f);
        const float gy = fmaxf(fmaxf(__fsub_rn(loy, cy), __fsub_rn(cy, hiy)), 0.f);
        const float gz = fmaxf(fmaxf(__fsub_rn(loz, cz), __fsub_rn(cz, hiz)), 0.f);
        const float lb = __fadd_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)), __fmul_rn(gz, gz));
        const unsigned act = __ballot_sync(0xffffffffu, lane < NBW && lb < __uint_as_float(bmax));
#pragma unroll
        for (int j = 0; j < NBW; j++) {
            if ((act >> j) & 1u) {            // warp-uniform
                const float dx = __fsub_rn(px[j], cx), dy = __fsub_rn(py[j], cy), dz = __fsub_rn(pz[j], cz);
                const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                if (d < dist[j]) dist[j] = d;
                const unsigned u = __float_as_uint(dist[j]);          // padding slots hold 0 and index INT_MAX
                const unsigned bm = __reduce_max_sync(0xffffffffu, u);
                const int bi = __reduce_min_sync(0xffffffffu, u == bm ? idx[j] : 0x7fffffff);
                if (lane == j) { bmax = bm; bidx = bi; }
            }
        }
        const unsigned wm = __reduce_max_sync(0xffffffffu, lane < NBW ? bmax : 0u);
        const int wi = __reduce_min_sync(0xffffffffu, (lane < NBW && bmax == wm) ? bidx : 0x7fffffff);
        if (NW == 1) { far = wi; continue; }
        const int slot = it & 1;
        if (lane == 0) { s_d[slot][warp] = wm; s_i[slot][warp] = wi; }
        __syncthreads();
        unsigned gm = 0u;
#pragma unroll
        for (int w = 0; w < NW; w++) gm = max(gm, s_d[slot][w]);
        int gi = 0x7fffffff;
#pragma unroll
        for (int w = 0; w < NW; w++) gi = min(gi, s_d[slot][w] == gm ? s_i[slot][w] : 0x7fffffff);
        far = gi;
    }
}

// ------------------------------------------------------------------------------------------------
// k_ball_query: query_ball_point (pointnet2_utils.py:89-109) for both radii of an SA level.
// One warp per centroid scans the cloud in index order (32 points per step, ballot + prefix
// popcount), keeps the first K members of each ball and pads with the first member.  Distances use
// the reference's expansion -2*a.b + |a|^2 + |b|^2 (square_distance, :39-41).
// small batches: one WARP per centroid (32 points per step, ballot + prefix popcount) -- more parallelism
// when there are few clouds; large batches use the one-thread-per-centroid kernel below
__global__ void __launch_bounds__(256) k_ball_query_warp(const float *xyz, int N, const float *new_xyz, int S,
                                                    float r0sq, int K0, float r1sq, int K1, int *g0, int *g1, int s0, int s1) {
    extern __shared__ float s_pts[];          // x[N] y[N] z[N] sq[N]
    float *sx = s_pts, *sy = sx + N, *sz = sy + N, *sq = sz + N;
    const int b = blockIdx.y;
    const float *X = xyz + (size_t)b * N * 3;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const float x = X[i * 3], y = X[i * 3 + 1], z = X[i * 3 + 2];
        sx[i] = x; sy[i] = y; sz[i] = z;
        sq[i] = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int s = s0 + blockIdx.x * 8 + (threadIdx.x >> 5);      // centroids [s0, s1) of every cloud
    if (s >= s1) return;
    const float *c = new_xyz + ((size_t)b * S + s) * 3;
    const float cx = c[0], cy = c[1], cz = c[2];
    const float cs = __fadd_rn(__fadd_rn(__fmul_rn(cx, cx), __fmul_rn(cy, cy)), __fmul_rn(cz, cz));
    int *o0 = g0 + ((size_t)b * S + s) * K0, *o1 = g1 + ((size_t)b * S + s) * K1;
    int cnt0 = 0, cnt1 = 0;
    const unsigned lt = (1u << lane) - 1u;
    for (int base = 0; base < N && (cnt0 < K0 || cnt1 < K1); base += 32) {
        const int i = base + lane;
        bool m0 = false, m1 = false;
        if (i < N) {
            const float dot = __fmaf_rn(cz, sz[i], __fmaf_rn(cy, sy[i], __fmul_rn(cx, sx[i])));
            float d = __fmul_rn(-2.f, dot);
            d = __fadd_rn(d, cs);
            d = __fadd_rn(d, sq[i]);
            m0 = !(d > r0sq); m1 = !(d > r1sq);
        }
        const unsigned b0 = __ballot_sync(0xffffffffu, m0), b1 = __ballot_sync(0xffffffffu, m1);
        if (m0) { const int pos = cnt0 + __popc(b0 & lt); if (pos < K0) o0[pos] = i; }
        if (m1) { const int pos = cnt1 + __popc(b1 & lt); if (pos < K1) o1[pos] = i; }
        cnt0 += __popc(b0); cnt1 += __popc(b1);
    }
    __syncwarp();
    // pad with the first member (group_first); an empty ball cannot occur: the centroid is a member
    if (cnt0 < K0) { const int f = cnt0 > 0 ? o0[0] : 0; for (int p = cnt0 + lane; p < K0; p += 32) o0[p] = f; }
    if (cnt1 < K1) { const int f = cnt1 > 0 ? o1[0] : 0; for (int p = cnt1 + lane; p < K1; p += 32) o1[p] = f; }
}

// Pruned version for radii well below the cloud's extent (levels 1-3): the coordinates are normalised to the unit ball, each
// axis is cut into kSlabs slabs, and bitmap[axis][slab] marks the points whose coordinate lies within `reach` (>= the large
// radius plus the rounding slack of the expansion formula) of the slab.  A centroid in slabs (ix, iy, iz) can only have members
// in bitmap[0][ix] & bitmap[1][iy] & bitmap[2][iz]; the thread walks the set bits in index order and applies the SAME float32
// test to them, so the groups are those of the exhaustive scan (1-15 % of the points are tested instead of all).
constexpr int kSlabs = 8;
constexpr int kSlabWordsMax = 128;        // 4096 points
__device__ __forceinline__ float slab_lo(int s) { return __fmaf_rn((float)s, 2.f / kSlabs, -1.f); }
__global__ void __launch_bounds__(1024) k_ball_query(const float *xyz, int N, const float *new_xyz, int S,
                                                    float r0sq, int K0, float r1sq, int K1, int *g0, int *g1, int s0, int s1, float reach) {
    extern __shared__ float4 s_p4[];          // [N] (x, y, z, |p|^2), then the bitmaps [3][kSlabs][W]
    const int Wn = (N + 31) >> 5, W = Wn + 1;   // row stride Wn + 1: the rows of different slabs start in different banks
    unsigned *bm = reinterpret_cast<unsigned *>(s_p4 + N);
    const int b = blockIdx.y;
    const float *X = xyz + (size_t)b * N * 3;
    for (int i = threadIdx.x; i < 3 * kSlabs * W; i += blockDim.x) bm[i] = 0u;
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const float v[3] = {X[i * 3], X[i * 3 + 1], X[i * 3 + 2]};
        s_p4[i] = make_float4(v[0], v[1], v[2], __fadd_rn(__fadd_rn(__fmul_rn(v[0], v[0]), __fmul_rn(v[1], v[1])), __fmul_rn(v[2], v[2])));
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int sl = 0; sl < kSlabs; sl++) {      // the end slabs are unbounded outwards
                const bool lo_ok = sl == 0 || v[a] >= slab_lo(sl) - reach;
                const bool hi_ok = sl == kSlabs - 1 || v[a] <= slab_lo(sl + 1) + reach;
                if (lo_ok && hi_ok) atomicOr(&bm[(a * kSlabs + sl) * W + (i >> 5)], 1u << (i & 31));
            }
    }
    __syncthreads();
    const int s = s0 + blockIdx.x * blockDim.x + threadIdx.x;      // centroids [s0, s1) of every cloud
    if (s >= s1) return;
    const float *c = new_xyz + ((size_t)b * S + s) * 3;
    const float cx = c[0], cy = c[1], cz = c[2];
    const float cs = __fadd_rn(__fadd_rn(__fmul_rn(cx, cx), __fmul_rn(cy, cy)), __fmul_rn(cz, cz));
    int *o0 = g0 + ((size_t)b * S + s) * K0, *o1 = g1 + ((size_t)b * S + s) * K1;
    int cnt0 = 0, cnt1 = 0, f0 = 0, f1 = 0;
    int ix = 0, iy = 0, iz = 0;               // the slab with slab_lo(s) <= c < slab_lo(s + 1) (clamped at the ends)
#pragma unroll
    for (int sl = 1; sl < kSlabs; sl++) {
        const float lo = slab_lo(sl);
        ix += cx >= lo ? 1 : 0; iy += cy >= lo ? 1 : 0; iz += cz >= lo ? 1 : 0;
    }
    const unsigned *bx = bm + (0 * kSlabs + ix) * W, *by = bm + (1 * kSlabs + iy) * W, *bz = bm + (2 * kSlabs + iz) * W;
    bool full = false;
    for (int w = 0; w < Wn && !full; w++) {
        unsigned m = bx[w] & by[w] & bz[w];
        while (m && !full) {
            const int i = (w << 5) + __ffs(m) - 1;
            m &= m - 1;
            const float4 p = s_p4[i];
            const float dot = __fmaf_rn(cz, p.z, __fmaf_rn(cy, p.y, __fmul_rn(cx, p.x)));
            float d = __fmul_rn(-2.f, dot);
            d = __fadd_rn(d, cs);
            d = __fadd_rn(d, p.w);
            if (!(d > r1sq)) {                       // r0 < r1: members of the small ball are members of the large one
                if (cnt1 < K1) { if (cnt1 == 0) f1 = i; o1[cnt1++] = i; }
                if (!(d > r0sq) && cnt0 < K0) { if (cnt0 == 0) f0 = i; o0[cnt0++] = i; }
                full = cnt1 >= K1 && cnt0 >= K0;
            }
        }
    }
    // pad with the first member (group_first); an empty ball cannot occur: the centroid is a member
    for (int p = cnt0; p < K0; p++) o0[p] = f0;
    for (int p = cnt1; p < K1; p++) o1[p] = f1;
}

// exhaustive version (large radii)
__global__ void __launch_bounds__(256) k_ball_query_bf(const float *xyz, int N, const float *new_xyz, int S,
                                                       float r0sq, int K0, float r1sq, int K1, int *g0, int *g1, int s0, int s1) {
    // points in pairs for the packed fp32 pipe: s_pp[2k] = (x_2k, x_2k+1, y_2k, y_2k+1), s_pp[2k + 1] = (z_2k, z_2k+1, |p|^2_2k,
    // |p|^2_2k+1): two broadcast LDS.128 per two points; an odd cloud is padded with a point far outside every ball
    extern __shared__ float4 s_pp[];
    const int b = blockIdx.y;
    const float *X = xyz + (size_t)b * N * 3;
    for (int k = threadIdx.x; 2 * k < N; k += blockDim.x) {
        float c[2][4];
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int i = 2 * k + u;
            const bool in = i < N;
            const float x = in ? X[i * 3] : 1e10f, y = in ? X[i * 3 + 1] : 1e10f, z = in ? X[i * 3 + 2] : 1e10f;
            c[u][0] = x; c[u][1] = y; c[u][2] = z;
            c[u][3] = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
        }
        s_pp[2 * k] = make_float4(c[0][0], c[1][0], c[0][1], c[1][1]);
        s_pp[2 * k + 1] = make_float4(c[0][2], c[1][2], c[0][3], c[1][3]);
    }
    __syncthreads();
    const int s = s0 + blockIdx.x * blockDim.x + threadIdx.x;      // centroids [s0, s1) of every cloud
    if (s >= s1) return;
    const float *c = new_xyz + ((size_t)b * S + s) * 3;
    const float cx = c[0], cy = c[1], cz = c[2];
    const float cs = __fadd_rn(__fadd_rn(__fmul_rn(cx, cx), __fmul_rn(cy, cy)), __fmul_rn(cz, cz));
    int *o0 = g0 + ((size_t)b * S + s) * K0, *o1 = g1 + ((size_t)b * S + s) * K1;
    int cnt0 = 0, cnt1 = 0, f0 = 0, f1 = 0;
    // the thread walks the cloud in index order, two points per step (FMUL2 / FFMA2 / FADD2: the reference's
    // -2 * (a . b) + |a|^2 + |b|^2 with the same roundings), and keeps the first K members of each ball
    const f32x2 cx2 = pack2(cx, cx), cy2 = pack2(cy, cy), cz2 = pack2(cz, cz), cs2 = pack2(cs, cs), m2 = pack2(-2.f, -2.f);
    bool full = false;
    for (int i = 0; i < N && !full; i += 2) {
        const float4 pa = s_pp[i], pb = s_pp[i + 1];
        const f32x2 dot = fma2(cz2, pack2(pb.x, pb.y), fma2(cy2, pack2(pa.z, pa.w), mul2(cx2, pack2(pa.x, pa.y))));
        float d[2];
        unpack2(add2(add2(mul2(m2, dot), cs2), pack2(pb.z, pb.w)), d[0], d[1]);      // -2 * dot is exact: a contraction changes nothing
        if (!(d[0] > r1sq) || !(d[1] > r1sq)) {
#pragma unroll
            for (int u = 0; u < 2; u++) {
                if (!(d[u] > r1sq) && i + u < N && !full) {     // r0 < r1: members of the small ball are members of the large one
                    if (cnt1 < K1) { if (cnt1 == 0) f1 = i + u; o1[cnt1++] = i + u; }
                    if (!(d[u] > r0sq) && cnt0 < K0) { if (cnt0 == 0) f0 = i + u; o0[cnt0++] = i + u; }
                    full = cnt1 >= K1 && cnt0 >= K0;
                }
            }
        }
    }
    // pad with the first member (group_first); an empty ball cannot occur: the centroid is a member
    for (int p = cnt0; p < K0; p++) o0[p] = f0;
    for (int p = cnt1; p < K1; p++) o1[p] = f1;
}

// ------------------------------------------------------------------------------------------------
// grouping (pointnet2_utils.py:246-253): rows of the first convolution's operand,
// [features of neighbour, xyz(neighbour) - xyz(centroid)], features first.  Coordinates are
// written as an fp16 hi part plus an fp16 lo (residual) part; the weight matrix repeats the
// corresponding columns, so the tensor cores see coordinates with ~22 significant bits.
__device__ __forceinline__ void split_half(float x, __half &hi, __half &lo) {
    hi = __float2half_rn(x);
    lo = __float2half_rn(__fsub_rn(x, __half2float(hi)));
}

// sa1: in6 f32 [B][N][6] -> rows of 16 halves:
//   [x y z start goal free | rx ry rz | x_lo y_lo z_lo | rx_lo ry_lo rz_lo | 0]
__global__ void __launch_bounds__(256) k_group_sa1(const float *in6, const float *xyz, int N, const float *new_xyz, int S,
                                                   const int *gidx, int K, int B, __half *out) {
    const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)B * S * K;
    if (r >= total) return;
    const size_t bs = r / K;
    const int b = (int)(bs / S);
    const int i = gidx[r];
    const float *f = in6 + ((size_t)b * N + i) * 6;
    const float *c = new_xyz + bs * 3;
    __half h[16];
    __half lo;
    float v[9];
    for (int k = 0; k < 6; k++) v[k] = f[k];
    for (int k = 0; k < 3; k++) v[6 + k] = __fsub_rn(f[k], c[k]);
    for (int k = 0; k < 3; k++) { split_half(v[k], h[k], lo); h[9 + k] = lo; }
    for (int k = 3; k < 6; k++) h[k] = __float2half_rn(v[k]);
    for (int k = 6; k < 9; k++) { split_half(v[k], h[k], lo); h[6 + k] = lo; }
    h[15] = __float2half_rn(0.f);
    uint4 *dst = reinterpret_cast<uint4 *>(out + r * 16);
    dst[0] = *reinterpret_cast<uint4 *>(h);
    dst[1] = *reinterpret_cast<uint4 *>(h + 8);
}

// sa2..sa4: feat fp16 [B][N][C] (C % 8 == 0) -> rows of Kpad halves:
//   [feat(C) | rx ry rz | rx_lo ry_lo rz_lo | 0...]; one thread per 16-byte chunk
// block = (Kpad / 8 chunks) x (rows): thread (x, y) writes chunk x of row blockIdx.x * blockDim.y + y, so no thread divides
// (K and S are powers of two: shifts); a warp still writes 32 consecutive chunks
__global__ void __launch_bounds__(256) k_group(const __half *feat, int C, const float *xyz, int N, const float *new_xyz,
                                               int s_shift, const int *gidx, int k_shift, unsigned rows, int Kpad, __half *out) {
    const int FC = C >> 3;
    const unsigned r = blockIdx.x * blockDim.y + threadIdx.y;
    const int ch = threadIdx.x;
    if (r >= rows) return;
    const unsigned bs = r >> k_shift;
    const int b = (int)(bs >> s_shift);
    const int i = gidx[r];
    uint4 val = make_uint4(0, 0, 0, 0);
    if (ch < FC) {
        val = *reinterpret_cast<const uint4 *>(feat + ((size_t)b * N + i) * C + ch * 8);
    } else if (ch == FC) {
        const float *p = xyz + ((size_t)b * N + i) * 3;
        const float *c = new_xyz + (size_t)bs * 3;
        __half h[8];
        for (int k = 0; k < 3; k++) split_half(__fsub_rn(p[k], c[k]), h[k], h[3 + k]);
        h[6] = h[7] = __float2half_rn(0.f);
        val = *reinterpret_cast<uint4 *>(h);
    }
    *reinterpret_cast<uint4 *>(out + (size_t)r * Kpad + ch * 8) = val;
}

// ------------------------------------------------------------------------------------------------
// k_interp: PointNetFeaturePropagation's interpolation (pointnet2_utils.py:298-311): the three
// nearest of the S coarse points by expansion distance, weights 1/(d+1e-8) normalised, weighted sum
// of their features, concatenated after the skip features: row = [points1(C1) | interpolated(C2)].
// One warp per fine point.
struct Top3 { float d[3]; int i[3]; };
__device__ __forceinline__ void top3_insert(Top3 &t, float d, int i) {
    if (d < t.d[2]) {
        if (d < t.d[1]) {
            t.d[2] = t.d[1]; t.i[2] = t.i[1];
            if (d < t.d[0]) { t.d[1] = t.d[0]; t.i[1] = t.i[0]; t.d[0] = d; t.i[0] = i; }
            else { t.d[1] = d; t.i[1] = i; }
        } else { t.d[2] = d; t.i[2] = i; }
    }
}
// small batches: one warp per fine point (lane-strided scan + 3 warp arg-min rounds)
// MODE 0: 3-NN search + interpolation; MODE 1: 3-NN search only (indices / weights -> knn_i / knn_w, runs
// on the geometry stream, it needs coordinates only); MODE 2: interpolation from stored indices / weights.
template <int MODE>
__global__ void __launch_bounds__(256) k_interp_warp(const float *xyz1, int N, const float *xyz2, int S, const __half *feat1, int C1,
                                                const __half *feat2, int C2, int B, __half *out, int4 *knn_i, float4 *knn_w) {
    const int lane = threadIdx.x & 31;
    const size_t p = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (p >= (size_t)B * N) return;
    const int b = (int)(p / N);
    float w0, w1, w2; int ni[3];
    if (MODE != 2) {
    const float *a = xyz1 + p * 3;
    const float ax = a[0], ay = a[1], az = a[2];
    const float as = __fadd_rn(__fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay)), __fmul_rn(az, az));
    const float *Q = xyz2 + (size_t)b * S * 3;
    Top3 t;
    t.d[0] = t.d[1] = t.d[2] = INFINITY; t.i[0] = t.i[1] = t.i[2] = 0x7fffffff;
    for (int s = lane; s < S; s += 32) {
        const float qx = Q[s * 3], qy = Q[s * 3 + 1], qz = Q[s * 3 + 2];
        const float qs = __fadd_rn(__fadd_rn(__fmul_rn(qx, qx), __fmul_rn(qy, qy)), __fmul_rn(qz, qz));
        const float dot = __fmaf_rn(az, qz, __fmaf_rn(ay, qy, __fmul_rn(ax, qx)));
        float d = __fmul_rn(-2.f, dot);
        d = __fadd_rn(d, as);
        d = __fadd_rn(d, qs);
        top3_insert(t, d, s);
    }
    float nd[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        float bd = t.d[0]; int bi = t.i[0];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float od = __shfl_xor_sync(0xffffffffu, bd, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
        }
        nd[r] = bd; ni[r] = bi;
        if (t.i[0] == bi) { t.d[0] = t.d[1]; t.i[0] = t.i[1]; t.d[1] = t.d[2]; t.i[1] = t.i[2]; t.d[2] = INFINITY; t.i[2] = 0x7fffffff; }
    }
    const float r0 = __fdiv_rn(1.f, __fadd_rn(nd[0], 1e-8f)), r1 = __fdiv_rn(1.f, __fadd_rn(nd[1], 1e-8f)),
                r2 = __fdiv_rn(1.f, __fadd_rn(nd[2], 1e-8f));
    const float norm = __fadd_rn(__fadd_rn(r0, r1), r2);
    w0 = __fdiv_rn(r0, norm); w1 = __fdiv_rn(r1, norm); w2 = __fdiv_rn(r2, norm);
    if (MODE == 1) {
        if (lane == 0) { knn_i[p] = make_int4(ni[0], ni[1], ni[2], 0); knn_w[p] = make_float4(w0, w1, w2, 0.f); }
        return;
    }
    } else {
        const int4 ki = knn_i[p]; const float4 kw = knn_w[p];
        ni[0] = ki.x; ni[1] = ki.y; ni[2] = ki.z; w0 = kw.x; w1 = kw.y; w2 = kw.z;
    }
    __half *o = out + p * (size_t)(C1 + C2);
    if (C1 > 0) {
        const uint4 *src = reinterpret_cast<const uint4 *>(feat1 + p * (size_t)C1);
        uint4 *dst = reinterpret_cast<uint4 *>(o);
        for (int k = lane; k < (C1 >> 3); k += 32) dst[k] = src[k];
    }
    const __half2 *f0 = reinterpret_cast<const __half2 *>(feat2 + ((size_t)b * S + ni[0]) * C2);
    const __half2 *f1 = reinterpret_cast<const __half2 *>(feat2 + ((size_t)b * S + ni[1]) * C2);
    const __half2 *f2 = reinterpret_cast<const __half2 *>(feat2 + ((size_t)b * S + ni[2]) * C2);
    __half2 *oi = reinterpret_cast<__half2 *>(o + C1);
    for (int k = lane; k < (C2 >> 1); k += 32) {
        const float2 a0 = __half22float2(f0[k]), a1 = __half22float2(f1[k]), a2 = __half22float2(f2[k]);
        const float x = __fadd_rn(__fadd_rn(__fmul_rn(a0.x, w0), __fmul_rn(a1.x, w1)), __fmul_rn(a2.x, w2));
        const float y = __fadd_rn(__fadd_rn(__fmul_rn(a0.y, w0), __fmul_rn(a1.y, w1)), __fmul_rn(a2.y, w2));
        oi[k] = __floats2half2_rn(fminf(x, 65504.f), fminf(y, 65504.f));
    }
}

// 3-NN search (indices + weights only, what MODE 1 of k_interp produces) with the slab bitmaps of k_ball_query over the
// COARSE points: the candidates within `reach` of the fine point's slabs are searched first; if the third-nearest of them is
// closer than every point outside the candidate set can be (squared distance <= ok_d < reach^2 minus the rounding slack of the
// expansion formula), the result is that of the exhaustive scan -- same distances, same (distance, index) order; otherwise the
// thread falls back to the exhaustive scan.
__global__ void __launch_bounds__(512) k_knn_pruned(const float *xyz1, int N, const float *xyz2, int S, int4 *knn_i, float4 *knn_w,
                                                    float reach, float ok_d) {
    extern __shared__ float4 s_q4[];          // [S] (x, y, z, |q|^2), then the bitmaps [3][kSlabs][W]
    const int Wn = (S + 31) >> 5, W = Wn + 1;   // row stride Wn + 1: the rows of different slabs start in different banks
    unsigned *bm = reinterpret_cast<unsigned *>(s_q4 + S);
    const int b = blockIdx.y;
    const float *Q = xyz2 + (size_t)b * S * 3;
    for (int i = threadIdx.x; i < 3 * kSlabs * W; i += blockDim.x) bm[i] = 0u;
    __syncthreads();
    for (int i = threadIdx.x; i < S; i += blockDim.x) {
        const float v[3] = {Q[i * 3], Q[i * 3 + 1], Q[i * 3 + 2]};
        s_q4[i] = make_float4(v[0], v[1], v[2], __fadd_rn(__fadd_rn(__fmul_rn(v[0], v[0]), __fmul_rn(v[1], v[1])), __fmul_rn(v[2], v[2])));
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int sl = 0; sl < kSlabs; sl++) {
                const bool lo_ok = sl == 0 || v[a] >= slab_lo(sl) - reach;
                const bool hi_ok = sl == kSlabs - 1 || v[a] <= slab_lo(sl + 1) + reach;
                if (lo_ok && hi_ok) atomicOr(&bm[(a * kSlabs + sl) * W + (i >> 5)], 1u << (i & 31));
            }
    }
    __syncthreads();
    const int pi = blockIdx.x * blockDim.x + threadIdx.x;
    if (pi >= N) return;
    const float *a = xyz1 + ((size_t)b * N + pi) * 3;
    const float ax = a[0], ay = a[1], az = a[2];
    const float as = __fadd_rn(__fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay)), __fmul_rn(az, az));
    int ix = 0, iy = 0, iz = 0;
#pragma unroll
    for (int sl = 1; sl < kSlabs; sl++) {
        const float lo = slab_lo(sl);
        ix += ax >= lo ? 1 : 0; iy += ay >= lo ? 1 : 0; iz += az >= lo ? 1 : 0;
    }
    const unsigned *bx = bm + (0 * kSlabs + ix) * W, *by = bm + (1 * kSlabs + iy) * W, *bz = bm + (2 * kSlabs + iz) * W;
    Top3 t;
    t.d[0] = t.d[1] = t.d[2] = INFINITY; t.i[0] = t.i[1] = t.i[2] = 0;
    auto visit = [&](int s) {
        const float4 q = s_q4[s];
        const float dot = __fmaf_rn(az, q.z, __fmaf_rn(ay, q.y, __fmul_rn(ax, q.x)));
        float d = __fmul_rn(-2.f, dot);
        d = __fadd_rn(d, as);
        d = __fadd_rn(d, q.w);
        top3_insert(t, d, s);                 // strict <: equal distances keep the lower index first (ascending visits)
    };
    for (int w = 0; w < Wn; w++) {
        unsigned m = bx[w] & by[w] & bz[w];
        while (m) { visit((w << 5) + __ffs(m) - 1); m &= m - 1; }
    }
    if (!(t.d[2] <= ok_d)) {                  // a point outside the candidate set could be among the three: exhaustive scan
        t.d[0] = t.d[1] = t.d[2] = INFINITY; t.i[0] = t.i[1] = t.i[2] = 0;
        for (int s = 0; s < S; s++) visit(s);
    }
    const float r0 = __fdiv_rn(1.f, __fadd_rn(t.d[0], 1e-8f)), r1 = __fdiv_rn(1.f, __fadd_rn(t.d[1], 1e-8f)),
                r2 = __fdiv_rn(1.f, __fadd_rn(t.d[2], 1e-8f));
    const float norm = __fadd_rn(__fadd_rn(r0, r1), r2);
    knn_i[(size_t)b * N + pi] = make_int4(t.i[0], t.i[1], t.i[2], 0);
    knn_w[(size_t)b * N + pi] = make_float4(__fdiv_rn(r0, norm), __fdiv_rn(r1, norm), __fdiv_rn(r2, norm), 0.f);
}

template <int MODE>
__global__ void __launch_bounds__(256) k_interp(const float *xyz1, int N, const float *xyz2, int S, const __half *feat1, int C1,
                                                const __half *feat2, int C2, int B, __half *out, int4 *knn_i, float4 *knn_w) {
    // phase 1: one thread per fine point walks the S coarse points (broadcast LDS.128) keeping its
    // three nearest in ascending (distance, index) order -- the first three entries of the
    // reference's sort; phase 2: the warp's 32 results are broadcast one at a time and all lanes
    // interpolate / copy the channels of that point.
    // coarse points in pairs (S is even at every level): s_qq[2k] = (x_2k, x_2k+1, y_2k, y_2k+1), s_qq[2k + 1] = (z.., |q|^2..)
    extern __shared__ float4 s_qq[];
    const int b = blockIdx.y;
    const float *Q = xyz2 + (size_t)b * S * 3;
    if (MODE != 2) {
        for (int k = threadIdx.x; 2 * k < S; k += blockDim.x) {
            float c[2][4];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int i = 2 * k + u;
                const float x = Q[i * 3], y = Q[i * 3 + 1], z = Q[i * 3 + 2];
                c[u][0] = x; c[u][1] = y; c[u][2] = z;
                c[u][3] = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
            }
            s_qq[2 * k] = make_float4(c[0][0], c[1][0], c[0][1], c[1][1]);
            s_qq[2 * k + 1] = make_float4(c[0][2], c[1][2], c[0][3], c[1][3]);
        }
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const int i0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31);     // first fine point of this warp
    const int pi = i0 + lane;
    Top3 t;
    t.d[0] = t.d[1] = t.d[2] = INFINITY; t.i[0] = t.i[1] = t.i[2] = 0;
    float w0 = 0.f, w1 = 0.f, w2 = 0.f;
    if (MODE == 2) {
        if (pi < N) {
            const int4 ki = knn_i[(size_t)b * N + pi]; const float4 kw = knn_w[(size_t)b * N + pi];
            t.i[0] = ki.x; t.i[1] = ki.y; t.i[2] = ki.z; w0 = kw.x; w1 = kw.y; w2 = kw.z;
        }
    } else {
    if (pi < N) {
        const float *a = xyz1 + ((size_t)b * N + pi) * 3;
        const float ax = a[0], ay = a[1], az = a[2];
        const float as = __fadd_rn(__fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay)), __fmul_rn(az, az));
        const f32x2 ax2 = pack2(ax, ax), ay2 = pack2(ay, ay), az2 = pack2(az, az), as2 = pack2(as, as), m2 = pack2(-2.f, -2.f);
        for (int s = 0; s < S; s += 2) {      // two coarse points per step on the packed fp32 pipe, the reference's roundings
            const float4 qa = s_qq[s], qb = s_qq[s + 1];
            const f32x2 dot = fma2(az2, pack2(qb.x, qb.y), fma2(ay2, pack2(qa.z, qa.w), mul2(ax2, pack2(qa.x, qa.y))));
            float d0, d1;
            unpack2(add2(add2(mul2(m2, dot), as2), pack2(qb.z, qb.w)), d0, d1);      // -2 * dot is exact
            if (fminf(d0, d1) < t.d[2]) {     // rare once the three running values have settled
                top3_insert(t, d0, s);        // strict <: equal distances keep the lower index first
                top3_insert(t, d1, s + 1);
            }
        }
    }
    const float r0 = __fdiv_rn(1.f, __fadd_rn(t.d[0], 1e-8f)), r1 = __fdiv_rn(1.f, __fadd_rn(t.d[1], 1e-8f)),
                r2 = __fdiv_rn(1.f, __fadd_rn(t.d[2], 1e-8f));
    const float norm = __fadd_rn(__fadd_rn(r0, r1), r2);
    w0 = __fdiv_rn(r0, norm); w1 = __fdiv_rn(r1, norm); w2 = __fdiv_rn(r2, norm);
    if (MODE == 1) {
        if (pi < N) { knn_i[(size_t)b * N + pi] = make_int4(t.i[0], t.i[1], t.i[2], 0); knn_w[(size_t)b * N + pi] = make_float4(w0, w1, w2, 0.f); }
        return;
    }
    }
    const int cnt = min(32, N - i0);
    // LPP lanes share a point, one 16-byte chunk (8 channels) per lane and step; 32 / LPP points per warp step.  Per element
    // the arithmetic is the reference's (a0 * w0 + a1 * w1) + a2 * w2 in fp32, whatever the vector width.
    const int chunks2 = C2 >> 3;
    const int LPP = chunks2 >= 32 ? 32 : (chunks2 >= 16 ? 16 : 8);
    const int sub = lane / LPP, c = lane % LPP, PPW = 32 / LPP;
    for (int j0 = 0; j0 < cnt; j0 += PPW) {
        const int j = j0 + sub, sl = min(j, 31);
        const int n0 = __shfl_sync(0xffffffffu, t.i[0], sl), n1 = __shfl_sync(0xffffffffu, t.i[1], sl), n2 = __shfl_sync(0xffffffffu, t.i[2], sl);
        const float u0 = __shfl_sync(0xffffffffu, w0, sl), u1 = __shfl_sync(0xffffffffu, w1, sl), u2 = __shfl_sync(0xffffffffu, w2, sl);
        if (j >= cnt) continue;
        const size_t p = (size_t)b * N + i0 + j;
        __half *o = out + p * (size_t)(C1 + C2);
        if (C1 > 0) {
            const uint4 *src = reinterpret_cast<const uint4 *>(feat1 + p * (size_t)C1);
            uint4 *dst = reinterpret_cast<uint4 *>(o);
            for (int k = c; k < (C1 >> 3); k += LPP) dst[k] = src[k];
        }
        const uint4 *f0 = reinterpret_cast<const uint4 *>(feat2 + ((size_t)b * S + n0) * C2);
        const uint4 *f1 = reinterpret_cast<const uint4 *>(feat2 + ((size_t)b * S + n1) * C2);
        const uint4 *f2 = reinterpret_cast<const uint4 *>(feat2 + ((size_t)b * S + n2) * C2);
        uint4 *oi = reinterpret_cast<uint4 *>(o + C1);
        for (int k = c; k < chunks2; k += LPP) {
            const uint4 q0 = __ldg(f0 + k), q1 = __ldg(f1 + k), q2 = __ldg(f2 + k);
            const __half2 *h0 = reinterpret_cast<const __half2 *>(&q0), *h1 = reinterpret_cast<const __half2 *>(&q1),
                          *h2 = reinterpret_cast<const __half2 *>(&q2);
            uint4 r;
            __half2 *hr = reinterpret_cast<__half2 *>(&r);
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const float2 a0 = __half22float2(h0[e]), a1 = __half22float2(h1[e]), a2 = __half22float2(h2[e]);
                const float x = __fadd_rn(__fadd_rn(__fmul_rn(a0.x, u0), __fmul_rn(a1.x, u1)), __fmul_rn(a2.x, u2));
                const float y = __fadd_rn(__fadd_rn(__fmul_rn(a0.y, u0), __fmul_rn(a1.y, u1)), __fmul_rn(a2.y, u2));
                hr[e] = __floats2half2_rn(fminf(x, 65504.f), fminf(y, 65504.f));
            }
            oi[k] = r;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_head: conv2 (128 -> 2) + log_softmax + argmax + softmax[:,1] (pointnet2.py:39-41,
// pointnet2_wrapper.py:61-62).  fp32 weights; one thread per point.
__global__ void __launch_bounds__(128) k_head(const __half *x, const float *w2, const float *b2, size_t total,
                                              long long *pred, float *score, float *logp) {
    __shared__ float s_w[2 * 128 + 2];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_w[i] = w2[i];
    if (threadIdx.x < 2) s_w[256 + threadIdx.x] = b2[threadIdx.x];
    __syncthreads();
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= total) return;
    const uint4 *row = reinterpret_cast<const uint4 *>(x + p * 128);
    float z0 = 0.f, z1 = 0.f;
#pragma unroll 4
    for (int k = 0; k < 16; k++) {
        const uint4 v = row[k];
        const __half2 *h = reinterpret_cast<const __half2 *>(&v);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float2 f = __half22float2(h[j]);
            const int c = k * 8 + j * 2;
            z0 = fmaf(f.x, s_w[c], z0); z0 = fmaf(f.y, s_w[c + 1], z0);
            z1 = fmaf(f.x, s_w[128 + c], z1); z1 = fmaf(f.y, s_w[128 + c + 1], z1);
        }
    }
    z0 += s_w[256]; z1 += s_w[257];
    const float m = fmaxf(z0, z1);
    const float lse = m + logf(expf(z0 - m) + expf(z1 - m));
    const float l0 = z0 - lse, l1 = z1 - lse;
    if (logp) { logp[p * 2] = l0; logp[p * 2 + 1] = l1; }
    pred[p] = (l1 > l0) ? 1 : 0;                      // np.argmax: first maximum
    const float e0 = expf(l0), e1 = expf(l1);
    score[p] = e1 / (e0 + e1);
}

// ------------------------------------------------------------------------------------------------
// host side

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;

static int load_encode() {
    if (g_encode) return NIRRT_OK;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn)
        return pfail(NIRRT_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    g_encode = (PFN_encodeTiled)fn;
    return NIRRT_OK;
}

// tensor map of a row-major fp16 matrix [rows][cols] with a {64, box_rows} box and 128-byte swizzle
static int make_tmap(CUtensorMap *m, const void *base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    PTRY(load_encode());
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)umma::kBK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return pfail(NIRRT_ERR_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    return NIRRT_OK;
}

static bool g_gemm_attr = false;
static int gemm_attr() {
    if (g_gemm_attr) return NIRRT_OK;
    PCUDA(cudaFuncSetAttribute(umma::k_gemm<umma::MODE_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    PCUDA(cudaFuncSetAttribute(umma::k_gemm<umma::MODE_POOL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    PCUDA(cudaFuncSetAttribute(k_ball_query, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
    PCUDA(cudaFuncSetAttribute(k_ball_query_bf, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024));
    PCUDA(cudaFuncSetAttribute(k_prep, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    PCUDA(cudaFuncSetAttribute(k_ball_query_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    PCUDA(cudaFuncSetAttribute(k_fps<256, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));   // 4096 points x 12 B + static
    PCUDA(cudaFuncSetAttribute(k_fps_bucket<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    g_gemm_attr = true;
    return NIRRT_OK;
}

// N tile per CTA (multiple of 16).  NIRRT_PN2_BN_MAX (development knob, default 256): narrower tiles need fewer accumulator
// columns, so more CTAs fit next to each other, at the price of re-reading the A tile once per N tile.
static int pick_bn(int N) {
    static const int bn_max = getenv("NIRRT_PN2_BN_MAX") ? atoi(getenv("NIRRT_PN2_BN_MAX")) : 256;
    if (bn_max >= 256 || bn_max < 32) return N <= 256 ? N : N / 2;
    if (N <= bn_max) return N;
    const int parts = (N + bn_max - 1) / bn_max;
    return (((N + parts - 1) / parts) + 15) & ~15;
}

struct Conv {
    int K = 0, N = 0, BN = 0;      // padded dims (multiples of 16)
    __half *w = nullptr;           // [N][K]
    float *b = nullptr;            // [N]
    CUtensorMap tmW;
    std::vector<__half> hw;        // host copies (used to build the fused kernels' shared-memory images)
    std::vector<float> hb;
};

static int g_num_sms = 0;

static int launch_gemm(const Conv &c, const __half *A, int M, int mode, __half *out, int ldo, int col_off, int group,
                       cudaStream_t s) {
    PTRY(gemm_attr());
    if (!g_num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    CUtensorMap tmA;
    PTRY(make_tmap(&tmA, A, (uint64_t)M, (uint64_t)c.K, umma::kBM));
    umma::GemmArgs g;
    g.M = M; g.N = c.N; g.K = c.K; g.BN = c.BN;
    g.ldo = ldo; g.col_off = col_off; g.group = group; g.bias = c.b; g.out = out;
    // CTAs per SM: bounded by TMEM (two accumulator buffers each) and shared memory; the stage count is
    // the largest of 4/3/2 that keeps that many CTAs resident
    int cps = 512 / (2 * umma::tmem_cols_for(c.BN));
    if (cps > 8) cps = 8;
    if (cps < 1) cps = 1;
    const int nkb = (c.K + umma::kBK - 1) / umma::kBK;
    int stages = nkb == 1 ? 2 : 4;      // one K block per tile: two stages already overlap consecutive tiles
    for (;; ) {
        const umma::SmemLayout L = umma::smem_layout(g.BN, stages);
        if ((size_t)cps * (L.total + 1024 + 1024) <= 224 * 1024) break;
        if (stages > 2) stages--;
        else if (cps > 1) { cps--; stages = nkb == 1 ? 2 : 4; }
        else break;
    }
    g.stages = stages;
    const umma::SmemLayout L = umma::smem_layout(g.BN, g.stages);
    const int ntiles = (M + umma::kBM - 1) / umma::kBM;
    const int gx = ntiles < g_num_sms * cps ? ntiles : g_num_sms * cps;
    const dim3 grid(gx, (c.N + c.BN - 1) / c.BN);
    const size_t smem = (size_t)L.total + 1024;
    if (mode == umma::MODE_STORE) umma::k_gemm<umma::MODE_STORE><<<grid, umma::kThreads, smem, s>>>(tmA, c.tmW, g);
    else umma::k_gemm<umma::MODE_POOL><<<grid, umma::kThreads, smem, s>>>(tmA, c.tmW, g);
    PCUDA(cudaGetLastError());
    return NIRRT_OK;
}

static const int kNp[5] = {0, 1024, 256, 64, 16};            // npoint of sa1..sa4 (pointnet2.py:11-14)
static const int kC[5] = {6, 96, 256, 512, 1024};            // channels of l0..l4 features
static const double kRad[4][2] = {{0.05, 0.1}, {0.1, 0.2}, {0.2, 0.4}, {0.4, 0.8}};
static const int kK[2] = {16, 32};
static const int kUpC[4] = {256, 256, 128, 128};             // outputs of fp4, fp3, fp2, fp1

struct nirrt_pn2 {
    int N0 = 0, maxB = 0, device = 0;
    int N0cap = 0;     // cloud size the buffers were allocated for (nirrt_pn2_set_n_points can select any size up to it)
    int n[5];
    Conv conv[34];
    float *w2 = nullptr, *b2 = nullptr;
    float *xyz[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    float *in6 = nullptr;
    __half *feat[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    int *fidx[4] = {nullptr, nullptr, nullptr, nullptr};
    float *fps_dist = nullptr;      // [maxB][N0cap] running distances carried between the chunks of a level's selection
    int *fps_far = nullptr;         // [maxB]
    int *grp[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    __half *bufA = nullptr, *bufB = nullptr;
    uint8_t *sa3_img[2] = {nullptr, nullptr};      // sa3: gather + first two layers fused (weight images of conv 12,13 / 15,16)
    float *sa3_bias[2] = {nullptr, nullptr};
    uint8_t *fp_img = nullptr;      // fused fp1 + conv1 + head kernel: four 128 x 128 weight images (null: layer by layer)
    float *fp_bias = nullptr;
    uint8_t *sa_img[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // fused sa1 / sa2 kernels: swizzled weight
    float *sa_bias[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};    // images + biases per level and radius
    int fused_levels = 2;                                               // 0: none, 1: sa1, 2: sa1 + sa2
    __half *up[4] = {nullptr, nullptr, nullptr, nullptr};     // outputs of fp4 (level 3) .. fp1 (level 0)
    // staging for the host-pointer entry point
    float *d_pc = nullptr, *d_sm = nullptr, *d_gm = nullptr, *d_score = nullptr, *d_logp = nullptr;
    int *d_fps = nullptr;
    long long *d_pred = nullptr;
    std::vector<void *> allocs;
    int lastB = 0;
    int64_t launches = 0;
    bool profiling = false;
    float stage_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    // geometry stream: everything that needs coordinates only (FPS, ball query, 3-NN of the feature
    // propagation levels) runs beside the feature stream's MLPs; events hand each level over
    cudaStream_t gs = nullptr, gs2 = nullptr;      // FPS chain | ball queries + 3-NN searches
    cudaEvent_t ev_fc[8] = {}, ev_bc[8] = {}, ev_fl[4] = {};      // level-1 FPS chunk / ball-query chunk done; FPS of a level done
    int l1_chunks = 1;
    bool fps_bucket = false;                       // NIRRT_PN2_FPS_BUCKET = 1: bucket-pruned selection for the larger levels
    bool bq_pruned = true, knn_pruned = true;     // slab-pruned ball query / 3-NN search (NIRRT_PN2_BQ / NIRRT_PN2_KNN = 0: exhaustive)
    float knn_reach = 0.3f;                        // NIRRT_PN2_KNN_REACH: candidate radius of the pruned 3-NN search
    cudaEvent_t ev_fork = nullptr, ev_bq[4] = {nullptr, nullptr, nullptr, nullptr}, ev_knn[4] = {nullptr, nullptr, nullptr, nullptr};
    int4 *knn_i[4] = {nullptr, nullptr, nullptr, nullptr};      // per FP level: the three nearest coarse points of every fine point
    float4 *knn_w[4] = {nullptr, nullptr, nullptr, nullptr};    // and their normalised inverse-distance weights
    bool two_streams = true;                                    // NIRRT_PN2_STREAMS=1 disables
};

template <typename T>
static int palloc(nirrt_pn2 *h, T **p, size_t count) {
    void *q = nullptr;
    cudaError_t e = nirrt_dev_malloc(&q, sizeof(T) * (count ? count : 1));
    if (e != cudaSuccess) return pfail(NIRRT_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    h->allocs.push_back(q);
    *p = (T *)q;
    return NIRRT_OK;
}

static int round16(int x) { return (x + 15) & ~15; }

// folds BatchNorm into the convolution, remaps/pads the input channels and uploads fp16 rows
//   remap: for each padded input column k the source column (or -1 for zero)
static int upload_conv(nirrt_pn2 *h, Conv &c, const nirrt_pn2_layer &L, const std::vector<int> &remap) {
    c.K = (int)remap.size();
    c.N = round16(L.c_out);
    c.BN = pick_bn(c.N);
    std::vector<__half> w((size_t)c.N * c.K, __float2half(0.f));
    std::vector<float> b((size_t)c.N, 0.f);
    for (int n = 0; n < L.c_out; n++) {
        float s = 1.f, shift = L.bias[n];
        if (L.bn_weight) {
            s = L.bn_weight[n] / sqrtf(L.bn_var[n] + 1e-5f);
            shift = (L.bias[n] - L.bn_mean[n]) * s + L.bn_bias[n];
        }
        b[n] = shift;
        for (int k = 0; k < c.K; k++)
            if (remap[k] >= 0) w[(size_t)n * c.K + k] = __float2half_rn(L.weight[(size_t)n * L.c_in + remap[k]] * s);
    }
    c.hw = w; c.hb = b;
    PTRY(palloc(h, &c.w, w.size()));
    PTRY(palloc(h, &c.b, b.size()));
    PCUDA(cudaMemcpy(c.w, w.data(), w.size() * sizeof(__half), cudaMemcpyHostToDevice));
    PCUDA(cudaMemcpy(c.b, b.data(), b.size() * sizeof(float), cudaMemcpyHostToDevice));
    PTRY(make_tmap(&c.tmW, c.w, (uint64_t)c.N, (uint64_t)c.K, (uint32_t)c.BN));
    return NIRRT_OK;
}

static std::vector<int> identity_remap(int c_in) {
    std::vector<int> r(round16(c_in), -1);
    for (int k = 0; k < c_in; k++) r[k] = k;
    return r;
}

extern "C" int nirrt_pn2_destroy(nirrt_pn2 *h) {
    if (!h) return NIRRT_OK;
    cudaSetDevice(h->device);
    for (void *p : h->allocs) nirrt_dev_free(p);
    for (int i = 0; i < 2; i++) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    if (h->gs) cudaStreamDestroy(h->gs);
    if (h->gs2) cudaStreamDestroy(h->gs2);
    for (int i = 0; i < 8; i++) { if (h->ev_fc[i]) cudaEventDestroy(h->ev_fc[i]); if (h->ev_bc[i]) cudaEventDestroy(h->ev_bc[i]); }
    for (int i = 0; i < 4; i++) if (h->ev_fl[i]) cudaEventDestroy(h->ev_fl[i]);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    for (int i = 0; i < 4; i++) { if (h->ev_bq[i]) cudaEventDestroy(h->ev_bq[i]); if (h->ev_knn[i]) cudaEventDestroy(h->ev_knn[i]); }
    delete h;
    return NIRRT_OK;
}

extern "C" int nirrt_pn2_create(const nirrt_pn2_layer *layers, int n_layers, int n_points, int max_batch, int device,
                                nirrt_pn2 **out) {
    if (!layers || !out) return pfail(NIRRT_ERR_INVALID, "nirrt_pn2_create: null argument");
    if (n_layers != NIRRT_PN2_NUM_LAYERS) return pfail(NIRRT_ERR_INVALID, "nirrt_pn2_create: expected 35 layers");
    // any cloud size the reference's forward accepts up to 4096 points: with fewer than 1024 points sa1's farthest point
    // sampling re-selects index 0 once every point is taken (torch.max on all-zero distances), which k_fps reproduces
    if (n_points < 16 || n_points > 4096) return pfail(NIRRT_ERR_INVALID, "n_points must be in [16, 4096]");
    if (max_batch < 1 || max_batch > 32768) return pfail(NIRRT_ERR_INVALID, "1 <= max_batch <= 32768 required (row indices are 32-bit)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return pfail(NIRRT_ERR_NO_DEVICE, "no CUDA device visible");
    if (device < 0 || device >= ndev) return pfail(NIRRT_ERR_INVALID, "bad device ordinal");
    cudaDeviceProp prop;
    PCUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return pfail(NIRRT_ERR_NO_DEVICE, "libnirrt_b200 is built for sm_100a only");
    PCUDA(cudaSetDevice(device));
    // expected channel plan
    static const int sa_mlp[4][2][3] = {{{16, 16, 32}, {32, 32, 64}}, {{64, 64, 128}, {64, 96, 128}},
                                        {{128, 196, 256}, {128, 196, 256}}, {{256, 256, 512}, {256, 384, 512}}};
    nirrt_pn2 *h = new nirrt_pn2();
    h->N0 = n_points; h->N0cap = n_points; h->maxB = max_batch; h->device = device;
    h->n[0] = n_points;
    for (int l = 1; l <= 4; l++) h->n[l] = kNp[l];
#define FAILC(msg) do { nirrt_pn2_destroy(h); return pfail(NIRRT_ERR_INVALID, msg); } while (0)
#define TRYC(expr) do { int _r = (expr); if (_r) { nirrt_pn2_destroy(h); return _r; } } while (0)
    int li = 0;
    for (int l = 1; l <= 4; l++)
        for (int sc = 0; sc < 2; sc++) {
            int last = kC[l - 1] + 3;
            for (int j = 0; j < 3; j++, li++) {
                const nirrt_pn2_layer &L = layers[li];
                if (L.c_in != last || L.c_out != sa_mlp[l - 1][sc][j] || !L.weight || !L.bias || !L.bn_weight || !L.bn_bias || !L.bn_mean || !L.bn_var)
                    FAILC("nirrt_pn2_create: SA layer " + std::to_string(li) + " has unexpected shape");
                std::vector<int> remap;
                if (j == 0) {
                    const int C = kC[l - 1];
                    if (l == 1) {
                        remap.assign(16, -1);
                        for (int k = 0; k < 9; k++) remap[k] = k;
                        for (int k = 0; k < 3; k++) { remap[9 + k] = k; remap[12 + k] = 6 + k; }
                    } else {
                        remap.assign(round16(C + 6), -1);
                        for (int k = 0; k < C + 3; k++) remap[k] = k;
                        for (int k = 0; k < 3; k++) remap[C + 3 + k] = C + k;
                    }
                } else remap = identity_remap(last);
                TRYC(upload_conv(h, h->conv[li], L, remap));
                last = L.c_out;
            }
        }
    static const int fp_in[4] = {1536, 512, 352, 128};
    static const int fp_mlp[4][3] = {{256, 256, 0}, {256, 256, 0}, {256, 128, 0}, {128, 128, 128}};
    for (int f = 0; f < 4; f++) {
        int last = fp_in[f];
        for (int j = 0; j < 3 && fp_mlp[f][j]; j++, li++) {
            const nirrt_pn2_layer &L = layers[li];
            if (L.c_in != last || L.c_out != fp_mlp[f][j] || !L.weight || !L.bias || !L.bn_weight)
                FAILC("nirrt_pn2_create: FP layer " + std::to_string(li) + " has unexpected shape");
            TRYC(upload_conv(h, h->conv[li], L, identity_remap(last)));
            last = L.c_out;
        }
    }
    if (layers[33].c_in != 128 || layers[33].c_out != 128 || !layers[33].bn_weight) FAILC("nirrt_pn2_create: conv1 has unexpected shape");
    TRYC(upload_conv(h, h->conv[33], layers[33], identity_remap(128)));
    if (layers[34].c_in != 128 || layers[34].c_out != 2) FAILC("nirrt_pn2_create: conv2 must be 128 -> 2 (num_classes = 2)");
    TRYC(palloc(h, &h->w2, 256)); TRYC(palloc(h, &h->b2, 2));
    if (cudaMemcpy(h->w2, layers[34].weight, 256 * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(h->b2, layers[34].bias, 2 * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess)
        FAILC("nirrt_pn2_create: weight upload failed");
    // activations
    const size_t B = (size_t)max_batch;
    for (int l = 0; l <= 4; l++) TRYC(palloc(h, &h->xyz[l], B * h->n[l] * 3));
    TRYC(palloc(h, &h->in6, B * h->N0 * 6));
    TRYC(palloc(h, &h->fps_dist, B * h->N0));
    TRYC(palloc(h, &h->fps_far, B));
    for (int l = 1; l <= 4; l++) {
        TRYC(palloc(h, &h->feat[l], B * h->n[l] * kC[l]));
        TRYC(palloc(h, &h->fidx[l - 1], B * h->n[l]));
        for (int sc = 0; sc < 2; sc++) TRYC(palloc(h, &h->grp[(l - 1) * 2 + sc], B * h->n[l] * kK[sc]));
    }
    size_t per_cloud = 1024 * 32 * 32;                     // sa1 scale 1: 32768 rows x 32 channels
    const size_t cand[] = {(size_t)256 * 32 * 112, (size_t)64 * 32 * 272, (size_t)16 * 32 * 528, (size_t)64 * 1536,
                           (size_t)256 * 512, (size_t)1024 * 352, (size_t)h->N0 * 128};
    for (size_t c : cand) if (c > per_cloud) per_cloud = c;
    TRYC(palloc(h, &h->bufA, B * per_cloud + 4096));
    TRYC(palloc(h, &h->bufB, B * per_cloud + 4096));
    for (int f = 0; f < 4; f++) TRYC(palloc(h, &h->up[f], B * h->n[3 - f] * kUpC[f]));
    for (int f = 0; f < 4; f++) { TRYC(palloc(h, &h->knn_i[f], B * h->n[3 - f])); TRYC(palloc(h, &h->knn_w[f], B * h->n[3 - f])); }
    {
        const char *st = getenv("NIRRT_PN2_STREAMS");
        h->two_streams = !(st && atoi(st) == 1);
        if (cudaStreamCreateWithFlags(&h->gs, cudaStreamNonBlocking) != cudaSuccess) FAILC("nirrt_pn2_create: cudaStreamCreate failed");
        if (cudaStreamCreateWithFlags(&h->gs2, cudaStreamNonBlocking) != cudaSuccess) FAILC("nirrt_pn2_create: cudaStreamCreate failed");
        for (int i = 0; i < 8; i++) { cudaEventCreateWithFlags(&h->ev_fc[i], cudaEventDisableTiming); cudaEventCreateWithFlags(&h->ev_bc[i], cudaEventDisableTiming); }
        for (int i = 0; i < 4; i++) cudaEventCreateWithFlags(&h->ev_fl[i], cudaEventDisableTiming);
        if (getenv("NIRRT_PN2_FPS_BUCKET")) h->fps_bucket = atoi(getenv("NIRRT_PN2_FPS_BUCKET")) != 0;
        if (getenv("NIRRT_PN2_BQ")) h->bq_pruned = atoi(getenv("NIRRT_PN2_BQ")) != 0;
        if (getenv("NIRRT_PN2_KNN")) h->knn_pruned = atoi(getenv("NIRRT_PN2_KNN")) != 0;
        if (getenv("NIRRT_PN2_KNN_REACH")) { const float r = (float)atof(getenv("NIRRT_PN2_KNN_REACH")); if (r > 0.005f && r < 1.f) h->knn_reach = r; }
        const char *ch = getenv("NIRRT_PN2_CHUNKS");
        if (ch) { const int c = atoi(ch); h->l1_chunks = (c == 1 || c == 2 || c == 4 || c == 8) ? c : 1; }
        cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
        for (int i = 0; i < 4; i++) { cudaEventCreateWithFlags(&h->ev_bq[i], cudaEventDisableTiming); cudaEventCreateWithFlags(&h->ev_knn[i], cudaEventDisableTiming); }
    }
    TRYC(palloc(h, &h->d_pc, B * h->N0 * 3)); TRYC(palloc(h, &h->d_sm, B * h->N0)); TRYC(palloc(h, &h->d_gm, B * h->N0));
    TRYC(palloc(h, &h->d_fps, B * 4)); TRYC(palloc(h, &h->d_pred, B * h->N0)); TRYC(palloc(h, &h->d_score, B * h->N0));
    TRYC(palloc(h, &h->d_logp, B * h->N0 * 2));
    for (int i = 0; i < 2; i++)
        if (cudaEventCreate(&h->ev[i]) != cudaSuccess) FAILC("cudaEventCreate failed");
    if (gemm_attr()) { nirrt_pn2_destroy(h); return NIRRT_ERR_CUDA; }
    {
        const char *f = getenv("NIRRT_PN2_FUSED");
        h->fused_levels = f ? atoi(f) : 2;
        if (h->fused_levels < 0 || h->fused_levels > 2) h->fused_levels = 2;
        for (int l = 0; l < h->fused_levels; l++)
            for (int sc = 0; sc < 2; sc++) {
                size_t bytes = 0;
                for (int j = 0; j < 3; j++) { const Conv &c = h->conv[l * 6 + sc * 3 + j]; bytes += (size_t)c.N * safused::nblk(c.K) * 128; }
                std::vector<uint8_t> img(bytes, 0);
                std::vector<float> bias;
                size_t base = 0;
                for (int j = 0; j < 3; j++) {
                    const Conv &c = h->conv[l * 6 + sc * 3 + j];
                    for (int n = 0; n < c.N; n++)
                        for (int k = 0; k < c.K; k++)
                            memcpy(&img[base + safused::w_off(c.N, n, k >> 3) + (k & 7) * 2], &c.hw[(size_t)n * c.K + k], 2);
                    bias.insert(bias.end(), c.hb.begin(), c.hb.end());
                    base += (size_t)c.N * safused::nblk(c.K) * 128;
                }
                TRYC(palloc(h, &h->sa_img[l][sc], img.size()));
                TRYC(palloc(h, &h->sa_bias[l][sc], bias.size()));
                if (cudaMemcpy(h->sa_img[l][sc], img.data(), img.size(), cudaMemcpyHostToDevice) != cudaSuccess ||
                    cudaMemcpy(h->sa_bias[l][sc], bias.data(), bias.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess)
                    FAILC("fused SA: weight image upload failed");
            }
        {
            const char *ff = getenv("NIRRT_PN2_FP_FUSED");
            bool ok = !(ff && atoi(ff) == 0);
            for (int j = 30; j < 34; j++) ok = ok && h->conv[j].K == safused::kFpC && h->conv[j].N == safused::kFpC;
            if (ok) {
                std::vector<uint8_t> img((size_t)4 * 32768, 0);
                std::vector<float> bias;
                for (int j = 0; j < 4; j++) {
                    const Conv &c = h->conv[30 + j];
                    for (int n = 0; n < c.N; n++)
                        for (int k = 0; k < c.K; k++)
                            memcpy(&img[(size_t)j * 32768 + safused::w_off(c.N, n, k >> 3) + (k & 7) * 2], &c.hw[(size_t)n * c.K + k], 2);
                    bias.insert(bias.end(), c.hb.begin(), c.hb.end());
                }
                TRYC(palloc(h, &h->fp_img, img.size()));
                TRYC(palloc(h, &h->fp_bias, bias.size()));
                if (cudaMemcpy(h->fp_img, img.data(), img.size(), cudaMemcpyHostToDevice) != cudaSuccess ||
                    cudaMemcpy(h->fp_bias, bias.data(), bias.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
                    cudaFuncSetAttribute(safused::k_fp1_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)safused::kFpSmem) != cudaSuccess)
                    FAILC("fused fp1: weight image upload failed");
            }
        }
        {
            // sa3: [256 features + xyz -> 128 -> 196 (padded 208)] fused up to the second layer (NIRRT_PN2_SA3_FUSED=0: layer by layer)
            const char *ff = getenv("NIRRT_PN2_SA3_FUSED");
            bool ok = !(ff && atoi(ff) == 0) && h->fused_levels >= 2;
            ok = ok && h->conv[12].K == 272 && h->conv[12].N == 128 && h->conv[13].N == 208 && h->conv[15].K == 272 && h->conv[15].N == 128 &&
                 h->conv[16].N == 208;
            for (int sc = 0; sc < 2 && ok; sc++) {
                size_t bytes = 0;
                for (int j = 0; j < 2; j++) { const Conv &c = h->conv[12 + sc * 3 + j]; bytes += (size_t)c.N * safused::nblk(c.K) * 128; }
                std::vector<uint8_t> img(bytes, 0);
                std::vector<float> bias;
                size_t base = 0;
                for (int j = 0; j < 2; j++) {
                    const Conv &c = h->conv[12 + sc * 3 + j];
                    for (int n = 0; n < c.N; n++)
                        for (int k = 0; k < c.K; k++)
                            memcpy(&img[base + safused::w_off(c.N, n, k >> 3) + (k & 7) * 2], &c.hw[(size_t)n * c.K + k], 2);
                    bias.insert(bias.end(), c.hb.begin(), c.hb.end());
                    base += (size_t)c.N * safused::nblk(c.K) * 128;
                }
                TRYC(palloc(h, &h->sa3_img[sc], img.size()));
                TRYC(palloc(h, &h->sa3_bias[sc], bias.size()));
                if (cudaMemcpy(h->sa3_img[sc], img.data(), img.size(), cudaMemcpyHostToDevice) != cudaSuccess ||
                    cudaMemcpy(h->sa3_bias[sc], bias.data(), bias.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess)
                    FAILC("fused sa3: weight image upload failed");
            }
            if (ok && (cudaFuncSetAttribute(safused::k_sa_fused<16, 256, 272, 128, 208, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)safused::Smem<272, 128, 208, 0>::kTotal) != cudaSuccess ||
                       cudaFuncSetAttribute(safused::k_sa_fused<32, 256, 272, 128, 208, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)safused::Smem<272, 128, 208, 0>::kTotal) != cudaSuccess))
                FAILC("fused sa3: cudaFuncSetAttribute failed");
        }
        if (h->fused_levels >= 2) {
            if (cudaFuncSetAttribute(safused::k_sa_fused<16, 96, 112, 64, 64, 128, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)safused::Smem<112, 64, 64, 128>::total(4)) != cudaSuccess ||
                cudaFuncSetAttribute(safused::k_sa_fused<32, 96, 112, 64, 96, 128, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)safused::Smem<112, 64, 96, 128>::total(4)) != cudaSuccess)
                FAILC("fused SA: cudaFuncSetAttribute failed");
        }
    }
#undef FAILC
#undef TRYC
    *out = h;
    return NIRRT_OK;
}

struct StageTimer {
    nirrt_pn2 *h; cudaStream_t s; int stage;
    StageTimer(nirrt_pn2 *h_, cudaStream_t s_, int st) : h(h_), s(s_), stage(st) { if (h->profiling) cudaEventRecord(h->ev[0], s); }
    ~StageTimer() {
        if (!h->profiling) return;
        cudaEventRecord(h->ev[1], s);
        cudaEventSynchronize(h->ev[1]);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
        h->stage_ms[stage] += ms;
    }
};

// development aid (NIRRT_PN2_TRACE=1): completion time of every launch of a forward on its stream, printed to stderr
static const bool g_trace = getenv("NIRRT_PN2_TRACE") && atoi(getenv("NIRRT_PN2_TRACE")) > 0;
static std::vector<std::pair<std::string, cudaEvent_t>> g_trace_ev;
static void trace_mark(const char *name, int a, int b, cudaStream_t s) {
    if (!g_trace) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, s);
    char buf[64];
    snprintf(buf, sizeof buf, "%s[%d,%d]", name, a, b);
    g_trace_ev.emplace_back(buf, e);
}
static void trace_dump() {
    if (!g_trace || g_trace_ev.empty()) return;
    cudaDeviceSynchronize();
    for (size_t i = 1; i < g_trace_ev.size(); i++) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, g_trace_ev[0].second, g_trace_ev[i].second);
        fprintf(stderr, "pn2-trace %-16s %8.1f us\n", g_trace_ev[i].first.c_str(), ms * 1e3f);
    }
    for (auto &p : g_trace_ev) cudaEventDestroy(p.second);
    g_trace_ev.clear();
}

// selections [it0, it1) of level l (it1 < 0: all of them)
static int fps_launch(nirrt_pn2 *h, int B, int l, const int *start, cudaStream_t s, int it0 = 0, int it1 = -1) {
    const int N = h->n[l - 1], np = h->n[l];
    if (it1 < 0) it1 = np;
    const size_t smem = (size_t)N * 3 * sizeof(float);
    static const int cfg = getenv("NIRRT_FPS_CFG") ? atoi(getenv("NIRRT_FPS_CFG")) : 0;      // development knob: threads per cloud
    if (h->fps_bucket && it0 == 0 && it1 == np && N > 512 && N <= 4096) {
        // spatially pruned selection (whole-level launches of the larger levels)
        const int nw = N > 2048 ? 8 : (N > 1024 ? 4 : 2);
        const size_t cap = (size_t)32 * nw * 16;
        const size_t sm = (size_t)N * 12 + (cap + (cap & 1)) * 2 + 513 * 4;
#define FPSB_ARGS h->xyz[l - 1], N, np, start, l - 1, h->fidx[l - 1], h->xyz[l]
        if (nw == 8) k_fps_bucket<8><<<B, 256, sm, s>>>(FPSB_ARGS);
        else if (nw == 4) k_fps_bucket<4><<<B, 128, sm, s>>>(FPSB_ARGS);
        else k_fps_bucket<2><<<B, 64, sm, s>>>(FPSB_ARGS);
#undef FPSB_ARGS
        PCUDA(cudaGetLastError());
        h->launches++;
        return NIRRT_OK;
    }
#define FPS_ARGS h->xyz[l - 1], N, np, start, l - 1, h->fidx[l - 1], h->xyz[l], it0, it1, h->fps_dist, h->fps_far
    // Few fat threads: 16 independent points per thread hide the arithmetic latency, and few warps keep the one barrier
    // per selection step short.
    if (N <= 64) k_fps<32, 2><<<B, 32, smem, s>>>(FPS_ARGS);
    else if (N <= 256) k_fps<32, 8><<<B, 32, smem, s>>>(FPS_ARGS);
    else if (N <= 1024) {
        if (cfg == 1) k_fps<128, 8><<<B, 128, smem, s>>>(FPS_ARGS);
        else if (cfg == 2) k_fps<256, 4><<<B, 256, smem, s>>>(FPS_ARGS);
        else k_fps<64, 16><<<B, 64, smem, s>>>(FPS_ARGS);
    } else if (N <= 2048) {
        if (cfg == 1) k_fps<256, 8><<<B, 256, smem, s>>>(FPS_ARGS);
        else if (cfg == 2) k_fps<512, 4><<<B, 512, smem, s>>>(FPS_ARGS);
        else k_fps<128, 16><<<B, 128, smem, s>>>(FPS_ARGS);
    } else k_fps<256, 16><<<B, 256, smem, s>>>(FPS_ARGS);
#undef FPS_ARGS
    PCUDA(cudaGetLastError());
    h->launches++;
    return NIRRT_OK;
}

// centroids [s0, s1) of level l (s1 < 0: all of them)
static int ball_query_launch(nirrt_pn2 *h, int B, int l, cudaStream_t s, int s0 = 0, int s1 = -1) {
    const int N = h->n[l - 1], S = h->n[l];
    if (s1 < 0) s1 = S;
    const int Sc = s1 - s0;
    const float r0 = (float)(kRad[l - 1][0] * kRad[l - 1][0]), r1 = (float)(kRad[l - 1][1] * kRad[l - 1][1]);
    if ((long long)B * S >= 65536) {
        const int bt = Sc >= 256 ? 256 : ((Sc + 31) / 32) * 32;
        // rounding slack: the float32 expansion -2 a.b + |a|^2 + |b|^2 is within ~1e-6 of the true squared distance for
        // coordinates in the unit ball, so a member is never farther than sqrt(r1^2 + 1e-5) from the centroid on any axis
        const float reach = sqrtf(r1 + 1e-5f) * 1.0001f + 1e-6f;
        if (h->bq_pruned && reach < 0.45f && N <= 32 * kSlabWordsMax) {
            // one CTA per cloud where possible: every CTA builds the cloud's bitmaps for itself
            const int pt = Sc >= 1024 ? 1024 : ((Sc + 31) / 32) * 32;
            k_ball_query<<<dim3((Sc + pt - 1) / pt, B), pt, (size_t)N * 16 + (size_t)3 * kSlabs * ((N + 31) / 32 + 1) * 4, s>>>(
                h->xyz[l - 1], N, h->xyz[l], S, r0, kK[0], r1, kK[1], h->grp[(l - 1) * 2], h->grp[(l - 1) * 2 + 1], s0, s1, reach);
        } else
            k_ball_query_bf<<<dim3((Sc + bt - 1) / bt, B), bt, (size_t)(N + 1) * 4 * sizeof(float), s>>>(
                h->xyz[l - 1], N, h->xyz[l], S, r0, kK[0], r1, kK[1], h->grp[(l - 1) * 2], h->grp[(l - 1) * 2 + 1], s0, s1);
    } else {
        k_ball_query_warp<<<dim3((Sc + 7) / 8, B), 256, (size_t)N * 4 * sizeof(float), s>>>(
            h->xyz[l - 1], N, h->xyz[l], S, r0, kK[0], r1, kK[1], h->grp[(l - 1) * 2], h->grp[(l - 1) * 2 + 1], s0, s1);
    }
    PCUDA(cudaGetLastError());
    h->launches++;
    return NIRRT_OK;
}

// feature propagation level f (fp4 .. fp1): mode 0 = 3-NN + interpolation, 1 = 3-NN only (coordinates),
// 2 = interpolation from the stored neighbours (PointNetFeaturePropagation.forward, pointnet2_utils.py:278-309)
static int interp_launch(nirrt_pn2 *h, int B, int f, int mode, const __half *upf, int upC, cudaStream_t s) {
    const int lo = 3 - f;
    const int N = h->n[lo], S = h->n[lo + 1];
    const int C1 = lo == 0 ? 0 : kC[lo];
    const int rows = B * N;
    const __half *f1 = C1 ? h->feat[lo] : nullptr;
#define INTERP_ARGS h->xyz[lo], N, h->xyz[lo + 1], S, f1, C1, upf, upC, B, h->bufA, h->knn_i[f], h->knn_w[f]
    if ((long long)B * N >= 65536) {
        const int bt = N >= 256 ? 256 : ((N + 31) / 32) * 32;
        const dim3 grid((N + bt - 1) / bt, B);
        const size_t smem = (size_t)S * sizeof(float4);
        if (mode == 0) k_interp<0><<<grid, bt, smem, s>>>(INTERP_ARGS);
        else if (mode == 1 && h->knn_pruned && S >= 512 && S <= 32 * kSlabWordsMax) {
            // the coarse level holds S FPS-spread points of the unit ball: the third-nearest is almost always within 0.3
            const float reach = h->knn_reach, ok_d = reach * reach - 1e-5f;
            const int pt = N >= 512 ? 512 : ((N + 31) / 32) * 32;
            k_knn_pruned<<<dim3((N + pt - 1) / pt, B), pt, (size_t)S * 16 + (size_t)3 * kSlabs * ((S + 31) / 32 + 1) * 4, s>>>(
                h->xyz[lo], N, h->xyz[lo + 1], S, h->knn_i[f], h->knn_w[f], reach, ok_d);
        } else if (mode == 1) k_interp<1><<<grid, bt, smem, s>>>(INTERP_ARGS);
        else k_interp<2><<<grid, bt, 0, s>>>(INTERP_ARGS);
    } else {
        if (mode == 0) k_interp_warp<0><<<(rows + 7) / 8, 256, 0, s>>>(INTERP_ARGS);
        else if (mode == 1) k_interp_warp<1><<<(rows + 7) / 8, 256, 0, s>>>(INTERP_ARGS);
        else k_interp_warp<2><<<(rows + 7) / 8, 256, 0, s>>>(INTERP_ARGS);
    }
#undef INTERP_ARGS
    PCUDA(cudaGetLastError());
    h->launches++;
    return NIRRT_OK;
}

// Selects the cloud size of the following classify calls (16 .. the n_points given at creation): level-0 buffers are
// sized for the creation value and every kernel takes the size as an argument, so one engine serves clouds of any
// size the reference's samplers produce (they only down-sample `if len(point_cloud) > n_points`).
extern "C" int nirrt_pn2_set_n_points(nirrt_pn2 *h, int n_points) {
    if (!h) return pfail(NIRRT_ERR_INVALID, "null handle");
    if (n_points < 16 || n_points > h->N0cap) return pfail(NIRRT_ERR_INVALID, "nirrt_pn2_set_n_points: 16 <= n_points <= creation size required");
    h->N0 = n_points; h->n[0] = n_points;
    return NIRRT_OK;
}

extern "C" int nirrt_pn2_classify_device(nirrt_pn2 *h, int batch, int dim, const float *pc, const float *start_mask,
                                         const float *goal_mask, const int32_t *fps_start, int64_t *path_pred,
                                         float *path_score, float *logp, void *stream) {
    if (!h || !pc || !start_mask || !goal_mask || !fps_start || !path_pred || !path_score)
        return pfail(NIRRT_ERR_INVALID, "nirrt_pn2_classify_device: null argument");
    if (batch < 1 || batch > h->maxB) return pfail(NIRRT_ERR_INVALID, "batch out of range (1..max_batch)");
    if (dim != 2 && dim != 3) return pfail(NIRRT_ERR_INVALID, "dim must be 2 or 3");
    PCUDA(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    const int B = batch;
    h->lastB = B;
    if (h->profiling) for (int i = 0; i < 8; i++) h->stage_ms[i] = 0.f;
    // g: the geometry stream (coordinates only).  With stage profiling on everything stays on s so that the
    // event brackets attribute serial times.
    const bool split = h->two_streams && !h->profiling;
    // Default: ONE geometry stream (FPS and ball query of a level alternate).  NIRRT_PN2_CHUNKS > 1 moves the ball queries
    // and 3-NN searches to a second stream and issues level 1 in chunks of centroids (FPS selects them in order, so ball
    // query, grouping and the first MLPs of the early centroids can run while the later ones are still being selected).
    // At 256 clouds the forward is throughput-bound (the kernels' serial times add up to the forward's), so the extra
    // overlap buys nothing there (measured 4.38 vs 4.41 ms); kept for small batches on an otherwise idle GPU.
    const int NC = (split && h->fused_levels >= 1) ? h->l1_chunks : 1;
    cudaStream_t g = split ? h->gs : s, g2 = (split && NC > 1) ? h->gs2 : g;
    trace_mark("start", 0, 0, s);
    if (split) { PCUDA(cudaEventRecord(h->ev_fork, s)); PCUDA(cudaStreamWaitEvent(g, h->ev_fork, 0)); }
    {
        StageTimer t(h, s, 0);
        k_prep<<<B, 256, (size_t)h->N0 * dim * sizeof(float), g>>>(pc, dim, start_mask, goal_mask, h->N0, h->xyz[0], h->in6);
        PCUDA(cudaGetLastError());
        h->launches++;
        trace_mark("prep", 0, 0, g);
    }
    if (split) {
        // the whole geometry is queued first.  g: the FPS chain of all four levels back to back (each level only needs
        // the previous level's centroids); g2: the ball queries behind their level's FPS, then the 3-NN searches
        const int Sc = h->n[1] / NC;
        for (int c = 0; c < NC; c++) {
            PTRY(fps_launch(h, B, 1, fps_start, g, c * Sc, (c + 1) * Sc));
            trace_mark("fps", 1, c, g);
            PCUDA(cudaEventRecord(h->ev_fc[c], g));
            PCUDA(cudaStreamWaitEvent(g2, h->ev_fc[c], 0));
            PTRY(ball_query_launch(h, B, 1, g2, c * Sc, (c + 1) * Sc));
            trace_mark("bq", 1, c, g2);
            PCUDA(cudaEventRecord(h->ev_bc[c], g2));
        }
        for (int l = 2; l <= 4; l++) {
            PTRY(fps_launch(h, B, l, fps_start, g));
            trace_mark("fps", l, 0, g);
            PCUDA(cudaEventRecord(h->ev_fl[l - 1], g));
            PCUDA(cudaStreamWaitEvent(g2, h->ev_fl[l - 1], 0));
            PTRY(ball_query_launch(h, B, l, g2));
            trace_mark("bq", l, 0, g2);
            PCUDA(cudaEventRecord(h->ev_bq[l - 1], g2));
        }
        for (int f = 0; f < 4; f++) {
            PTRY(interp_launch(h, B, f, 1, nullptr, 0, g2));
            trace_mark("knn", f, 0, g2);
            PCUDA(cudaEventRecord(h->ev_knn[f], g2));
        }
    }
    int li = 0;
    for (int l = 1; l <= 4; l++) {
        const int N = h->n[l - 1], S = h->n[l];
        if (!split) { StageTimer t(h, s, 1); PTRY(fps_launch(h, B, l, fps_start, s)); }
        if (!split) { StageTimer t(h, s, 2); PTRY(ball_query_launch(h, B, l, s)); }
        if (l <= h->fused_levels) {
            // gather + 3 layers + max-pool in one persistent tcgen05 kernel per radius (sa_fused.cuh); level 1 per chunk of
            // centroids as soon as that chunk's groups exist
            const int nc = l == 1 ? NC : 1;
            for (int c = 0; c < nc; c++) {
                if (split) PCUDA(cudaStreamWaitEvent(s, l == 1 ? h->ev_bc[c] : h->ev_bq[l - 1], 0));
                for (int sc = 0; sc < 2; sc++) {
                    StageTimer t(h, s, 4);
                    const int K = kK[sc];
                    safused::Args fa;
                    fa.in6 = h->in6; fa.feat = h->feat[l - 1]; fa.xyz = h->xyz[l - 1]; fa.new_xyz = h->xyz[l]; fa.gidx = h->grp[(l - 1) * 2 + sc];
                    fa.wimg = h->sa_img[l - 1][sc]; fa.bias = h->sa_bias[l - 1][sc];
                    fa.out = h->feat[l]; fa.N = N; fa.S = S; fa.B = B; fa.ldo = kC[l]; fa.col_off = sc == 0 ? 0 : h->conv[li + 2].N;
                    fa.tpc = S * K / 128; fa.chunk_tiles = fa.tpc / nc; fa.chunk_lo = c * fa.chunk_tiles;
                    if (!g_num_sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev); if (g_num_sms <= 0) g_num_sms = 148; }
                    const int ntiles = B * fa.chunk_tiles;
                    // sa1: 8 CTAs of one tile per SM (tensor memory is the limit); sa2: one CTA of four tiles per SM (its weight
                    // images are shared by the four groups)
                    const int units = l == 1 ? ntiles : (ntiles + 3) / 4;
                    const int cps = l == 1 ? 8 : 1;
                    const int grid = units < g_num_sms * cps ? units : g_num_sms * cps;
                    if (l == 1 && sc == 0) safused::k_sa_fused<16, 0, 16, 16, 16, 32, 1><<<grid, 128, safused::Smem<16, 16, 16, 32>::kTotal, s>>>(fa);
                    else if (l == 1) safused::k_sa_fused<32, 0, 16, 32, 32, 64, 1><<<grid, 128, safused::Smem<16, 32, 32, 64>::kTotal, s>>>(fa);
                    else if (sc == 0) safused::k_sa_fused<16, 96, 112, 64, 64, 128, 4><<<grid, 512, safused::Smem<112, 64, 64, 128>::total(4), s>>>(fa);
                    else safused::k_sa_fused<32, 96, 112, 64, 96, 128, 4><<<grid, 512, safused::Smem<112, 64, 96, 128>::total(4), s>>>(fa);
                    PCUDA(cudaGetLastError());
                    h->launches++;
                    trace_mark(sc ? "sa_fused1" : "sa_fused0", l, c, s);
                }
            }
            li += 6;
            continue;
        }
        if (split) PCUDA(cudaStreamWaitEvent(s, l == 1 ? h->ev_bc[NC - 1] : h->ev_bq[l - 1], 0));
        for (int sc = 0; sc < 2; sc++) {
            const int K = kK[sc];
            const int rows = B * S * K;
            const Conv &c0 = h->conv[li], &c1 = h->conv[li + 1], &c2 = h->conv[li + 2];
            if (l == 3 && h->sa3_img[sc]) {
                // gather + the first two layers in one kernel, then the last layer with the pooling epilogue
                StageTimer t(h, s, 4);
                safused::Args fa;
                fa.in6 = h->in6; fa.feat = h->feat[l - 1]; fa.xyz = h->xyz[l - 1]; fa.new_xyz = h->xyz[l]; fa.gidx = h->grp[(l - 1) * 2 + sc];
                fa.wimg = h->sa3_img[sc]; fa.bias = h->sa3_bias[sc];
                fa.out = h->bufA; fa.N = N; fa.S = S; fa.B = B; fa.ldo = c1.N; fa.col_off = 0;
                fa.tpc = S * K / 128; fa.chunk_tiles = fa.tpc; fa.chunk_lo = 0;
                if (!g_num_sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev); if (g_num_sms <= 0) g_num_sms = 148; }
                const int ntiles = B * fa.tpc;
                const int grid = ntiles < g_num_sms ? ntiles : g_num_sms;
                if (sc == 0) safused::k_sa_fused<16, 256, 272, 128, 208, 0, 1><<<grid, 128, safused::Smem<272, 128, 208, 0>::kTotal, s>>>(fa);
                else safused::k_sa_fused<32, 256, 272, 128, 208, 0, 1><<<grid, 128, safused::Smem<272, 128, 208, 0>::kTotal, s>>>(fa);
                PCUDA(cudaGetLastError());
                const int off = sc == 0 ? 0 : h->conv[li - 1].N;
                PTRY(launch_gemm(c2, h->bufA, rows, umma::MODE_POOL, h->feat[l], kC[l], off, K, s));
                h->launches += 2;
                trace_mark("sa3_fused", l, sc, s);
                li += 3;
                continue;
            }
            {
                StageTimer t(h, s, 3);
                if (l == 1) {
                    k_group_sa1<<<(rows + 255) / 256, 256, 0, s>>>(h->in6, h->xyz[0], N, h->xyz[1], S, h->grp[sc], K, B, h->bufA);
                } else {
                    const int CH = c0.K >> 3;                 // <= 66 chunks per row
                    const dim3 blk(CH, 256 / CH);
                    int s_shift = 0, k_shift = 0;
                    while ((1 << s_shift) < S) s_shift++;
                    while ((1 << k_shift) < K) k_shift++;
                    k_group<<<(rows + blk.y - 1) / blk.y, blk, 0, s>>>(h->feat[l - 1], kC[l - 1], h->xyz[l - 1], N, h->xyz[l], s_shift,
                                                                      h->grp[(l - 1) * 2 + sc], k_shift, (unsigned)rows, c0.K, h->bufA);
                }
                PCUDA(cudaGetLastError());
                h->launches++;
            }
            {
                StageTimer t(h, s, 4);
                PTRY(launch_gemm(c0, h->bufA, rows, umma::MODE_STORE, h->bufB, c0.N, 0, 0, s));
                PTRY(launch_gemm(c1, h->bufB, rows, umma::MODE_STORE, h->bufA, c1.N, 0, 0, s));
                // scale 0 channels first, then scale 1 (torch.cat(new_points_list, dim=1)); the last
                // layer widths (32|64, 128|128, 256|256, 512|512) need no padding
                const int off = sc == 0 ? 0 : h->conv[li - 1].N;
                PTRY(launch_gemm(c2, h->bufA, rows, umma::MODE_POOL, h->feat[l], kC[l], off, K, s));
                h->launches += 3;
                trace_mark("sa_gemm", l, sc, s);
            }
            li += 3;
        }
    }
    // feature propagation: fp4 (l3 <- l4), fp3, fp2, fp1 (l0 <- l1, no skip features)
    const __half *upf = h->feat[4];
    int upC = kC[4];
    for (int f = 0; f < 4; f++) {
        const int lo = 3 - f;
        const int rows = B * h->n[lo];
        if (f == 3 && h->fp_img) {
            // fp1 (3-NN interpolation + three 128-wide layers), conv1 and the head in one kernel (sa_fused.cuh)
            {
                StageTimer t(h, s, 5);
                if (split) PCUDA(cudaStreamWaitEvent(s, h->ev_knn[f], 0));
                else PTRY(interp_launch(h, B, f, 1, nullptr, 0, s));      // indices / weights only
            }
            StageTimer t(h, s, 6);
            safused::FpArgs fa;
            fa.feat2 = upf; fa.knn_i = h->knn_i[f]; fa.knn_w = h->knn_w[f]; fa.wimg = h->fp_img; fa.bias = h->fp_bias;
            fa.w2 = h->w2; fa.b2 = h->b2; fa.pred = (long long *)path_pred; fa.score = path_score; fa.logp = logp;
            fa.N = h->n[0]; fa.S = h->n[1]; fa.rows = (unsigned)rows;
            if (!g_num_sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev); if (g_num_sms <= 0) g_num_sms = 148; }
            const int pairs = (rows + 255) / 256;
            safused::k_fp1_fused<<<pairs < g_num_sms ? pairs : g_num_sms, 256, safused::kFpSmem, s>>>(fa);
            PCUDA(cudaGetLastError());
            h->launches++;
            trace_mark("fp1_fused", 0, 0, s);
            trace_dump();
            return NIRRT_OK;
        }
        {
            StageTimer t(h, s, 5);
            if (split) PCUDA(cudaStreamWaitEvent(s, h->ev_knn[f], 0));
            PTRY(interp_launch(h, B, f, split ? 2 : 0, upf, upC, s));
            trace_mark("interp", f, 0, s);
        }
        {
            StageTimer t(h, s, 6);
            const int nl = f == 3 ? 3 : 2;
            const __half *in = h->bufA;
            __half *pp[2] = {h->bufB, h->bufA};
            for (int j = 0; j < nl; j++, li++) {
                const Conv &c = h->conv[li];
                __half *o = (j == nl - 1) ? h->up[f] : pp[j & 1];
                PTRY(launch_gemm(c, in, rows, umma::MODE_STORE, o, c.N, 0, 0, s));
                h->launches++;
                in = o;
            }
        }
        trace_mark("fp_gemm", f, 0, s);
        upf = h->up[f];
        upC = kUpC[f];
    }
    {
        StageTimer t(h, s, 6);
        PTRY(launch_gemm(h->conv[33], h->up[3], B * h->N0, umma::MODE_STORE, h->bufA, 128, 0, 0, s));
        h->launches++;
    }
    {
        StageTimer t(h, s, 7);
        const size_t total = (size_t)B * h->N0;
        k_head<<<(unsigned)((total + 127) / 128), 128, 0, s>>>(h->bufA, h->w2, h->b2, total, (long long *)path_pred, path_score, logp);
        PCUDA(cudaGetLastError());
        h->launches++;
        trace_mark("head", 0, 0, s);
    }
    trace_dump();
    return NIRRT_OK;
}

extern "C" int nirrt_pn2_classify_sync(nirrt_pn2 *h, int batch, int dim, const float *pc, const float *start_mask,
                                       const float *goal_mask, const int32_t *fps_start, int64_t *path_pred,
                                       float *path_score, float *logp, void *stream) {
    if (!h || !pc || !start_mask || !goal_mask || !fps_start || !path_pred || !path_score)
        return pfail(NIRRT_ERR_INVALID, "nirrt_pn2_classify_sync: null argument");
    if (batch < 1 || batch > h->maxB) return pfail(NIRRT_ERR_INVALID, "batch out of range (1..max_batch)");
    if (dim != 2 && dim != 3) return pfail(NIRRT_ERR_INVALID, "dim must be 2 or 3");
    PCUDA(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    const size_t BN = (size_t)batch * h->N0;
    PCUDA(cudaMemcpyAsync(h->d_pc, pc, BN * dim * sizeof(float), cudaMemcpyHostToDevice, s));
    PCUDA(cudaMemcpyAsync(h->d_sm, start_mask, BN * sizeof(float), cudaMemcpyHostToDevice, s));
    PCUDA(cudaMemcpyAsync(h->d_gm, goal_mask, BN * sizeof(float), cudaMemcpyHostToDevice, s));
    PCUDA(cudaMemcpyAsync(h->d_fps, fps_start, (size_t)batch * 4 * sizeof(int), cudaMemcpyHostToDevice, s));
    PTRY(nirrt_pn2_classify_device(h, batch, dim, h->d_pc, h->d_sm, h->d_gm, h->d_fps, (int64_t *)h->d_pred, h->d_score,
                                   h->d_logp, s));
    PCUDA(cudaMemcpyAsync(path_pred, h->d_pred, BN * sizeof(long long), cudaMemcpyDeviceToHost, s));
    PCUDA(cudaMemcpyAsync(path_score, h->d_score, BN * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (logp) PCUDA(cudaMemcpyAsync(logp, h->d_logp, BN * 2 * sizeof(float), cudaMemcpyDeviceToHost, s));
    PCUDA(cudaStreamSynchronize(s));
    return NIRRT_OK;
}

extern "C" int64_t nirrt_pn2_read_buffer_sync(nirrt_pn2 *h, const char *name, void *out, int64_t bytes, void *stream) {
    if (!h || !name) return pfail(NIRRT_ERR_INVALID, "nirrt_pn2_read_buffer_sync: null argument");
    const size_t B = (size_t)(h->lastB > 0 ? h->lastB : 1);
    const void *src = nullptr;
    size_t size = 0;
    const std::string nm(name);
    if (nm == "xyz0") { src = h->xyz[0]; size = B * h->N0 * 3 * sizeof(float); }
    else if (nm.size() == 4 && nm.compare(0, 3, "fps") == 0 && nm[3] >= '0' && nm[3] <= '3') {
        const int l = nm[3] - '0'; src = h->fidx[l]; size = B * h->n[l + 1] * sizeof(int);
    } else if (nm.size() == 6 && nm.compare(0, 5, "group") == 0 && nm[5] >= '0' && nm[5] <= '7') {
        const int g = nm[5] - '0'; src = h->grp[g]; size = B * h->n[g / 2 + 1] * kK[g & 1] * sizeof(int);
    } else if (nm.size() == 5 && nm.compare(0, 4, "feat") == 0 && nm[4] >= '1' && nm[4] <= '4') {
        const int l = nm[4] - '0'; src = h->feat[l]; size = B * h->n[l] * kC[l] * sizeof(__half);
    } else if (nm.size() == 3 && nm.compare(0, 2, "up") == 0 && nm[2] >= '0' && nm[2] <= '3') {
        const int lo = nm[2] - '0'; const int f = 3 - lo; src = h->up[f]; size = B * h->n[lo] * kUpC[f] * sizeof(__half);
    } else return pfail(NIRRT_ERR_INVALID, "nirrt_pn2_read_buffer_sync: unknown buffer name");
    cudaStream_t s = (cudaStream_t)stream;
    PCUDA(cudaSetDevice(h->device));
    if (out && bytes > 0) {
        const size_t m = (size_t)bytes < size ? (size_t)bytes : size;
        PCUDA(cudaMemcpyAsync(out, src, m, cudaMemcpyDeviceToHost, s));
        PCUDA(cudaStreamSynchronize(s));
    }
    return (int64_t)size;
}

extern "C" int nirrt_pn2_set_profiling(nirrt_pn2 *h, int enabled) {
    if (!h) return pfail(NIRRT_ERR_INVALID, "null handle");
    h->profiling = enabled != 0;
    return NIRRT_OK;
}
extern "C" int nirrt_pn2_last_stage_ms(nirrt_pn2 *h, float *ms8) {
    if (!h || !ms8) return pfail(NIRRT_ERR_INVALID, "null argument");
    for (int i = 0; i < 8; i++) ms8[i] = h->stage_ms[i];
    return NIRRT_OK;
}
extern "C" int64_t nirrt_pn2_launch_count(nirrt_pn2 *h) { return h ? h->launches : 0; }

extern "C" int nirrt_gemm_f16_sync(const uint16_t *A, const uint16_t *W, const float *bias, int m, int n, int k, int mode,
                                   int group, uint16_t *out, void *stream) {
    if (!A || !W || !bias || !out) return pfail(NIRRT_ERR_INVALID, "nirrt_gemm_f16_sync: null argument");
    if (m < 1 || n < 16 || k < 16 || (n & 15) || (k & 15)) return pfail(NIRRT_ERR_INVALID, "n and k must be positive multiples of 16");
    if (mode != 0 && mode != 1) return pfail(NIRRT_ERR_INVALID, "mode must be 0 or 1");
    if (mode == 1 && ((group != 16 && group != 32) || m % group)) return pfail(NIRRT_ERR_INVALID, "group must be 16 or 32 and divide m");
    cudaStream_t s = (cudaStream_t)stream;
    __half *dA = nullptr, *dW = nullptr, *dO = nullptr;
    float *dB = nullptr;
    const size_t orow = mode == 0 ? (size_t)m : (size_t)(m / group);
    int rc = NIRRT_OK;
    auto cleanup = [&]() { cudaFree(dA); cudaFree(dW); cudaFree(dO); cudaFree(dB); };
#define GT(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cleanup(); return pfail(NIRRT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); } } while (0)
    GT(cudaMalloc(&dA, (size_t)m * k * 2)); GT(cudaMalloc(&dW, (size_t)n * k * 2)); GT(cudaMalloc(&dO, orow * n * 2)); GT(cudaMalloc(&dB, (size_t)n * 4));
    GT(cudaMemcpyAsync(dA, A, (size_t)m * k * 2, cudaMemcpyHostToDevice, s));
    GT(cudaMemcpyAsync(dW, W, (size_t)n * k * 2, cudaMemcpyHostToDevice, s));
    GT(cudaMemcpyAsync(dB, bias, (size_t)n * 4, cudaMemcpyHostToDevice, s));
    GT(cudaMemsetAsync(dO, 0, orow * n * 2, s));
    Conv c;
    c.K = k; c.N = n; c.BN = pick_bn(n); c.w = dW; c.b = dB;
    if (c.BN & 15) c.BN = 256;
    rc = make_tmap(&c.tmW, dW, (uint64_t)n, (uint64_t)k, (uint32_t)c.BN);
    if (!rc) rc = launch_gemm(c, dA, m, mode, dO, n, 0, group, s);
    if (rc) { cleanup(); return rc; }
    GT(cudaMemcpyAsync(out, dO, orow * n * 2, cudaMemcpyDeviceToHost, s));
    GT(cudaStreamSynchronize(s));
#undef GT
    cleanup();
    return NIRRT_OK;
}

// ------------------------------------------------------------------------------------------------
// Neural Connect graph analysis (wrapper/utils/bfs_connect_heuristic.py):
//   bfs_point_cloud            :32-78   r-disc graph over [src, dst, predicted path points], component
//                                       of src, has_path <=> dst is in it
//   get_boundary_mask          :5-29    visited path points with a non-path point closer than r
// float32 arithmetic as numpy evaluates it on float32 clouds: norm = sqrt((dx*dx + dy*dy) + dz*dz).
// One CTA: vertex coordinates in shared memory, level-synchronous frontier expansion (the visited
// SET of an exhausted BFS is order-independent; when dst is reached the reference's caller ignores
// the set), then the boundary test of the visited path points against all non-path points.
constexpr int kConnMax = 4096;         // points per cloud
__device__ __forceinline__ float norm_f32(float dx, float dy, float dz, int dim) {
    float s = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    if (dim == 3) s = __fadd_rn(s, __fmul_rn(dz, dz));
    return __fsqrt_rn(s);
}
// One CTA per analysis (blockIdx.x): clouds [n_clouds][n_max][dim] and masks [n_clouds][n_max], analysis b works on cloud
// b % n_clouds with n_pts[b] valid points (0: skipped); outputs [B][n_max], src / dst [B][3].
__global__ void __launch_bounds__(1024) k_connect(const float *pc_all, const int *n_pts, int n_max, int dim, const uint8_t *path_mask_all,
                                                  const float *src_all, const float *dst_all, float radius, int *has_path_all,
                                                  uint8_t *visited_all, uint8_t *boundary_all, int n_clouds) {
    const int b = blockIdx.x, n = n_pts[b];
    if (n == 0) { if (threadIdx.x == 0) has_path_all[b] = 0; return; }
    const float *pc = pc_all + (size_t)(b % n_clouds) * n_max * dim;
    const uint8_t *path_mask = path_mask_all + (size_t)(b % n_clouds) * n_max;
    const float *src = src_all + 3 * b, *dst = dst_all + 3 * b;
    int *has_path = has_path_all + b;
    uint8_t *visited_mask = visited_all + (size_t)b * n_max, *boundary_mask = boundary_all + (size_t)b * n_max;
    extern __shared__ unsigned char s_raw[];
    float *vx = reinterpret_cast<float *>(s_raw);            // [m] vertices: 0 = src, 1 = dst, 2.. = path points
    float *vy = vx + (kConnMax + 2), *vz = vy + (kConnMax + 2);
    int *vid = reinterpret_cast<int *>(vz + (kConnMax + 2)); // point index of vertex (>= 2)
    int *fr = vid + (kConnMax + 2);                          // frontier lists, double buffered
    int *fr2 = fr + (kConnMax + 2);
    unsigned char *state = reinterpret_cast<unsigned char *>(fr2 + (kConnMax + 2));   // 0 unvisited, 1 visited
    __shared__ int s_m, s_cnt[2], s_found;
    const int tid = threadIdx.x;
    if (tid == 0) { s_m = 2; s_cnt[0] = 1; s_cnt[1] = 0; s_found = 0; }
    __syncthreads();
    // compact the path points in ascending point order (order only matters for determinism of vid)
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + tid;
        const bool p = i < n && path_mask[i];
        // block-ordered append via warp ballots
        const unsigned bal = __ballot_sync(0xffffffffu, p);
        __shared__ int s_w[32];
        const int w = tid >> 5, l = tid & 31;
        if (l == 0) s_w[w] = __popc(bal);
        __syncthreads();
        int off = s_m;
        for (int q = 0; q < w; q++) off += s_w[q];
        if (p) {
            const int pos = off + __popc(bal & ((1u << l) - 1u));
            vx[pos] = pc[(size_t)i * dim]; vy[pos] = pc[(size_t)i * dim + 1]; vz[pos] = dim == 3 ? pc[(size_t)i * dim + 2] : 0.f;
            vid[pos] = i;
        }
        __syncthreads();
        if (tid == 0) { int t = 0; for (int q = 0; q < (int)(blockDim.x >> 5); q++) t += s_w[q]; s_m += t; }
        __syncthreads();
    }
    const int m = s_m;
    if (tid == 0) {
        vx[0] = src[0]; vy[0] = src[1]; vz[0] = dim == 3 ? src[2] : 0.f;
        vx[1] = dst[0]; vy[1] = dst[1]; vz[1] = dim == 3 ? dst[2] : 0.f;
        fr[0] = 0;
    }
    for (int j = tid; j < m; j += blockDim.x) state[j] = j == 0 ? 1 : 0;
    __syncthreads();
    int *cur = fr, *nxt = fr2;
    int level = 0;
    while (true) {
        const int nf = s_cnt[level & 1];
        if (nf == 0 || s_found) break;
        for (int j = tid; j < m; j += blockDim.x) {
            if (state[j]) continue;
            bool hit = false;
            for (int k = 0; k < nf && !hit; k++) {
                const int f = cur[k];
                hit = norm_f32(__fsub_rn(vx[j], vx[f]), __fsub_rn(vy[j], vy[f]), __fsub_rn(vz[j], vz[f]), dim) < radius;
            }
            if (hit) {
                state[j] = 1;
                if (j == 1) s_found = 1;
                else nxt[atomicAdd(&s_cnt[(level + 1) & 1], 1)] = j;
            }
        }
        __syncthreads();
        if (tid == 0) s_cnt[level & 1] = 0;
        __syncthreads();
        int *t = cur; cur = nxt; nxt = t;
        level++;
    }
    __syncthreads();
    if (tid == 0) *has_path = s_found;
    for (int i = tid; i < n; i += blockDim.x) { visited_mask[i] = 0; boundary_mask[i] = 0; }
    __syncthreads();
    for (int j = 2 + tid; j < m; j += blockDim.x) {
        if (!state[j]) continue;
        const int i = vid[j];
        visited_mask[i] = 1;
        bool b = false;
        for (int u = 0; u < n && !b; u++) {
            if (path_mask[u]) continue;
            const float ux = pc[(size_t)u * dim], uy = pc[(size_t)u * dim + 1], uz = dim == 3 ? pc[(size_t)u * dim + 2] : 0.f;
            b = norm_f32(__fsub_rn(vx[j], ux), __fsub_rn(vy[j], uy), __fsub_rn(vz[j], uz), dim) < radius;
        }
        boundary_mask[i] = b ? 1 : 0;
    }
}

// Grow-only device workspace of the connect analyses (one per process; the entry points are not re-entrant)
struct ConnWs { void *p = nullptr; size_t bytes = 0; };
static ConnWs g_conn_ws;

extern "C" int nirrt_connect_analyse_batch_sync(const float *pc, const int *n_pts, int n_max, int dim, int batch, const uint8_t *path_mask,
                                                const float *src, const float *dst, float radius, int *has_path,
                                                uint8_t *visited_mask, uint8_t *boundary_mask, void *stream) {
    if (!pc || !n_pts || !path_mask || !src || !dst || !has_path || !visited_mask || !boundary_mask)
        return pfail(NIRRT_ERR_INVALID, "nirrt_connect_analyse_batch_sync: null argument");
    if (batch < 1 || n_max < 1 || n_max > kConnMax || (dim != 2 && dim != 3))
        return pfail(NIRRT_ERR_INVALID, "nirrt_connect_analyse_batch_sync: batch >= 1, 1 <= n_max <= 4096, dim 2 or 3");
    for (int b = 0; b < batch; b++)
        if (n_pts[b] < 1 || n_pts[b] > n_max) return pfail(NIRRT_ERR_INVALID, "nirrt_connect_analyse_batch_sync: 1 <= n_pts[b] <= n_max");
    if (nirrt_device_count() <= 0) return pfail(NIRRT_ERR_NO_DEVICE, "no sm_100 device");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t B = (size_t)batch, N = (size_t)n_max;
    const size_t o_pc = 0, o_sd = o_pc + sizeof(float) * B * N * dim, o_n = o_sd + sizeof(float) * B * 6, o_hp = o_n + sizeof(int) * B,
                 o_mask = o_hp + sizeof(int) * B, o_out = o_mask + B * N, total = o_out + 2 * B * N;
    if (g_conn_ws.bytes < total) {
        if (g_conn_ws.p) cudaFree(g_conn_ws.p);
        g_conn_ws.p = nullptr; g_conn_ws.bytes = 0;
        PCUDA(cudaMalloc(&g_conn_ws.p, total + (total >> 2)));
        g_conn_ws.bytes = total + (total >> 2);
    }
    unsigned char *w = (unsigned char *)g_conn_ws.p;
    float *d_pc = (float *)(w + o_pc), *d_sd = (float *)(w + o_sd);
    int *d_n = (int *)(w + o_n), *d_hp = (int *)(w + o_hp);
    uint8_t *d_mask = w + o_mask, *d_out = w + o_out;
    PCUDA(cudaMemcpyAsync(d_pc, pc, sizeof(float) * B * N * dim, cudaMemcpyHostToDevice, s));
    PCUDA(cudaMemcpyAsync(d_sd, src, sizeof(float) * B * 3, cudaMemcpyHostToDevice, s));
    PCUDA(cudaMemcpyAsync(d_sd + 3 * B, dst, sizeof(float) * B * 3, cudaMemcpyHostToDevice, s));
    PCUDA(cudaMemcpyAsync(d_n, n_pts, sizeof(int) * B, cudaMemcpyHostToDevice, s));
    PCUDA(cudaMemcpyAsync(d_mask, path_mask, B * N, cudaMemcpyHostToDevice, s));
    const size_t smem = (size_t)(kConnMax + 2) * (3 * sizeof(float) + 3 * sizeof(int) + 1) + 16;
    static bool attr = false;
    if (!attr) { PCUDA(cudaFuncSetAttribute(k_connect, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
    k_connect<<<batch, 1024, smem, s>>>(d_pc, d_n, n_max, dim, d_mask, d_sd, d_sd + 3 * B, radius, d_hp, d_out, d_out + B * N, batch);
    PCUDA(cudaGetLastError());
    PCUDA(cudaMemcpyAsync(has_path, d_hp, sizeof(int) * B, cudaMemcpyDeviceToHost, s));
    PCUDA(cudaMemcpyAsync(visited_mask, d_out, B * N, cudaMemcpyDeviceToHost, s));
    PCUDA(cudaMemcpyAsync(boundary_mask, d_out + B * N, B * N, cudaMemcpyDeviceToHost, s));
    PCUDA(cudaStreamSynchronize(s));
    return NIRRT_OK;
}

// ---- one Neural Connect trial for a batch of problems, device resident (generate_connected_path_points,
// pointnet2_wrapper_connect_bfs.py:76-240, the part after the network call): path_mask |= prediction, both searches of
// every active problem, and for the problems still unconnected the boundary-point heuristic and the new start / goal
// neighbourhood masks (select_heuristic_boundary_point, bfs_connect_heuristic.py:142-181; get_point_cloud_mask_around_points).
__global__ void __launch_bounds__(256) k_connect_merge(const uint8_t *active, const long long *pred, uint8_t *path_mask, int n, int B, int *n_act) {
    const int b = blockIdx.x;
    const bool on = active[b] != 0;
    if (threadIdx.x == 0) { n_act[b] = on ? n : 0; n_act[b + B] = on ? n : 0; }
    if (!on) return;
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        if (pred[(size_t)b * n + i] != 0) path_mask[(size_t)b * n + i] = 1;
}
// search j of 2B (j < B: start -> goal of problem j; j >= B: goal -> start of problem j - B).  The ranks of the reference
// come from np.argsort, whose order among EQUAL keys is an implementation detail: a search with equal keys is reported in
// ties[] and left untouched for the caller to decide with numpy itself.
__global__ void __launch_bounds__(256) k_connect_pick(const float *pc_all, int n, int dim, int B, const int *n_act, const int *hp,
                                                      const uint8_t *bnd_all, const float *src_all, const float *dst_all, float radius,
                                                      float *start_mask, float *goal_mask, int *ties) {
    const int j = blockIdx.x, b = j % B, dir = j / B, tid = threadIdx.x;
    if (tid == 0) ties[j] = 0;
    if (n_act[j] == 0 || hp[b] || hp[b + B]) return;
    const float *pc = pc_all + (size_t)b * n * dim;
    const uint8_t *bnd = bnd_all + (size_t)j * n;
    const float *src = src_all + 3 * j, *dst = dst_all + 3 * j;
    extern __shared__ unsigned char s_pick[];
    int *idx = reinterpret_cast<int *>(s_pick);
    float *fs = reinterpret_cast<float *>(idx + n), *tot = fs + n;
    __shared__ int s_m, s_w[8], s_tie;
    __shared__ unsigned s_best;
    if (tid == 0) { s_m = 0; s_tie = 0; s_best = 0xffffffffu; }
    __syncthreads();
    for (int base = 0; base < n; base += blockDim.x) {       // boundary points in index order
        const int i = base + tid;
        const bool p = i < n && bnd[i];
        const unsigned bal = __ballot_sync(0xffffffffu, p);
        const int w = tid >> 5, l = tid & 31;
        if (l == 0) s_w[w] = __popc(bal);
        __syncthreads();
        int off = s_m;
        for (int q = 0; q < w; q++) off += s_w[q];
        if (p) idx[off + __popc(bal & ((1u << l) - 1u))] = i;
        __syncthreads();
        if (tid == 0) { int t = 0; for (int q = 0; q < 8; q++) t += s_w[q]; s_m += t; }
        __syncthreads();
    }
    const int m = s_m;
    if (m == 0) return;                                      // no boundary point: the mask stays
    for (int i = tid; i < m; i += blockDim.x) {
        const float *p = pc + (size_t)idx[i] * dim;
        const float pz = dim == 3 ? p[2] : 0.f;
        const float a = norm_f32(__fsub_rn(p[0], src[0]), __fsub_rn(p[1], src[1]), __fsub_rn(pz, dim == 3 ? src[2] : 0.f), dim);
        const float g = norm_f32(__fsub_rn(p[0], dst[0]), __fsub_rn(p[1], dst[1]), __fsub_rn(pz, dim == 3 ? dst[2] : 0.f), dim);
        fs[i] = a; tot[i] = __fadd_rn(a, g);
    }
    __syncthreads();
    // rank of the total cost ascending + rank of the cost from the start descending; lowest sum, first boundary point on ties
    unsigned best = 0xffffffffu;
    bool tie = false;
    for (int i = tid; i < m; i += blockDim.x) {
        const float ti = tot[i], fi = fs[i];
        int score = 0;
        for (int k = 0; k < m; k++) {
            const float tk = tot[k], fk = fs[k];
            score += (tk < ti ? 1 : 0) + (fk > fi ? 1 : 0);
            tie = tie || (k != i && (tk == ti || fk == fi));
        }
        best = min(best, ((unsigned)score << 12) | (unsigned)i);
    }
    if (tie) s_tie = 1;
    atomicMin(&s_best, best);
    __syncthreads();
    if (s_tie) { if (tid == 0) ties[j] = 1; return; }
    const float *q = pc + (size_t)idx[s_best & 4095u] * dim;
    const float qx = q[0], qy = q[1], qz = dim == 3 ? q[2] : 0.f;
    float *out = (dir == 0 ? start_mask : goal_mask) + (size_t)b * n;
    for (int i = tid; i < n; i += blockDim.x) {
        const float *p = pc + (size_t)i * dim;
        out[i] = norm_f32(__fsub_rn(p[0], qx), __fsub_rn(p[1], qy), __fsub_rn(dim == 3 ? p[2] : 0.f, qz), dim) < radius ? 1.f : 0.f;
    }
}

// get_point_cloud_mask_around_points(pc, x, radius) of the float32 cloud for x = the problem's start and goal: the masks the
// first trial's network call sees (pointnet2_wrapper_connect_bfs.py works on pc.astype(float32))
__global__ void __launch_bounds__(256) k_connect_masks(const float *pc_all, int n, int dim, int B, const float *src, float radius,
                                                       float *start_mask, float *goal_mask) {
    const int b = blockIdx.x;
    const float *pc = pc_all + (size_t)b * n * dim, *a = src + 3 * b, *g = src + 3 * (B + b);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float *p = pc + (size_t)i * dim;
        const float pz = dim == 3 ? p[2] : 0.f;
        start_mask[(size_t)b * n + i] = norm_f32(__fsub_rn(p[0], a[0]), __fsub_rn(p[1], a[1]), __fsub_rn(pz, dim == 3 ? a[2] : 0.f), dim) < radius ? 1.f : 0.f;
        goal_mask[(size_t)b * n + i] = norm_f32(__fsub_rn(p[0], g[0]), __fsub_rn(p[1], g[1]), __fsub_rn(pz, dim == 3 ? g[2] : 0.f), dim) < radius ? 1.f : 0.f;
    }
}
extern "C" int nirrt_connect_masks_device(const float *pc, int n_points, int dim, int batch, const float *src, float radius,
                                          float *start_mask, float *goal_mask, void *stream) {
    if (!pc || !src || !start_mask || !goal_mask) return pfail(NIRRT_ERR_INVALID, "nirrt_connect_masks_device: null argument");
    if (batch < 1 || n_points < 1 || (dim != 2 && dim != 3)) return pfail(NIRRT_ERR_INVALID, "nirrt_connect_masks_device: bad argument");
    if (nirrt_device_count() <= 0) return pfail(NIRRT_ERR_NO_DEVICE, "no sm_100 device");
    k_connect_masks<<<batch, 256, 0, (cudaStream_t)stream>>>(pc, n_points, dim, batch, src, radius, start_mask, goal_mask);
    PCUDA(cudaGetLastError());
    return NIRRT_OK;
}

extern "C" int nirrt_connect_trial_device(const float *pc, int n_points, int dim, int batch, const uint8_t *active, uint8_t *path_mask,
                                          const int64_t *pred, const float *src, const float *dst, float radius, float *start_mask,
                                          float *goal_mask, int32_t *has_path, int32_t *ties, uint8_t *tie_boundary, void *stream) {
    if (!pc || !active || !path_mask || !pred || !src || !dst || !start_mask || !goal_mask || !has_path || !ties || !tie_boundary)
        return pfail(NIRRT_ERR_INVALID, "nirrt_connect_trial_device: null argument");
    if (batch < 1 || n_points < 1 || n_points > 4095 || (dim != 2 && dim != 3))
        return pfail(NIRRT_ERR_INVALID, "nirrt_connect_trial_device: batch >= 1, 1 <= n_points <= 4095, dim 2 or 3");
    if (nirrt_device_count() <= 0) return pfail(NIRRT_ERR_NO_DEVICE, "no sm_100 device");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t B = (size_t)batch, N = (size_t)n_points;
    const size_t o_n = 0, o_hp = o_n + sizeof(int) * 2 * B, o_tie = o_hp + sizeof(int) * 2 * B, o_out = o_tie + sizeof(int) * 2 * B,
                 total = o_out + 4 * B * N;
    if (g_conn_ws.bytes < total) {
        if (g_conn_ws.p) cudaFree(g_conn_ws.p);
        g_conn_ws.p = nullptr; g_conn_ws.bytes = 0;
        PCUDA(cudaMalloc(&g_conn_ws.p, total + (total >> 2)));
        g_conn_ws.bytes = total + (total >> 2);
    }
    unsigned char *w = (unsigned char *)g_conn_ws.p;
    int *d_n = (int *)(w + o_n), *d_hp = (int *)(w + o_hp), *d_tie = (int *)(w + o_tie);
    uint8_t *d_vis = w + o_out, *d_bnd = d_vis + 2 * B * N;
    k_connect_merge<<<batch, 256, 0, s>>>(active, (const long long *)pred, path_mask, n_points, batch, d_n);
    PCUDA(cudaGetLastError());
    const size_t smem = (size_t)(kConnMax + 2) * (3 * sizeof(float) + 3 * sizeof(int) + 1) + 16;
    static bool attr = false;
    if (!attr) { PCUDA(cudaFuncSetAttribute(k_connect, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
    k_connect<<<2 * batch, 1024, smem, s>>>(pc, d_n, n_points, dim, path_mask, src, dst, radius, d_hp, d_vis, d_bnd, batch);
    PCUDA(cudaGetLastError());
    k_connect_pick<<<2 * batch, 256, N * 12, s>>>(pc, n_points, dim, batch, d_n, d_hp, d_bnd, src, dst, radius, start_mask, goal_mask, d_tie);
    PCUDA(cudaGetLastError());
    PCUDA(cudaMemcpyAsync(has_path, d_hp, sizeof(int) * 2 * B, cudaMemcpyDeviceToHost, s));
    PCUDA(cudaMemcpyAsync(ties, d_tie, sizeof(int) * 2 * B, cudaMemcpyDeviceToHost, s));
    PCUDA(cudaStreamSynchronize(s));
    bool any = false;
    for (size_t j = 0; j < 2 * B; j++)
        if (ties[j]) { any = true; PCUDA(cudaMemcpyAsync(tie_boundary + j * N, d_bnd + j * N, N, cudaMemcpyDeviceToHost, s)); }
    if (any) PCUDA(cudaStreamSynchronize(s));
    return NIRRT_OK;
}

extern "C" int nirrt_connect_analyse_sync(const float *pc, int n, int dim, const uint8_t *path_mask, const float *src,
                                          const float *dst, float radius, int *has_path, uint8_t *visited_mask,
                                          uint8_t *boundary_mask, void *stream) {
    if (!src || !dst) return pfail(NIRRT_ERR_INVALID, "nirrt_connect_analyse_sync: null argument");
    float s3[3] = {src[0], src[1], dim == 3 ? src[2] : 0.f}, d3[3] = {dst[0], dst[1], dim == 3 ? dst[2] : 0.f};
    return nirrt_connect_analyse_batch_sync(pc, &n, n, dim, 1, path_mask, s3, d3, radius, has_path, visited_mask, boundary_mask, stream);
}
