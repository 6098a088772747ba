// sa_fused.cuh -- a set-abstraction scale of PointNet++ as ONE persistent kernel:
// grouping (gather + centroid subtraction), the three 1x1 convolutions (+ folded BatchNorm + ReLU)
// and the max over the K neighbours (pointnet2_utils.py:243-259), for 128 (centroid, neighbour) rows
// per tile.  The tile's activations never leave the SM: the gathered operand and every intermediate
// layer live in one shared-memory tile in the tcgen05 K-major / 128-byte-swizzle layout (one 16 KB
// block per 64 channels), the accumulators in TMEM; the weights of all three layers are copied into
// shared memory once per CTA.  Used for sa1 (6 input channels) and sa2 (96), for sa3 up to its second
// layer (N3 = 0), and -- k_fp1_fused below -- for fp1 + conv1 + the classifier head; the remaining
// layers keep their weights in L2 and go through umma::k_gemm.
//
//   Groups of 128 threads (NG per CTA), thread = row = TMEM lane of its group's tile.  Per layer: all threads write their row of the operand,
//   fence.proxy.async + __syncthreads, thread 0 issues the tcgen05.mma instructions (K / 16 of them)
//   and commits to an mbarrier, all threads wait, tcgen05.ld their accumulator row.  Several CTAs per
//   SM hide each other's round trips.
//
// Operand rows (coordinates travel as fp16 hi + lo pairs, the weight columns are repeated):
//   sa1   [x y z start goal free | rx ry rz | x_lo y_lo z_lo | rx_lo ry_lo rz_lo | 0]          K0 = 16
//   sa2   [96 features | rx ry rz rx_lo ry_lo rz_lo 0 0 | 0 x 8]                               K0 = 112
#pragma once
#include "umma_gemm.cuh"

namespace safused {

using umma::smem_u32;

// byte offset of 16-byte chunk `c` (0..7) of row `r` in a K-major SWIZZLE_128B block (rows of 128
// bytes, 8-row groups of 1024 bytes, chunk index XOR (r mod 8))
__host__ __device__ __forceinline__ int sw128_off(int r, int c) { return (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4); }
// operand tile: logical chunk cc of row r (block cc / 8 of 16 KB)
__device__ __forceinline__ int a_off(int r, int cc) { return (cc >> 3) * 16384 + sw128_off(r, cc & 7); }
// weight image of a layer with n rows: block kb is n * 128 bytes
__host__ __device__ __forceinline__ size_t w_off(int n_rows, int n, int cc) { return (size_t)(cc >> 3) * n_rows * 128 + sw128_off(n, cc & 7); }
__host__ __device__ constexpr int nblk(int k) { return (k + 63) / 64; }
__host__ __device__ constexpr int imax(int a, int b) { return a > b ? a : b; }

struct Args {
    const float *in6;       // sa1: [B][N][6] network input
    const __half *feat;     // sa2: [B][N][C] features of the previous level
    const float *xyz;       // [B][N][3] coordinates of the previous level (sa2)
    const float *new_xyz;   // [B][S][3] centroids
    const int *gidx;        // [B][S][G]
    const uint8_t *wimg;    // W1 | W2 | W3 shared-memory images
    const float *bias;      // [N1 + N2 + N3]
    __half *out;            // [B][S][ldo]
    int N, S, B, ldo, col_off;
    // tiles of 128 rows: tpc per cloud (S * G / 128); this launch covers tiles [chunk_lo, chunk_lo + chunk_tiles) of every
    // cloud (a range of centroids -- level 1 is launched per chunk of FPS selections)
    int tpc, chunk_lo, chunk_tiles;
};

__device__ __forceinline__ void split_hi_lo(float x, __half &hi, __half &lo) {
    hi = __float2half_rn(x);
    lo = __float2half_rn(__fsub_rn(x, __half2float(hi)));
}

// accumulator row -> relu(. + bias) -> fp16 -> this thread's row of the next operand tile
template <int NCOLS>
__device__ __forceinline__ void epilogue_to_operand(uint32_t taddr, uint8_t *sA, int r, const float *bias) {
    constexpr int W = NCOLS % 32 == 0 ? 32 : 16;          // accumulator columns per TMEM round trip
#pragma unroll 1
    for (int h = 0; h < NCOLS; h += W) {
        uint32_t u[W];
        if (W == 32) umma::tmem_ld32(taddr + h, u);
        else umma::tmem_ld16(taddr + h, u);
#pragma unroll
        for (int q = 0; q < W; q += 8) {
            // the biases of 8 columns as two 16-byte shared-memory loads (layer offsets are multiples of 16 floats)
            const float4 b0 = *reinterpret_cast<const float4 *>(bias + h + q), b1 = *reinterpret_cast<const float4 *>(bias + h + q + 4);
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            uint32_t p[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const float a = fmaxf(__uint_as_float(u[q + 2 * i]) + bb[2 * i], 0.f);
                const float b = fmaxf(__uint_as_float(u[q + 2 * i + 1]) + bb[2 * i + 1], 0.f);
                p[i] = umma::pack_half2_sat(a, b);
            }
            *reinterpret_cast<uint4 *>(sA + a_off(r, (h + q) >> 3)) = make_uint4(p[0], p[1], p[2], p[3]);
        }
    }
}

template <int K, int NROWS>
__device__ __forceinline__ void issue(uint32_t d_tmem, uint32_t sa, uint32_t sw, uint64_t *bar) {
    const uint32_t idesc = umma::make_idesc(NROWS);
#pragma unroll
    for (int kb = 0; kb < nblk(K); kb++) {
        const int ksteps = (K - 64 * kb < 64 ? K - 64 * kb : 64) / 16;
#pragma unroll
        for (int k = 0; k < ksteps; k++)
            umma::mma_f16(d_tmem, umma::make_smem_desc(sa + kb * 16384 + k * 32), umma::make_smem_desc(sw + kb * NROWS * 128 + k * 32),
                          idesc, (uint32_t)((kb | k) != 0));
    }
    umma::mma_commit(bar);
}

template <int K0, int N1, int N2, int N3>
struct Smem {
    static constexpr int kABytes = imax(imax(nblk(K0), nblk(N1)), nblk(N2)) * 16384;
    static constexpr int kW1 = N1 * nblk(K0) * 128, kW2 = N2 * nblk(N1) * 128, kW3 = N3 * nblk(N2) * 128;
    static constexpr int kBias = (N1 + N2 + N3 + 2) * 4;
    static constexpr size_t kTotal = 1024 + (size_t)kABytes + kW1 + kW2 + kW3 + kBias + 64;
    static constexpr size_t total(int ng) { return kTotal + (size_t)(ng - 1) * kABytes; }      // ng groups per CTA
};

// C == 0: sa1 (in6 input); C > 0: features [.,C] of the previous level + relative coordinates
// NG: independent groups of 128 threads per CTA (thread = row = TMEM lane of its group's tile).  The groups share the weight
// images and each owns an operand tile, an mbarrier and TCOLS accumulator columns: where the weights are what limits the number
// of CTAs per SM (sa2: 40-60 KB of weights next to a 32 KB tile), NG = 4 keeps four tiles in flight per SM instead of two or
// three, so that one group's gather and epilogues overlap the others' round trips.
// N3 == 0: two layers only -- the fp16 activations of the second layer are written to a.out as rows [row][N2] (ld a.ldo) for a
// following umma::k_gemm with the pooling epilogue.  Used for sa3, whose first two weight matrices (80 + 52 KB) fit next to the
// 80 KB operand tile while the third does not: the gathered operand and the first activation never touch HBM.
template <int G, int C, int K0, int N1, int N2, int N3, int NG>
__global__ void __launch_bounds__(128 * NG, (N3 == 0 || NG > 1) ? 1 : 8) k_sa_fused(const Args a) {
    typedef Smem<K0, N1, N2, N3> SM;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
    const int grp = threadIdx.x >> 7, tid = threadIdx.x & 127, warp = tid >> 5, lane = tid & 31;      // within the group
    uint8_t *sA = smem + grp * SM::kABytes;               // this group's operand tile (also its pooling scratch)
    uint8_t *sW1 = smem + NG * SM::kABytes, *sW2 = sW1 + SM::kW1, *sW3 = sW2 + SM::kW2;
    float *s_bias = reinterpret_cast<float *>(sW3 + SM::kW3);
    uint64_t *bar0 = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(s_bias) + SM::kBias);
    uint64_t *bar = bar0 + grp;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bar0 + NG);
    constexpr int NMAX = imax(imax(N1, N2), N3);
    constexpr int TCOLS = NMAX <= 32 ? 32 : (NMAX <= 64 ? 64 : (NMAX <= 128 ? 128 : 256));
    static_assert(TCOLS * NG <= 512, "tensor memory holds 512 columns");

    if (threadIdx.x == 0) {
        for (int g = 0; g < NG; g++) umma::mbar_init(bar0 + g, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(TCOLS * NG) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(a.wimg);
        uint4 *dst = reinterpret_cast<uint4 *>(sW1);
        for (int i = threadIdx.x; i < (SM::kW1 + SM::kW2 + SM::kW3) / 16; i += 128 * NG) dst[i] = src[i];
        for (int i = threadIdx.x; i < N1 + N2 + N3; i += 128 * NG) s_bias[i] = a.bias[i];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_all = *tmem_ptr;
    const uint32_t tmem_base = tmem_all + (uint32_t)(grp * TCOLS);
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const uint32_t sa_u = smem_u32(sA), sw1_u = smem_u32(sW1), sw2_u = smem_u32(sW2), sw3_u = smem_u32(sW3);
    uint32_t phase = 0;
    const int nvt = a.B * a.chunk_tiles;
    auto group_sync = [&]() {
        if (NG == 1) __syncthreads();
        else asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
    };

#define SA_LAYER_SYNC()                                                   \
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          \
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");      \
    group_sync();
#define SA_WAIT()                                                         \
    umma::mbar_wait(bar, phase); phase ^= 1u;                             \
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    for (int vt = blockIdx.x * NG + grp; vt < nvt; vt += gridDim.x * NG) {
        const int tile = (vt / a.chunk_tiles) * a.tpc + a.chunk_lo + vt % a.chunk_tiles;
        // ---- grouping: this thread's row (pointnet2_utils.py:246-253)
        {
            // 32-bit index arithmetic (rows < 2^31 is checked at launch): a 64-bit division per row costs more than the gather
            const unsigned r = (unsigned)tile * 128u + (unsigned)tid;
            const unsigned bs = r / G;
            const int b = tile / a.tpc;
            const int i = a.gidx[r];
            const float *c = a.new_xyz + (size_t)bs * 3;
            if (C == 0) {
                const float *f = a.in6 + ((size_t)b * a.N + i) * 6;
                __half h[16];
                float v[9];
#pragma unroll
                for (int k = 0; k < 6; k++) v[k] = f[k];
#pragma unroll
                for (int k = 0; k < 3; k++) v[6 + k] = __fsub_rn(f[k], c[k]);
#pragma unroll
                for (int k = 0; k < 3; k++) split_hi_lo(v[k], h[k], h[9 + k]);
#pragma unroll
                for (int k = 3; k < 6; k++) h[k] = __float2half_rn(v[k]);
#pragma unroll
                for (int k = 6; k < 9; k++) split_hi_lo(v[k], h[k], h[6 + k]);
                h[15] = __float2half_rn(0.f);
                *reinterpret_cast<uint4 *>(sA + a_off(tid, 0)) = *reinterpret_cast<uint4 *>(h);
                *reinterpret_cast<uint4 *>(sA + a_off(tid, 1)) = *reinterpret_cast<uint4 *>(h + 8);
            } else {
                // thread-per-row gather: 12-14 independent 16-byte loads in flight per thread.  (A warp-cooperative walk with
                // whole rows per load instruction, as in k_fp1_fused, was measured 0.47 ms SLOWER here: with 2-3 CTAs per SM
                // the memory-level parallelism of the per-thread version matters more than the coalescing.)
                const uint4 *src = reinterpret_cast<const uint4 *>(a.feat + ((size_t)b * a.N + i) * C);
#pragma unroll
                for (int cc = 0; cc < C / 8; cc++) *reinterpret_cast<uint4 *>(sA + a_off(tid, cc)) = __ldg(src + cc);
                const float *p = a.xyz + ((size_t)b * a.N + i) * 3;
                __half h[8];
#pragma unroll
                for (int k = 0; k < 3; k++) split_hi_lo(__fsub_rn(p[k], c[k]), h[k], h[3 + k]);
                h[6] = h[7] = __float2half_rn(0.f);
                *reinterpret_cast<uint4 *>(sA + a_off(tid, C / 8)) = *reinterpret_cast<uint4 *>(h);
#pragma unroll
                for (int cc = C / 8 + 1; cc < K0 / 8; cc++) *reinterpret_cast<uint4 *>(sA + a_off(tid, cc)) = make_uint4(0, 0, 0, 0);
            }
        }
        SA_LAYER_SYNC()
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            issue<K0, N1>(tmem_base, sa_u, sw1_u, bar);
        }
        SA_WAIT()
        epilogue_to_operand<N1>(taddr, sA, tid, s_bias);
        SA_LAYER_SYNC()
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            issue<N1, N2>(tmem_base, sa_u, sw2_u, bar);
        }
        SA_WAIT()
        if (N3 == 0) {
            // ---- two-layer mode: relu(. + bias) -> fp16 -> this thread's row of the activation buffer
            __half *orow = a.out + (size_t)((unsigned)tile * 128u + (unsigned)tid) * a.ldo;
#pragma unroll 1
            for (int h = 0; h < N2; h += 16) {
                uint32_t u[16];
                umma::tmem_ld16(taddr + h, u);
                uint32_t p[8];
#pragma unroll
                for (int i = 0; i < 8; i += 2) {
                    const float4 bv = *reinterpret_cast<const float4 *>(s_bias + N1 + h + 2 * i);
                    p[i] = umma::pack_half2_sat(fmaxf(__uint_as_float(u[2 * i]) + bv.x, 0.f), fmaxf(__uint_as_float(u[2 * i + 1]) + bv.y, 0.f));
                    p[i + 1] = umma::pack_half2_sat(fmaxf(__uint_as_float(u[2 * i + 2]) + bv.z, 0.f), fmaxf(__uint_as_float(u[2 * i + 3]) + bv.w, 0.f));
                }
                *reinterpret_cast<uint4 *>(orow + h) = make_uint4(p[0], p[1], p[2], p[3]);
                *reinterpret_cast<uint4 *>(orow + h + 8) = make_uint4(p[4], p[5], p[6], p[7]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            group_sync();         // every accumulator row has been read before the next tile's first MMA
            continue;
        }
        epilogue_to_operand<N2>(taddr, sA, tid, s_bias + N1);
        SA_LAYER_SYNC()
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            issue<N2, (N3 > 0 ? N3 : 16)>(tmem_base, sa_u, sw3_u, bar);
        }
        SA_WAIT()
        // ---- last layer: max over the G rows of each group, then bias + ReLU (they commute with max)
        {
            float *sc = reinterpret_cast<float *>(sA) + warp * (32 * 17);      // the operand tile is free again
            const int cl = lane & 15, half = lane >> 4;
            const long long row0 = (long long)tile * 128 + warp * 32;
#pragma unroll 1
            for (int c = 0; c < N3; c += 16) {
                uint32_t u[16];
                umma::tmem_ld16(taddr + c, u);
#pragma unroll
                for (int i = 0; i < 16; i++) sc[lane * 17 + i] = __uint_as_float(u[i]);
                __syncwarp();
                float mx = -INFINITY;
#pragma unroll
                for (int r = 0; r < 16; r++) mx = fmaxf(mx, sc[(half * 16 + r) * 17 + cl]);
                __syncwarp();
                const float bv = s_bias[N1 + N2 + c + cl];
                if (G == 32) {
                    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
                    if (half == 0)
                        a.out[(size_t)(row0 / 32) * a.ldo + a.col_off + c + cl] = __float2half_rn(fminf(fmaxf(mx + bv, 0.f), 65504.f));
                } else {
                    a.out[(size_t)(row0 / 16 + half) * a.ldo + a.col_off + c + cl] = __float2half_rn(fminf(fmaxf(mx + bv, 0.f), 65504.f));
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        group_sync();             // pooling scratch (aliasing the operand tile) is free before the next gather
    }
#undef SA_LAYER_SYNC
#undef SA_WAIT
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        __syncwarp();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_all), "n"(TCOLS * NG) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// fp1 + classifier head as ONE persistent kernel (PointNetFeaturePropagation pointnet2_utils.py:278-317 for l0 <- l1, then
// conv1 + bn1 + relu and conv2 + log_softmax, pointnet2.py:33-41): per tile of 128 fine points the 3-NN interpolation of the
// level-1 features (indices / weights from the 3-NN search) is the operand producer, four 128 x 128 layers run back to back
// with the activations in shared memory / TMEM, and the last epilogue applies the 128 -> 2 head.  The four round trips of
// [rows][128] fp16 activations through HBM of the layer-by-layer path (1.3 GB per forward of 256 clouds) disappear; every
// rounding point (fp16 operand, fp16 activations, fp32 head in channel order) is the one of the un-fused kernels.
// 256 threads = two independent halves of 128 (thread = row = TMEM lane; half h owns operand tile h and TMEM columns
// [128 h, 128 h + 128)) that share the weight images: one half's gather overlaps the other's layers.
struct FpArgs {
    const __half *feat2;    // [B][S][128] level-1 features after fp2
    const int4 *knn_i;      // [B][N] three nearest level-1 points of every level-0 point
    const float4 *knn_w;    // [B][N] their normalised inverse-distance weights
    const uint8_t *wimg;    // 4 weight images (fp1 conv 0..2, conv1), 128 rows x 2 blocks x 128 bytes each
    const float *bias;      // [4][128]
    const float *w2, *b2;   // head [2][128], [2]
    long long *pred;        // [B * N]
    float *score, *logp;    // [B * N], [B * N][2] or null
    int N, S;
    unsigned rows;          // B * N
};
constexpr int kFpC = 128;
constexpr size_t kFpSmem = 1024 + 2 * 32768 + 4 * 32768 + (4 * kFpC + 2 * kFpC + 2) * 4 + 64;

__device__ __forceinline__ void half_sync(int half) { asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory"); }

__global__ void __launch_bounds__(256, 1) k_fp1_fused(const FpArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
    const int tid = threadIdx.x, half = tid >> 7, t = tid & 127, warp = tid >> 5, lane = tid & 31;
    uint8_t *sA = smem + half * 32768;                       // this half's operand tile
    uint8_t *sW = smem + 2 * 32768;                          // 4 x 32 KB
    float *s_bias = reinterpret_cast<float *>(sW + 4 * 32768);
    float *s_w2 = s_bias + 4 * kFpC;                         // [2][128] + [2]
    uint64_t *bar = reinterpret_cast<uint64_t *>(s_w2 + 2 * kFpC + 2);      // one per half
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bar + 2);
    if (tid == 0) {
        umma::mbar_init(bar, 1); umma::mbar_init(bar + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(a.wimg);
        uint4 *dst = reinterpret_cast<uint4 *>(sW);
        for (int i = tid; i < 4 * 32768 / 16; i += 256) dst[i] = src[i];
        for (int i = tid; i < 4 * kFpC; i += 256) s_bias[i] = a.bias[i];
        for (int i = tid; i < 2 * kFpC; i += 256) s_w2[i] = a.w2[i];
        if (tid < 2) s_w2[2 * kFpC + tid] = a.b2[tid];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_ptr + (uint32_t)half * 128u;
    const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t sa_u = smem_u32(sA), sw_u = smem_u32(sW);
    uint64_t *mybar = bar + half;
    uint32_t phase = 0;
    const unsigned ntiles = (a.rows + 127u) / 128u;

#define FP_LAYER_SYNC()                                                   \
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          \
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");      \
    half_sync(half);
#define FP_WAIT()                                                         \
    umma::mbar_wait(mybar, phase); phase ^= 1u;                           \
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    for (unsigned tile = blockIdx.x * 2u + (unsigned)half; tile < ntiles; tile += gridDim.x * 2u) {
        const unsigned r = tile * 128u + (unsigned)t;
        const bool valid = r < a.rows;
        // ---- operand: interpolated level-1 features, (a0 * w0 + a1 * w1) + a2 * w2 in fp32 -> fp16.  Warp-cooperative: the
        // warp walks its 32 rows two at a time, 16 lanes per row, one 16-byte chunk per lane -- every load instruction reads
        // whole 256-byte feature rows (a thread-per-row gather touches 32 different rows per instruction).
        {
            int4 ki = make_int4(0, 0, 0, 0);
            float4 kw = make_float4(0.f, 0.f, 0.f, 0.f);
            unsigned base = 0u;                 // first feature row of this point's cloud
            if (valid) { ki = a.knn_i[r]; kw = a.knn_w[r]; base = (r / (unsigned)a.N) * (unsigned)a.S; }
            const int sub = lane >> 4, cc = lane & 15;
#pragma unroll 4
            for (int j = 0; j < 32; j += 2) {
                const int src = j + sub;
                const unsigned rb = __shfl_sync(0xffffffffu, base, src);
                const int n0 = __shfl_sync(0xffffffffu, ki.x, src), n1 = __shfl_sync(0xffffffffu, ki.y, src), n2 = __shfl_sync(0xffffffffu, ki.z, src);
                const float w0 = __shfl_sync(0xffffffffu, kw.x, src), w1 = __shfl_sync(0xffffffffu, kw.y, src), w2 = __shfl_sync(0xffffffffu, kw.z, src);
                const uint4 q0 = __ldg(reinterpret_cast<const uint4 *>(a.feat2 + ((size_t)rb + n0) * kFpC) + cc);
                const uint4 q1 = __ldg(reinterpret_cast<const uint4 *>(a.feat2 + ((size_t)rb + n1) * kFpC) + cc);
                const uint4 q2 = __ldg(reinterpret_cast<const uint4 *>(a.feat2 + ((size_t)rb + n2) * kFpC) + cc);
                const __half2 *h0 = reinterpret_cast<const __half2 *>(&q0), *h1 = reinterpret_cast<const __half2 *>(&q1),
                              *h2 = reinterpret_cast<const __half2 *>(&q2);
                uint4 o;
                __half2 *ho = reinterpret_cast<__half2 *>(&o);
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const float2 a0 = __half22float2(h0[e]), a1 = __half22float2(h1[e]), a2 = __half22float2(h2[e]);
                    const float x = __fadd_rn(__fadd_rn(__fmul_rn(a0.x, w0), __fmul_rn(a1.x, w1)), __fmul_rn(a2.x, w2));
                    const float y = __fadd_rn(__fadd_rn(__fmul_rn(a0.y, w0), __fmul_rn(a1.y, w1)), __fmul_rn(a2.y, w2));
                    ho[e] = __floats2half2_rn(fminf(x, 65504.f), fminf(y, 65504.f));
                }
                // rows past the end carry weights 0 and neighbour 0 of cloud 0: finite values that nobody reads
                *reinterpret_cast<uint4 *>(sA + a_off((warp & 3) * 32 + src, cc)) = o;
            }
        }
        // ---- fp1 conv 0..2
#pragma unroll 1
        for (int l = 0; l < 3; l++) {
            FP_LAYER_SYNC()
            if (t == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                issue<kFpC, kFpC>(tmem_base, sa_u, sw_u + l * 32768, mybar);
            }
            FP_WAIT()
            epilogue_to_operand<kFpC>(taddr, sA, t, s_bias + l * kFpC);
        }
        // ---- conv1 + bn1 + relu, then the 128 -> 2 head on the fp16-rounded activations in channel order
        FP_LAYER_SYNC()
        if (t == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            issue<kFpC, kFpC>(tmem_base, sa_u, sw_u + 3 * 32768, mybar);
        }
        FP_WAIT()
        float z0 = 0.f, z1 = 0.f;
#pragma unroll 1
        for (int h = 0; h < kFpC; h += 32) {
            uint32_t u[32];
            umma::tmem_ld32(taddr + h, u);
#pragma unroll
            for (int i = 0; i < 32; i += 4) {            // bias and head weights of 4 columns per 16-byte shared-memory load
                const float4 bv = *reinterpret_cast<const float4 *>(s_bias + 3 * kFpC + h + i);
                const float4 wa = *reinterpret_cast<const float4 *>(s_w2 + h + i), wb = *reinterpret_cast<const float4 *>(s_w2 + kFpC + h + i);
                const float x0 = fmaxf(__uint_as_float(u[i]) + bv.x, 0.f), x1 = fmaxf(__uint_as_float(u[i + 1]) + bv.y, 0.f);
                const float x2 = fmaxf(__uint_as_float(u[i + 2]) + bv.z, 0.f), x3 = fmaxf(__uint_as_float(u[i + 3]) + bv.w, 0.f);
                const uint32_t pk0 = umma::pack_half2_sat(x0, x1), pk1 = umma::pack_half2_sat(x2, x3);
                const float2 f0 = __half22float2(*reinterpret_cast<const __half2 *>(&pk0));
                const float2 f1 = __half22float2(*reinterpret_cast<const __half2 *>(&pk1));
                z0 = fmaf(f0.x, wa.x, z0); z0 = fmaf(f0.y, wa.y, z0); z0 = fmaf(f1.x, wa.z, z0); z0 = fmaf(f1.y, wa.w, z0);
                z1 = fmaf(f0.x, wb.x, z1); z1 = fmaf(f0.y, wb.y, z1); z1 = fmaf(f1.x, wb.z, z1); z1 = fmaf(f1.y, wb.w, z1);
            }
        }
        if (valid) {
            z0 += s_w2[2 * kFpC]; z1 += s_w2[2 * kFpC + 1];
            const float m = fmaxf(z0, z1);
            const float lse = m + logf(expf(z0 - m) + expf(z1 - m));
            const float l0 = z0 - lse, l1 = z1 - lse;
            if (a.logp) { a.logp[(size_t)r * 2] = l0; a.logp[(size_t)r * 2 + 1] = l1; }
            a.pred[r] = (l1 > l0) ? 1 : 0;
            const float e0 = expf(l0), e1 = expf(l1);
            a.score[r] = e1 / (e0 + e1);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        half_sync(half);          // every thread of the half has read its accumulator row before the next tile's first MMA
    }
#undef FP_LAYER_SYNC
#undef FP_WAIT
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_ptr), "n"(256) : "memory");
    }
}

}  // namespace safused
