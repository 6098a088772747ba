// sa_fused.cuh -- a set-abstraction scale of PointNet++ as ONE persistent kernel:
// grouping (gather + centroid subtraction), the three 1x1 convolutions (+ folded BatchNorm + ReLU)
// and the max over the K neighbours (pointnet2_utils.py:243-259), for 128 (centroid, neighbour) rows
// per tile.  The tile's activations never leave the SM: the gathered operand and every intermediate
// layer live in one shared-memory tile in the tcgen05 K-major / 128-byte-swizzle layout (one 16 KB
// block per 64 channels), the accumulators in TMEM; the weights of all three layers are copied into
// shared memory once per CTA.  Used for sa1 (6 input channels) and sa2 (96); the wider levels keep
// their weights in L2 and go through umma::k_gemm.
//
//   128 threads, thread = row = TMEM lane.  Per layer: all threads write their row of the operand,
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
#pragma unroll 1
    for (int h = 0; h < NCOLS; h += 16) {
        uint32_t u[16];
        umma::tmem_ld16(taddr + h, u);
        uint32_t p[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float a = fmaxf(__uint_as_float(u[2 * i]) + bias[h + 2 * i], 0.f);
            const float b = fmaxf(__uint_as_float(u[2 * i + 1]) + bias[h + 2 * i + 1], 0.f);
            p[i] = umma::pack_half2_sat(a, b);
        }
        *reinterpret_cast<uint4 *>(sA + a_off(r, h >> 3)) = make_uint4(p[0], p[1], p[2], p[3]);
        *reinterpret_cast<uint4 *>(sA + a_off(r, (h >> 3) + 1)) = make_uint4(p[4], p[5], p[6], p[7]);
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
};

// C == 0: sa1 (in6 input); C > 0: features [.,C] of the previous level + relative coordinates
template <int G, int C, int K0, int N1, int N2, int N3>
__global__ void __launch_bounds__(128) k_sa_fused(const Args a) {
    typedef Smem<K0, N1, N2, N3> SM;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
    uint8_t *sA = smem;                                   // operand tile (also the pooling scratch)
    uint8_t *sW1 = sA + SM::kABytes, *sW2 = sW1 + SM::kW1, *sW3 = sW2 + SM::kW2;
    float *s_bias = reinterpret_cast<float *>(sW3 + SM::kW3);
    uint64_t *bar = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(s_bias) + SM::kBias);
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int NMAX = imax(imax(N1, N2), N3);
    constexpr int TCOLS = NMAX <= 32 ? 32 : (NMAX <= 64 ? 64 : (NMAX <= 128 ? 128 : 256));

    if (tid == 0) {
        umma::mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(TCOLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(a.wimg);
        uint4 *dst = reinterpret_cast<uint4 *>(sW1);
        for (int i = tid; i < (SM::kW1 + SM::kW2 + SM::kW3) / 16; i += 128) dst[i] = src[i];
        for (int i = tid; i < N1 + N2 + N3; i += 128) s_bias[i] = a.bias[i];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const uint32_t sa_u = smem_u32(sA), sw1_u = smem_u32(sW1), sw2_u = smem_u32(sW2), sw3_u = smem_u32(sW3);
    uint32_t phase = 0;
    const int nvt = a.B * a.chunk_tiles;

#define SA_LAYER_SYNC()                                                   \
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          \
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");      \
    __syncthreads();
#define SA_WAIT()                                                         \
    umma::mbar_wait(bar, phase); phase ^= 1u;                             \
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    for (int vt = blockIdx.x; vt < nvt; vt += gridDim.x) {
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
        epilogue_to_operand<N2>(taddr, sA, tid, s_bias + N1);
        SA_LAYER_SYNC()
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            issue<N2, N3>(tmem_base, sa_u, sw3_u, bar);
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
        __syncthreads();          // pooling scratch (aliasing the operand tile) is free before the next gather
    }
#undef SA_LAYER_SYNC
#undef SA_WAIT
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TCOLS) : "memory");
    }
}

}  // namespace safused
