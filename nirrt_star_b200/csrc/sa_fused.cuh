// sa_fused.cuh -- first set-abstraction level of PointNet++ as ONE persistent kernel per radius:
// grouping (gather + centroid subtraction), the three 1x1 convolutions (+ folded BatchNorm + ReLU)
// and the max over the K neighbours (pointnet2_utils.py:243-259), for 128 (centroid, neighbour) rows
// per tile.  The tile's activations never leave the SM: the gathered operand and every intermediate
// layer live in one 16 KB shared-memory tile in the tcgen05 K-major / 128-byte-swizzle layout, the
// accumulators in TMEM; weights (<= 16 KB) are copied into shared memory once per CTA.
//
//   128 threads, thread = row = TMEM lane.  Per layer: all threads write their row of the operand,
//   fence.proxy.async + __syncthreads, thread 0 issues tcgen05.mma (K <= 64: one or two
//   instructions) and commits to an mbarrier, all threads wait, tcgen05.ld their accumulator row.
//   Up to 8 CTAs per SM hide each other's round trips.
//
// sa1 has the 6-channel network input as point features: rows are the 16 halves
//   [x y z start goal free | rx ry rz | x_lo y_lo z_lo | rx_lo ry_lo rz_lo | 0]
// (hi + lo split of the coordinates, see k_group_sa1), K = 16; layer widths 16/16/32 (K = 16
// neighbours) and 32/32/64 (K = 32).
#pragma once
#include "umma_gemm.cuh"

namespace safused {

using umma::smem_u32;

// byte offset of 16-byte chunk `c` of row `r` in a K-major SWIZZLE_128B tile (rows of 128 bytes,
// 8-row groups of 1024 bytes, chunk index XOR (r mod 8))
__host__ __device__ __forceinline__ int sw128_off(int r, int c) { return (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4); }

struct Args {
    const float *in6;       // [B][N][6]
    const float *new_xyz;   // [B][S][3]
    const int *gidx;        // [B][S][G]
    const uint8_t *wimg;    // W1 | W2 | W3 shared-memory images (N_l rows x 128 B, swizzled)
    const float *bias;      // [N1 + N2 + N3]
    __half *out;            // [B][S][ldo]
    int N, S, B, ldo, col_off;
};

template <int NCOLS>
__device__ __forceinline__ void ld_row(uint32_t taddr, float *v) {       // NCOLS accumulator columns of this thread's lane
#pragma unroll
    for (int h = 0; h < NCOLS; h += 16) {
        uint32_t u[16];
        umma::tmem_ld16(taddr + h, u);
#pragma unroll
        for (int i = 0; i < 16; i++) v[h + i] = __uint_as_float(u[i]);
    }
}

// relu(v + bias) -> fp16 -> this thread's row of the next operand tile
template <int NCOLS>
__device__ __forceinline__ void store_row(uint8_t *sA, int r, const float *v, const float *bias) {
#pragma unroll
    for (int c = 0; c < NCOLS / 8; c++) {
        uint32_t p[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float a = fmaxf(v[c * 8 + 2 * i] + bias[c * 8 + 2 * i], 0.f);
            const float b = fmaxf(v[c * 8 + 2 * i + 1] + bias[c * 8 + 2 * i + 1], 0.f);
            p[i] = umma::pack_half2_sat(a, b);
        }
        *reinterpret_cast<uint4 *>(sA + sw128_off(r, c)) = make_uint4(p[0], p[1], p[2], p[3]);
    }
}

template <int K>
__device__ __forceinline__ void issue(uint32_t d_tmem, uint32_t sa, uint32_t sw, int n, uint64_t *bar) {
    const uint32_t idesc = umma::make_idesc(n);
#pragma unroll
    for (int k = 0; k < K / 16; k++)
        umma::mma_f16(d_tmem, umma::make_smem_desc(sa + k * 32), umma::make_smem_desc(sw + k * 32), idesc, (uint32_t)(k != 0));
    umma::mma_commit(bar);
}

template <int G, int N1, int N2, int N3>
__global__ void __launch_bounds__(128) k_sa1_fused(const Args a) {
    constexpr int K0 = 16;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
    uint8_t *sA = smem;                                   // 16 KB operand tile (also the pooling scratch)
    uint8_t *sW1 = sA + 16384, *sW2 = sW1 + N1 * 128, *sW3 = sW2 + N2 * 128;
    float *s_bias = reinterpret_cast<float *>(sW3 + N3 * 128);
    uint64_t *bar = reinterpret_cast<uint64_t *>(s_bias + N1 + N2 + N3 + ((N1 + N2 + N3) & 1));
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int TCOLS = N3 <= 32 ? 32 : 64;

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
        for (int i = tid; i < (N1 + N2 + N3) * 8; i += 128) dst[i] = src[i];
        for (int i = tid; i < N1 + N2 + N3; i += 128) s_bias[i] = a.bias[i];
        // rows of the operand tile beyond the valid K columns are never read by the MMAs (K <= 64)
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const uint32_t sa_u = smem_u32(sA), sw1_u = smem_u32(sW1), sw2_u = smem_u32(sW2), sw3_u = smem_u32(sW3);
    uint32_t phase = 0;
    const long long rows = (long long)a.B * a.S * G;
    const int ntiles = (int)(rows / 128);

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        // ---- grouping: this thread's row (pointnet2_utils.py:246-253)
        {
            const long long r = (long long)tile * 128 + tid;
            const long long bs = r / G;
            const int b = (int)(bs / a.S);
            const int i = a.gidx[r];
            const float *f = a.in6 + ((size_t)b * a.N + i) * 6;
            const float *c = a.new_xyz + bs * 3;
            __half h[16];
            float v[9];
#pragma unroll
            for (int k = 0; k < 6; k++) v[k] = f[k];
#pragma unroll
            for (int k = 0; k < 3; k++) v[6 + k] = __fsub_rn(f[k], c[k]);
#pragma unroll
            for (int k = 0; k < 3; k++) { h[k] = __float2half_rn(v[k]); h[9 + k] = __float2half_rn(__fsub_rn(v[k], __half2float(h[k]))); }
#pragma unroll
            for (int k = 3; k < 6; k++) h[k] = __float2half_rn(v[k]);
#pragma unroll
            for (int k = 6; k < 9; k++) { h[k] = __float2half_rn(v[k]); h[6 + k] = __float2half_rn(__fsub_rn(v[k], __half2float(h[k]))); }
            h[15] = __float2half_rn(0.f);
            *reinterpret_cast<uint4 *>(sA + sw128_off(tid, 0)) = *reinterpret_cast<uint4 *>(h);
            *reinterpret_cast<uint4 *>(sA + sw128_off(tid, 1)) = *reinterpret_cast<uint4 *>(h + 8);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            issue<K0>(tmem_base, sa_u, sw1_u, N1, bar);
        }
        umma::mbar_wait(bar, phase); phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        {
            float v[N1];
            ld_row<N1>(taddr, v);
            store_row<N1>(sA, tid, v, s_bias);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            issue<N1>(tmem_base, sa_u, sw2_u, N2, bar);
        }
        umma::mbar_wait(bar, phase); phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        {
            float v[N2];
            ld_row<N2>(taddr, v);
            store_row<N2>(sA, tid, v, s_bias + N1);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            issue<N2>(tmem_base, sa_u, sw3_u, N3, bar);
        }
        umma::mbar_wait(bar, phase); phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
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
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TCOLS) : "memory");
    }
}

template <int N1, int N2, int N3>
constexpr size_t smem_bytes() { return 1024 + 16384 + (size_t)(N1 + N2 + N3) * 128 + (size_t)(N1 + N2 + N3 + 2) * 4 + 64; }

}  // namespace safused
