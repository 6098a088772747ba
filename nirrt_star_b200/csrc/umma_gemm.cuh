// umma_gemm.cuh -- sm_100a tensor-core GEMM for the PointNet++ 1x1 convolutions.
//
//   D[m, n] = sum_k A[m, k] * W[n, k]          A: [M][K] fp16 (activations, K-major)
//                                              W: [N][K] fp16 (conv weight with BatchNorm folded,
//                                                 PyTorch [C_out][C_in] order == K-major)
//   MODE_STORE: out[m][n]            = fp16(relu(D + bias[n]))                 (conv+BN+ReLU)
//   MODE_POOL : out[m / G][off + n]  = fp16(relu(max_{rows of group} D + bias[n]))
//               (conv+BN+ReLU followed by torch.max over the K neighbours of a group,
//                pointnet2_utils.py:255-259; relu and +bias commute with max)
//
// One CTA computes a 128 x BN tile: tcgen05.mma (cta_group::1, kind::f16, M=128, N=BN, K=16)
// issued by one thread, operands staged in shared memory by TMA (cp.async.bulk.tensor.2d, 128-byte
// swizzle, 64-element K blocks, a ring of mbarrier-guarded stages), fp32 accumulators in TMEM,
// read back with tcgen05.ld by four epilogue warps (one TMEM lane quarter each).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace umma {

constexpr int kBM = 128;        // rows per CTA tile == TMEM lanes
constexpr int kBK = 64;         // fp16 elements per K block == one 128-byte swizzle row
constexpr int kThreads = 192;
constexpr int kMaxStages = 4;
constexpr int MODE_STORE = 0, MODE_POOL = 1;

struct GemmArgs {
    int M, N, K;           // rows, output channels (multiple of 16), K (multiple of 16)
    int BN;                // N tile per CTA (multiple of 16, <= 256); grid.y = ceil(N / BN)
    int stages;
    int ldo;               // output row pitch (elements)
    int col_off;           // MODE_POOL: column offset inside the output row (scale concat)
    int group;             // MODE_POOL: rows per group (16 or 32)
    const float *bias;     // [N]
    __half *out;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a mis-programmed pipeline traps (the launch fails) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (uint32_t spins = 0;; spins++) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return;
        if (spins > (1u << 24)) { __trap(); }
    }
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// K-major, 128-byte-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart
// (SBO = 1024 >> 4), LBO unused for swizzled K-major (1), descriptor version 1 (sm_100),
// layout type 2 = SWIZZLE_128B.  The tile base is 1024-byte aligned; advancing by UMMA_K = 16
// fp16 (32 bytes) inside the swizzle row adds 32 >> 4 to the start-address field.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)1 << 16;                 // leading byte offset (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;       // stride byte offset
    d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
    return d;
}
// instruction descriptor: D = f32, A = B = f16, both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t make_idesc(int n) {
    return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t *v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 32 accumulator columns per round trip (one wait for both halves)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t *v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t pack_half2_sat(float a, float b) {
    a = fminf(a, 65504.f); b = fminf(b, 65504.f);      // inputs are >= 0 (post-ReLU)
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}

struct SmemLayout {
    int stage_bytes, a_bytes, bars_off, bias_off, scratch_off, total;
};
__host__ __device__ inline SmemLayout smem_layout(int BN, int stages) {
    SmemLayout L;
    L.a_bytes = kBM * 128;
    L.stage_bytes = L.a_bytes + ((BN * 128 + 1023) & ~1023);
    L.bars_off = stages * L.stage_bytes;
    L.bias_off = L.bars_off + 128;                       // 2*stages+4 barriers + tmem pointer
    L.scratch_off = L.bias_off + 256 * 4;
    L.total = L.scratch_off + 4 * 32 * 20 * 4;           // per epilogue warp: 32 x 17 floats (pool) or 32 x 80 B (store staging)
    return L;
}
__host__ __device__ inline int tmem_cols_for(int bn) {
    int c = 32;
    while (c < bn) c <<= 1;
    return c;
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Persistent: each CTA walks the M tiles blockIdx.x, blockIdx.x + gridDim.x, ... of its N slice.
// The TMA ring keeps running across tiles, and the accumulator is double buffered in TMEM
// (2 x tmem_cols_for(bn) columns) so the epilogue of tile t overlaps the loads and MMAs of tile t+1.
template <int MODE>
__global__ void __launch_bounds__(kThreads) k_gemm(const __grid_constant__ CUtensorMap tmA,
                                                   const __grid_constant__ CUtensorMap tmW, const GemmArgs g) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment of the operand tiles (swizzle atom)
    uint8_t *smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
    const SmemLayout L = smem_layout(g.BN, g.stages);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + L.bars_off);
    uint64_t *empty = full + kMaxStages;
    uint64_t *tfull = empty + kMaxStages;       // [2] accumulator buffer ready for the epilogue
    uint64_t *tempty = tfull + 2;               // [2] accumulator buffer drained
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tempty + 2);
    float *s_bias = reinterpret_cast<float *>(smem + L.bias_off);
    float *s_scratch = reinterpret_cast<float *>(smem + L.scratch_off);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.y * g.BN;
    const int bn = min(g.BN, g.N - n0);                  // multiple of 16
    const int nkb = (g.K + kBK - 1) / kBK;
    const int ntiles = (g.M + kBM - 1) / kBM;
    const int buf_cols = tmem_cols_for(bn);
    const int tmem_cols = 2 * buf_cols;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmW);
        for (int s = 0; s < g.stages; s++) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int b = 0; b < 2; b++) { mbar_init(tfull + b, 1); mbar_init(tempty + b, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < bn; i += kThreads) s_bias[i] = g.bias[n0 + i];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t tx = (uint32_t)(L.a_bytes + g.BN * 128);
            int it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int m0 = tile * kBM;
                for (int kb = 0; kb < nkb; kb++, it++) {
                    const int s = it % g.stages;
                    const uint32_t ph = (uint32_t)(it / g.stages) & 1u;
                    mbar_wait(empty + s, ph ^ 1u);
                    mbar_expect_tx(full + s, tx);
                    uint8_t *sa = smem + s * L.stage_bytes;
                    tma_load_2d(sa, &tmA, kb * kBK, m0, full + s);
                    tma_load_2d(sa + L.a_bytes, &tmW, kb * kBK, n0, full + s);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc(bn);
            int it = 0, tcount = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, tcount++) {
                const int buf = tcount & 1;
                mbar_wait(tempty + buf, (uint32_t)((tcount >> 1) & 1) ^ 1u);     // epilogue drained this buffer
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * buf_cols);
                for (int kb = 0; kb < nkb; kb++, it++) {
                    const int s = it % g.stages;
                    const uint32_t ph = (uint32_t)(it / g.stages) & 1u;
                    mbar_wait(full + s, ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sa = smem_u32(smem + s * L.stage_bytes);
                    const uint32_t sw = sa + L.a_bytes;
                    const int ksteps = min(kBK, g.K - kb * kBK) >> 4;
                    for (int k = 0; k < ksteps; k++)
                        mma_f16(d_tmem, make_smem_desc(sa + k * 32), make_smem_desc(sw + k * 32), idesc, (uint32_t)((kb | k) != 0));
                    mma_commit(empty + s);          // frees the stage once these MMAs have read it
                }
                mma_commit(tfull + buf);            // accumulator of this tile complete
            }
        }
    } else {
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        float *sc = s_scratch + q * (32 * 20);
        const int cl = lane & 15, half = lane >> 4;
        int tcount = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, tcount++) {
            const int buf = tcount & 1;
            const int m0 = tile * kBM;
            mbar_wait(tfull + buf, (uint32_t)((tcount >> 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int row = m0 + q * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * buf_cols);
            if (MODE == MODE_STORE) {
                // Each thread owns one output row; a thread can store only 16 B per instruction, and
                // 32 threads writing 16 B at the row pitch would hit 32 half-filled sectors (measured:
                // 3.6x DRAM write amplification).  The 32 x (<=32 columns) block of this warp is
                // therefore transposed through shared memory and written as whole 32-byte sectors:
                // 4 (or 2) lanes per row, 8 (or 16) rows per instruction.
                uint4 *stg = reinterpret_cast<uint4 *>(sc);          // [32 rows][5 x 16 B] (80-byte pitch)
                const int row0 = m0 + q * 32;
                for (int c = 0; c < bn; c += 32) {
                    const int w = min(32, bn - c);                   // 16 or 32 columns
                    for (int h = 0; h < w; h += 16) {
                        uint32_t v[16];
                        tmem_ld16(taddr + c + h, v);
                        uint32_t p[8];
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const float a = fmaxf(__uint_as_float(v[2 * i]) + s_bias[c + h + 2 * i], 0.f);
                            const float b = fmaxf(__uint_as_float(v[2 * i + 1]) + s_bias[c + h + 2 * i + 1], 0.f);
                            p[i] = pack_half2_sat(a, b);
                        }
                        stg[lane * 5 + (h >> 3)] = make_uint4(p[0], p[1], p[2], p[3]);
                        stg[lane * 5 + (h >> 3) + 1] = make_uint4(p[4], p[5], p[6], p[7]);
                    }
                    __syncwarp();
                    const int ppr = w >> 3;                          // 16-byte pieces per row: 2 or 4
                    const int rpi = 32 / ppr;                        // rows per store instruction
                    const int piece = lane % ppr;
                    for (int j = 0; j < 32; j += rpi) {
                        const int r = j + lane / ppr;
                        if (row0 + r < g.M)
                            *reinterpret_cast<uint4 *>(g.out + (size_t)(row0 + r) * g.ldo + n0 + c + piece * 8) = stg[r * 5 + piece];
                    }
                    __syncwarp();
                }
            } else {
                for (int c = 0; c < bn; c += 16) {
                    uint32_t v[16];
                    tmem_ld16(taddr + c, v);
#pragma unroll
                    for (int i = 0; i < 16; i++) sc[lane * 17 + i] = (row < g.M) ? __uint_as_float(v[i]) : -INFINITY;
                    __syncwarp();
                    float mx = -INFINITY;
#pragma unroll
                    for (int r = 0; r < 16; r++) mx = fmaxf(mx, sc[(half * 16 + r) * 17 + cl]);
                    __syncwarp();
                    if (g.group == 32) {
                        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
                        const int grp = (m0 + q * 32) / 32;
                        if (half == 0 && m0 + q * 32 < g.M)
                            g.out[(size_t)grp * g.ldo + g.col_off + n0 + c + cl] =
                                __float2half_rn(fminf(fmaxf(mx + s_bias[c + cl], 0.f), 65504.f));
                    } else {
                        const int grp = (m0 + q * 32) / 16 + half;
                        if (m0 + q * 32 + half * 16 < g.M)
                            g.out[(size_t)grp * g.ldo + g.col_off + n0 + c + cl] =
                                __float2half_rn(fminf(fmaxf(mx + s_bias[c + cl], 0.f), 65504.f));
                    }
                }
            }
            // this warp is done reading the buffer: let the MMA warp reuse it
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty + buf);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

}  // namespace umma
