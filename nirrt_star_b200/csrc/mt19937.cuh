// mt19937.cuh -- numpy legacy (np.random.*) MT19937 stream, one state per planning problem.
//
// The reference draws every sample from the process-global numpy generator
// (rrt_base_3d.py:49-58, irrt_star_3d.py:147-157, nirrt_star_png_3d.py:116,130), so identical
// trees under a fixed seed require consuming the identical 32-bit word stream.  Seeding stays on
// the host (np.random.RandomState(seed).get_state() -> 624 key words + pos); generation runs here.
//
// Layout: two 624-word blocks per stream.  The block that follows the current one is produced by
// the whole thread block ahead of time (mt_prepare_next, three barrier-separated phases); the
// single sampling thread then only flips an index when it runs out of words.  A serial in-place
// regeneration remains as the fallback for the (rare) iteration that consumes more than one block.
#pragma once
#include <stdint.h>

namespace nirrt {

struct MtState {
    uint32_t key[2][624];
    int pos;       // next word in key[cur], 0..624
    int cur;       // which block is current
    int has_next;  // key[cur^1] already holds the following block
    int pad;
};

__device__ __forceinline__ uint32_t mt_twist(uint32_t a, uint32_t b) {
    uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
    return (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

// Called by ALL threads of the block (uniform branch on global state).
__device__ __forceinline__ void mt_prepare_next(MtState *s, int margin) {
    const int pos = s->pos, has_next = s->has_next, cur = s->cur;
    if (has_next || pos < 624 - margin) return;   // uniform: every thread reads the same words
    const uint32_t *o = s->key[cur];
    uint32_t *nx = s->key[cur ^ 1];
    for (int i = threadIdx.x; i < 227; i += blockDim.x) nx[i] = o[i + 397] ^ mt_twist(o[i], o[i + 1]);
    __syncthreads();
    for (int i = 227 + threadIdx.x; i < 454; i += blockDim.x) nx[i] = nx[i - 227] ^ mt_twist(o[i], o[i + 1]);
    __syncthreads();
    for (int i = 454 + threadIdx.x; i < 623; i += blockDim.x) nx[i] = nx[i - 227] ^ mt_twist(o[i], o[i + 1]);
    if (threadIdx.x == 0) nx[623] = nx[396] ^ mt_twist(o[623], nx[0]);
    __syncthreads();
    if (threadIdx.x == 0) s->has_next = 1;
    __syncthreads();
}

// Single-thread word stream over an MtState (registers hold pos/cur; call flush() when done).
__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}
// mt19937_next_double from two consecutive raw state words: (a >> 5, b >> 6) -> 53-bit fraction
__device__ __forceinline__ double mt_double(uint32_t raw0, uint32_t raw1) {
    const int a = (int)(mt_temper(raw0) >> 5), b = (int)(mt_temper(raw1) >> 6);
    return __ddiv_rn(__dadd_rn(__dmul_rn((double)a, 67108864.0), (double)b), 9007199254740992.0);
}

constexpr int kMtCache = 64;   // words of the current block staged in shared memory ahead of the sampling thread

// Called by ALL threads: copies the next kMtCache words of the stream's current block into `cache`
// (one coalesced round trip instead of one dependent global load per word drawn).
__device__ __forceinline__ void mt_stage_words(const MtState *s, uint32_t *cache) {
    const int pos = s->pos;
    const uint32_t *k = s->key[s->cur];
    for (int j = threadIdx.x; j < kMtCache; j += blockDim.x) cache[j] = pos + j < 624 ? k[pos + j] : 0u;
}

struct MtStream {
    MtState *s;
    uint32_t *k;
    int pos;
    const uint32_t *cache;   // words [lo, lo + kMtCache) of the current block, or null
    int lo;
    __device__ __forceinline__ explicit MtStream(MtState *st, const uint32_t *staged = nullptr)
        : s(st), k(st->key[st->cur]), pos(st->pos), cache(staged), lo(st->pos) {}
    __device__ void refill() {
        cache = nullptr;
        if (s->has_next) {
            s->cur ^= 1;
            s->has_next = 0;
            k = s->key[s->cur];
        } else {  // serial in-place regeneration (mt19937_gen of numpy's randomkit)
            int i;
            for (i = 0; i < 624 - 397; i++) k[i] = k[i + 397] ^ mt_twist(k[i], k[i + 1]);
            for (; i < 623; i++) k[i] = k[i - 227] ^ mt_twist(k[i], k[i + 1]);
            k[623] = k[396] ^ mt_twist(k[623], k[0]);
        }
        pos = 0;
    }
    __device__ __forceinline__ uint32_t next() {
        if (pos == 624) refill();
        const uint32_t y = (cache && pos - lo < kMtCache) ? cache[pos - lo] : k[pos];
        pos++;
        return mt_temper(y);
    }
    // words already consumed elsewhere from the staged cache (same block, pos + n <= 624)
    __device__ __forceinline__ void skip(int n) { pos += n; }
    // mt19937_next_double: (a >> 5, b >> 6) -> 53-bit fraction
    __device__ __forceinline__ double next_double() {
        int a = (int)(next() >> 5), b = (int)(next() >> 6);
        return __ddiv_rn(__dadd_rn(__dmul_rn((double)a, 67108864.0), (double)b), 9007199254740992.0);
    }
    // np.random.uniform(lo, hi) == lo + (hi - lo) * next_double()
    __device__ __forceinline__ double uniform(double lo, double hi) {
        return __dadd_rn(lo, __dmul_rn(__dsub_rn(hi, lo), next_double()));
    }
    // np.random.randint(0, high): masked rejection on 32-bit words; no draw when high == 1
    __device__ __forceinline__ long long randint(long long high) {
        unsigned long long rng = (unsigned long long)(high - 1);
        if (rng == 0) return 0;
        uint32_t mask = (uint32_t)rng;
        mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
        uint32_t v;
        do { v = next() & mask; } while (v > rng);
        return (long long)v;
    }
    __device__ __forceinline__ void flush() { s->pos = pos; }
};

}  // namespace nirrt
