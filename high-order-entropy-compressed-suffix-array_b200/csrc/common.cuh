// common.cuh -- shared helpers for libhkcsa (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <atomic>

#include "../../include/hkcsa.h"

namespace hkcsa {

// ---------------------------------------------------------------- errors
void set_error(const char *fmt, ...);

#define HK_CUDA(call)                                                                      \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            hkcsa::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,                  \
                             cudaGetErrorString(e__));                                     \
            return HKCSA_ECUDA;                                                            \
        }                                                                                  \
    } while (0)

// every kernel launch of the library is counted (bench.py reports gpu_launches)
extern std::atomic<unsigned long long> g_launches;
static inline void count_launch(unsigned long long k = 1) { g_launches.fetch_add(k, std::memory_order_relaxed); }

#define HK_LAUNCH_CHECK()                                                                  \
    do {                                                                                   \
        hkcsa::count_launch();                                                             \
        HK_CUDA(cudaGetLastError());                                                       \
    } while (0)

#define HK_REQUIRE(cond, code, msg)                                                        \
    do {                                                                                   \
        if (!(cond)) {                                                                     \
            hkcsa::set_error("%s:%d %s", __FILE__, __LINE__, msg);                         \
            return (code);                                                                 \
        }                                                                                  \
    } while (0)

// One pinned host page per calling thread for scalar read-backs (the library's only
// allocation).  Returns nullptr on failure (error already set).
void *pinned_page();   // 4096 bytes

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Carves aligned sub-buffers out of one caller-provided scratch allocation.
struct Carver {
    uint8_t *base;
    size_t used;
    explicit Carver(void *p) : base(static_cast<uint8_t *>(p)), used(0) {}
    template <typename T>
    T *take(size_t count)
    {
        used = align_up(used, 256);
        T *p = base ? reinterpret_cast<T *>(base + used) : nullptr;
        used += count * sizeof(T);
        return p;
    }
    size_t total() const { return align_up(used, 256); }
};

static inline uint32_t bits_for(uint64_t maxval)  // bits needed to represent values 0..maxval
{
    uint32_t b = 0;
    while (b < 64 && (maxval >> b) != 0) ++b;
    return b;
}

static inline int num_sms()
{
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

// ---------------------------------------------------------------- device helpers
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ULL;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t lanemask_lt()
{
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t *p, uint32_t v)
{
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// L2 eviction policies (createpolicy, sm_80+): streams that are touched once are marked evict-first so they do
// not push a randomly gathered table (the text under the BWT gather) out of the 126 MB L2.
__device__ __forceinline__ uint64_t l2_policy_evict_last(float fraction = 1.0f)
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, %1;" : "=l"(p) : "f"(fraction));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint32_t ld_u8_hint(const uint8_t *p, uint64_t pol)
{
    uint32_t v;
    asm volatile("ld.global.nc.L2::cache_hint.u8 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ uint4 ld_u32x4_hint(const uint32_t *p, uint64_t pol)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void st_u32_hint(uint32_t *p, uint32_t v, uint64_t pol)
{
    asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
}

// Exclusive warp scan (sum) of one uint32 per lane.
__device__ __forceinline__ uint32_t warp_excl_sum(uint32_t v, uint32_t &total)
{
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane_id() >= (uint32_t)o) x += y;
    }
    total = __shfl_sync(0xffffffffu, x, 31);
    return x - v;
}
__device__ __forceinline__ uint32_t warp_incl_max(uint32_t v)
{
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane_id() >= (uint32_t)o) x = max(x, y);
    }
    return x;
}

// ---------------------------------------------------------------- rank blocks
// One 32-byte block: low 32 bits of word 0 = ones before the block (relative to
// its superblock); the other 224 bits are payload.  See include/hkcsa.h.
struct __align__(32) RankBlock {
    uint64_t w[4];
};

// ones among payload bits [0, o) of a block, o in [0, 224]
__device__ __forceinline__ uint32_t block_rank(const RankBlock &b, uint32_t o)
{
    const uint32_t t = o + 32u;         // bit index inside the 256-bit block
    const uint32_t wq = t >> 6, r = t & 63u;
    const uint64_t part = (r ? ((1ULL << r) - 1ULL) : 0ULL);
    uint64_t w0 = b.w[0] & 0xFFFFFFFF00000000ULL;
    uint32_t c = 0;
    c += __popcll(wq > 0 ? w0 : (w0 & part));
    c += __popcll(wq > 1 ? b.w[1] : (wq == 1 ? (b.w[1] & part) : 0ULL));
    c += __popcll(wq > 2 ? b.w[2] : (wq == 2 ? (b.w[2] & part) : 0ULL));
    c += __popcll(wq > 3 ? b.w[3] : (wq == 3 ? (b.w[3] & part) : 0ULL));
    return c;
}
__device__ __forceinline__ uint32_t block_bit(const RankBlock &b, uint32_t o)
{
    const uint32_t t = o + 32u;
    return (uint32_t)(b.w[t >> 6] >> (t & 63u)) & 1u;
}
// one 256-bit read-only load (sm_100: LDG.E.256) of a 32-byte-aligned sector
__device__ __forceinline__ void ld_nc_256(const void *p, uint64_t &w0, uint64_t &w1, uint64_t &w2, uint64_t &w3)
{
    asm("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(w0), "=l"(w1), "=l"(w2), "=l"(w3) : "l"(p));
}
__device__ __forceinline__ RankBlock load_block(const RankBlock *p)
{
    RankBlock b;          // the whole block = one sector = one load instruction
    ld_nc_256(p, b.w[0], b.w[1], b.w[2], b.w[3]);
    return b;
}

// A level bit-vector as seen by device code.
struct BitVec {
    const RankBlock *blocks;
    const uint64_t *super;   // absolute ones before each superblock
    uint64_t len;            // bits
};

// rank1(i) = ones in bits [0, i), i in [0, len]
__device__ __forceinline__ uint64_t bv_rank(const BitVec &v, uint64_t i)
{
    const uint64_t blk = i / HKCSA_BLOCK_BITS;
    const uint32_t o = (uint32_t)(i - blk * HKCSA_BLOCK_BITS);
    const RankBlock b = load_block(v.blocks + blk);
    return v.super[blk / HKCSA_SUPER_BLOCKS] + (uint32_t)(b.w[0] & 0xFFFFFFFFu) + block_rank(b, o);
}

}  // namespace hkcsa
