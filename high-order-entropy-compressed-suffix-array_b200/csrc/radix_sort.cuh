// radix_sort.cuh -- single-pass-per-digit ("onesweep") LSD radix sort for sm_100a.
//
// One kernel launch per 8-bit digit.  Each CTA takes a tile through a ticket
// counter, ranks its keys with warp-level match histograms, publishes its
// per-digit counts to a look-back table, resolves the per-digit exclusive
// prefix over the preceding tiles by decoupled look-back, stages the tile in
// shared memory in digit order and writes it out as coalesced runs.
// Digit histograms for ALL passes are accumulated up front (they do not depend
// on element order) by whichever kernel produces the keys.
//
// Used for: (uint64 key, uint32 value) pairs in suffix-array construction (K1)
// and byte keys ranked through a look-up table (node id per symbol) in
// wavelet-tree level construction (K3) and FMIndex.precompute_rank.
#pragma once
#include "common.cuh"

namespace hkcsa {

constexpr int RADIX = 256;
constexpr uint32_t LB_AGG = 1u << 30;   // tile aggregate available
constexpr uint32_t LB_INC = 2u << 30;   // inclusive prefix available
constexpr uint32_t LB_VAL = (1u << 30) - 1u;

constexpr int SORT64_TILE = 2048;   // smallest (uint64, uint32) tile of any onesweep64 variant: sizes the look-back table

constexpr int SORT8_THREADS = 256;
constexpr int SORT8_IPT = 32;
constexpr int SORT8_TILE = SORT8_THREADS * SORT8_IPT;      // 8192 bytes per CTA

// Adds one key's digits (8 bits each, `passes` of them from bit 0) to a CTA's
// shared histogram hist[pass][256].  When every lane of the warp holds the same
// digit (sorted high digits of later rounds) one lane adds the whole count;
// otherwise plain shared-memory atomics.  Must be called by all 32 lanes.
__device__ __forceinline__ void hist_add_key(uint32_t *s_hist, uint64_t key, int passes, bool valid)
{
    const uint32_t lane = lane_id();
    const uint32_t nvalid = __popc(__ballot_sync(0xffffffffu, valid));
    const uint64_t key0 = __shfl_sync(0xffffffffu, key, 0);
    for (int p = 0; p < passes; ++p) {
        const uint32_t d = (uint32_t)((key >> (8 * p)) & 0xFFu);
        const uint32_t d0 = (uint32_t)((key0 >> (8 * p)) & 0xFFu);
        // lanes are valid in a prefix (tail of the array), so lane 0 is valid whenever any lane is
        const bool uniform = __all_sync(0xffffffffu, !valid || d == d0);
        if (uniform) {
            if (lane == 0 && nvalid) atomicAdd(&s_hist[p * RADIX + d0], nvalid);
        } else if (valid) {
            atomicAdd(&s_hist[p * RADIX + d], 1u);
        }
    }
}
// one shared-memory atomic per digit, the digit taken from the key's 32-bit halves already scaled to a byte offset
// (shift + mask: two instructions per digit, no 64-bit shifts, no per-pass loop test)
template <int PASSES>
__device__ __forceinline__ void hist_add_digits(uint32_t *s_hist, uint64_t key)
{
    const uint32_t lo = (uint32_t)key, hi = (uint32_t)(key >> 32);
    char *base = reinterpret_cast<char *>(s_hist);
#pragma unroll
    for (int p = 0; p < PASSES; ++p) {
        const uint32_t half = p < 4 ? lo : hi;
        const int sh = 8 * (p & 3);
        const uint32_t off = (sh >= 2 ? (half >> (sh >= 2 ? sh - 2 : 0)) : (half << 2)) & 0x3FCu;      // digit * 4
        atomicAdd(reinterpret_cast<uint32_t *>(base + p * RADIX * 4 + off), 1u);
    }
}
// Same contract as hist_add_key, for keys in no particular order (round 0), PASSES known at compile time: one
// uniformity test on the whole key (runs of one symbol) instead of one per digit; otherwise plain shared-memory atomics.
template <int PASSES>
__device__ __forceinline__ void hist_add_key_unsorted(uint32_t *s_hist, uint64_t key, bool valid)
{
    const uint64_t key0 = __shfl_sync(0xffffffffu, key, 0);
    const bool uniform = __all_sync(0xffffffffu, !valid || key == key0);
    if (uniform) {
        const uint32_t nvalid = __popc(__ballot_sync(0xffffffffu, valid));
        if (lane_id() == 0 && nvalid)
            for (int p = 0; p < PASSES; ++p) atomicAdd(&s_hist[p * RADIX + ((uint32_t)(key0 >> (8 * p)) & 0xFFu)], nvalid);
    } else if (valid) {
        hist_add_digits<PASSES>(s_hist, key);
    }
}
__device__ __forceinline__ void hist_zero(uint32_t *s_hist, int passes)
{
    for (int i = threadIdx.x; i < passes * RADIX; i += blockDim.x) s_hist[i] = 0;
}
__device__ __forceinline__ void hist_flush(const uint32_t *s_hist, uint32_t *g_hist, int passes)
{
    for (int i = threadIdx.x; i < passes * RADIX; i += blockDim.x) {
        const uint32_t v = s_hist[i];
        if (v) atomicAdd(&g_hist[i], v);
    }
}

struct SortScratch {
    uint32_t *hist;      // [8][256] digit histograms (filled by the key producer)
    uint32_t *base;      // [8][256] exclusive digit offsets
    uint32_t *lookback;  // [tiles][256]
    uint32_t *ticket;    // [8] one per pass
    size_t lookback_words;
};

size_t sort_scratch_words(uint64_t n);   // uint32 words needed by carve_sort_scratch
SortScratch carve_sort_scratch(Carver &c, uint64_t n);

// Histogram-only pass over existing keys (when the producer could not fuse it).
cudaError_t radix_histogram_u64(const uint64_t *d_keys, uint32_t n, int passes, const SortScratch &s,
                                cudaStream_t st);
// Sorts on the low 8*passes bits.  s.hist must hold the digit histograms of the
// input.  Input in (k0,v0); (k1,v1) is the ping-pong buffer.  The result is in
// (k0,v0) when passes is even, (k1,v1) when odd.
// identity_vals: the input values are 0, 1, 2, ... and v0 is not read (it is still the ping-pong buffer).
cudaError_t radix_sort_pairs_u64(uint64_t *k0, uint32_t *v0, uint64_t *k1, uint32_t *v1, uint32_t n,
                                 int passes, const SortScratch &s, cudaStream_t st, bool identity_vals = false);
// Stable bucket partition of bytes: element x goes to bucket lut[x]; bucket b
// starts at d_bucket_base[b] in d_out.  No histogram pass (the caller knows the
// bucket sizes).  If d_pos_out != nullptr the source positions are carried as values.
cudaError_t radix_partition_bytes(const uint8_t *d_in, uint8_t *d_out, uint32_t *d_pos_out, uint32_t n,
                                  const uint8_t *d_lut, const uint32_t *d_bucket_base,
                                  const SortScratch &s, cudaStream_t st);

}  // namespace hkcsa
