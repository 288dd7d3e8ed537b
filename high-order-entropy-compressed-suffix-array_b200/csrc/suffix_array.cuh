// suffix_array.cuh -- pieces of K1 shared by the single-GPU builder (suffix_array.cu) and the distributed
// builder (dist_sa.cu): the order-preserving prefix code of round 0, the shared-memory bit-stream packer, and the
// segmented rank-update kernels.
#pragma once
#include "common.cuh"
#include "radix_sort.cuh"

namespace hkcsa {

constexpr int SEG_THREADS = 256;
constexpr int SEG_IPT = 8;
constexpr int SEG_TILE = SEG_THREADS * SEG_IPT;

struct CodeMap {
    uint16_t code[256];  // byte -> dense code + 1 (0 for bytes that do not occur)
};

// Order-preserving prefix code ("alphabetic code") over {past-the-end} + the symbols of the text: frequent
// symbols get short codes, and comparing two code streams bit by bit equals comparing the symbol strings,
// the end of the text being smallest.  The round-0 key of a suffix is the first `bits` bits of its code
// stream, so a key covers as many symbols as a fixed-width packing would in fewer bits -- fewer radix passes
// (DNA + '$': 21 symbols in 48 bits instead of 63).  Equal keys share at least bits / max_len symbols.
struct AlphaCode {
    uint32_t code[257];   // right-aligned; [256] = past the end
    uint8_t len[257];
};

// returns false when the code would be degenerate (then the caller keeps fixed-width codes)
bool build_alpha_code(const uint64_t *h_hist, AlphaCode &ac, double &avg_len, int &max_len);

// The same idea over k-GRAMS (single-GPU builder): the order-preserving prefix code is built over the B^k strings of k
// symbols (B = sigma + 1 digits: 0 = past the end, 1.. = the symbols in byte order; a gram is its base-B number, so
// numeric order = lexicographic order), weighted by a sampled k-gram histogram of the text.  The code stream of a
// suffix is code(gram at i) code(gram at i + k) ...: comparing streams still equals comparing suffixes, but a gram's
// code length follows the k-th order statistics of the text instead of the symbol frequencies -- DNA + '$' (k = 6):
// 1.9 instead of 2.25 bits per symbol, so 40 key bits cover what 48 did and round 0 sorts in five passes, not six; the
// 97-symbol English-like text (k = 2): six or seven passes instead of eight, and a free top byte for the BWT symbol.
constexpr int GRAM_MIN_K = 4;              // shorter grams see too little context to beat the per-symbol code
constexpr int GRAM_MAX_K = 8;
constexpr int GRAM_MAX_G = 1 << 17;        // table entries (B^k <= this)
constexpr int GRAM_PER_KEY = 16;           // a key takes at most this many grams (degenerate texts: one-bit codes)
constexpr int GRAM_LEN_BITS = 5;           // table entry = code word left-aligned in the upper 27 bits | len, len <= 27
constexpr int GRAM_MAX_LEN = 27;
constexpr uint32_t GRAM_LEN_MASK = (1u << GRAM_LEN_BITS) - 1u;
struct GramCode {
    const uint32_t *tab;   // [G]; nullptr = the per-symbol code (AlphaCode) is in use
    uint32_t k, B, G, pw;  // pw = B^(k-1)
    uint8_t digit[256];    // byte -> digit (bytes that do not occur: 0)
};

// first `bits` bits of the gram code stream of the suffix at `pos` (what sa_pack0_gram_kernel writes for it)
__device__ __forceinline__ uint64_t gram_key(const GramCode &gc, const uint8_t *__restrict__ text, uint64_t n, uint64_t pos,
                                             int bits)
{
    uint64_t acc = 0;
    int used = 0;
    for (int t = 0; t < GRAM_PER_KEY && used < bits; ++t) {
        uint32_t g = 0;
        for (uint32_t q = 0; q < gc.k; ++q) {
            const uint64_t p = pos + (uint64_t)t * gc.k + q;
            g = g * gc.B + (p < n ? (uint32_t)gc.digit[text[p]] : 0u);
        }
        const uint32_t c = __ldg(gc.tab + g);
        acc |= ((uint64_t)(c & ~GRAM_LEN_MASK) << 32) >> used;      // the code word sits left-aligned in the entry
        used += (int)(c & GRAM_LEN_MASK);
    }
    return acc >> (64 - bits);
}

// Round-0 parameters derived from the byte histogram of the text -- the single-GPU and the distributed builder
// must key suffixes identically.
struct Round0Plan {
    AlphaCode ac;
    GramCode gram;  // gram.tab != nullptr: keys come from the k-gram code
    uint32_t sigma;
    int b;          // fixed-width code size
    int max_len;    // longest code word
    int bits0;      // key width (multiple of 8, 16..64)
    int k0;         // symbols every key is guaranteed to cover
    int passes0;    // radix passes of round 0
};
// want_carry: the caller wants the BWT symbol carried in the top byte of the key (single-GPU builder with a BWT
// output): with HKCSA_CARRY56=1 a plan that would use all 64 bits settles for 56
void make_round0_plan(const uint64_t *h_hist, Round0Plan &p, bool want_carry = false);

// bwt[i] = text[SA[i] - 1] (wrapping to text[n - 1]) by the phased gather of bwt.cu
int bwt_gather(const uint8_t *d_text, const uint32_t *d_sa, uint64_t n, uint8_t *d_bwt, cudaStream_t st);

// first `bits` bits of the code stream of the suffix whose symbols are produced by next_sym(t), t = 0, 1, ...
template <typename NextSym>
__device__ __forceinline__ uint64_t alpha_pack(const uint32_t *s_code, const uint8_t *s_len, int bits, NextSym next_sym)
{
    uint64_t acc = 0;
    int used = 0;
    for (int t = 0; used < bits; ++t) {
        const uint32_t c = next_sym(t);
        const int L = s_len[c];
        acc |= ((uint64_t)s_code[c] << (64 - L)) >> used;   // bits beyond 64 fall off: the last code is truncated
        used += L;
    }
    return acc >> (64 - bits);
}

// ---------------------------------------------------------------- round 0: the bit-stream packer
constexpr int PACK_THREADS = 256;
constexpr int PACK_IPT = 8;
constexpr int PACK_TILE = PACK_THREADS * PACK_IPT;
constexpr int PACK_LOOK = 64;                                   // >= 64 bits of look-ahead (code words >= 1 bit)
constexpr int PACK_VT = PACK_THREADS + PACK_LOOK / PACK_IPT;    // "virtual threads" incl. the look-ahead groups
constexpr int PACK_MAX_LEN = 24;                                // build_alpha_code never exceeds it
constexpr int PACK_STREAM_WORDS = (PACK_TILE + PACK_LOOK) * PACK_MAX_LEN / 32 + 4;

// loads 8 consecutive symbols starting at g0 as (code << 8 | len) entries; returns the sum of the lengths
template <typename LenT>
__device__ __forceinline__ uint32_t pack_load8(const uint8_t *__restrict__ text, LenT n, uint64_t g0, bool aligned8,
                                               const uint32_t *s_tab, uint32_t cl[PACK_IPT])
{
    uint32_t total = 0;
    if (aligned8 && g0 + PACK_IPT <= n) {
        const uint2 v = __ldg(reinterpret_cast<const uint2 *>(text + g0));
#pragma unroll
        for (int e = 0; e < PACK_IPT; ++e) {
            const uint32_t c = ((e < 4 ? v.x : v.y) >> (8 * (e & 3))) & 0xFFu;
            cl[e] = s_tab[c];
            total += cl[e] & 0xFFu;
        }
    } else {
#pragma unroll
        for (int e = 0; e < PACK_IPT; ++e) {
            const uint64_t g = g0 + e;
            cl[e] = s_tab[g < n ? (uint32_t)text[g] : 256u];
            total += cl[e] & 0xFFu;
        }
    }
    return total;
}

// ORs the 8 code words into the stream from bit offset `off` (bit 0 = MSB of word 0) and records the bit
// offset of every symbol
__device__ __forceinline__ void pack_emit8(uint32_t *s_stream, uint16_t *s_off8, uint32_t off, const uint32_t cl[PACK_IPT])
{
    uint32_t wi = off >> 5;
    uint32_t nacc = off & 31u;
    uint64_t acc = 0;
    uint32_t o = off;
    uint32_t offs[PACK_IPT];
#pragma unroll
    for (int e = 0; e < PACK_IPT; ++e) {
        const uint32_t L = cl[e] & 0xFFu;
        offs[e] = o;
        o += L;
        acc |= (uint64_t)(cl[e] >> 8) << (64u - nacc - L);     // nacc + L <= 31 + 24
        nacc += L;
        if (nacc >= 32u) {
            atomicOr(&s_stream[wi++], (uint32_t)(acc >> 32));
            acc <<= 32;
            nacc -= 32u;
        }
    }
    if (nacc) atomicOr(&s_stream[wi], (uint32_t)(acc >> 32));
    uint4 q;
    q.x = offs[0] | (offs[1] << 16);
    q.y = offs[2] | (offs[3] << 16);
    q.z = offs[4] | (offs[5] << 16);
    q.w = offs[6] | (offs[7] << 16);
    *reinterpret_cast<uint4 *>(s_off8) = q;
}

cudaError_t byte_hist(const uint8_t *d_text, uint64_t n, uint64_t *d_hist, cudaStream_t st);

// 64-bit suffix ids of the distributed build (texts beyond 4 GB, n <= 2^40): id in the low 40 bits, then WIDE_X_BITS
// bits that continue the round-0 key (the code stream behind its 64 bits; 0 when the packer cannot provide them),
// the BWT symbol in the top byte
constexpr int WIDE_ID_BITS = 40;
constexpr int WIDE_X_BITS = 16;
constexpr uint64_t WIDE_ID_MASK = (1ull << WIDE_ID_BITS) - 1ull;
static_assert(WIDE_ID_BITS + WIDE_X_BITS <= 56, "the top byte carries the BWT symbol");

// ---------------------------------------------------------------- group-local refinement (suffix_array.cu)
constexpr int GS_MAX = 16;       // groups up to this size are ordered by one thread
constexpr int GS_DEPTH = 64;     // symbols compared beyond the depth the group shares
// keys[j] = (cgrp[j] << 32 | sub-group) with every group of at most GS_MAX elements reordered in place (cidx) by text
// comparison from `depth` on; ids64 != nullptr: cidx holds ordinals into ids64 (WIDE_ID_MASK = suffix id, then the key extension bits)
cudaError_t group_local_keys(uint32_t *cidx, const uint32_t *cgrp, uint32_t m, const uint8_t *text, uint64_t n,
                             uint64_t depth, const uint64_t *ids64, uint64_t *keys, cudaStream_t st);

// ---------------------------------------------------------------- segmented rank update (suffix_array.cu)
// mask = the key bits that were sorted on; carried_out != nullptr: byte j = top byte of sorted key j
__global__ void __launch_bounds__(SEG_THREADS) seg_reduce_kernel(const uint64_t *__restrict__ skey, uint32_t m, uint32_t *__restrict__ agg_head,
                                  uint32_t *__restrict__ agg_keep, uint16_t *__restrict__ flags, uint64_t mask,
                                  uint8_t *__restrict__ carried_out);
__global__ void __launch_bounds__(1024) seg_scan_kernel(uint32_t *__restrict__ agg_head, uint32_t *__restrict__ agg_keep, uint32_t tiles,
                                uint32_t *__restrict__ out_m);
__global__ void __launch_bounds__(SEG_THREADS, 6) seg_apply_kernel(const uint16_t *__restrict__ flags, const uint32_t *__restrict__ sidx,
                                 const uint32_t *__restrict__ pos /* nullptr = identity */, uint32_t m,
                                 const uint32_t *__restrict__ carry_head, const uint32_t *__restrict__ carry_keep,
                                 uint32_t *__restrict__ sa, uint32_t *__restrict__ rank, uint32_t *__restrict__ cpos,
                                 uint32_t *__restrict__ cidx, uint32_t *__restrict__ cgrp, bool write_sa, bool scatter_all,
                                 const uint8_t *__restrict__ text, uint32_t n, uint8_t *__restrict__ bwt);

}  // namespace hkcsa
