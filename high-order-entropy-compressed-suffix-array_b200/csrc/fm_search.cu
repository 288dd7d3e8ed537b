// fm_search.cu -- K4: batched FM-index backward search (count) and locate.
//
// Replaces EnhancedFMIndex.find_range / .rank / .find (reference
// csa/enhanced_fm_index.py:15-40).  occ[c][i] of build_occ (utils/utils.py:
// 26-32) is answered by a rank walk over the wavelet tree; C[] and the node
// tables live in shared memory; every rank step reads one 32-byte block.
#include "common.cuh"
#include "prof.cuh"
#include "radix_sort.cuh"
#include "wavelet.cuh"

namespace hkcsa {

// wavelet.cu: mark bit-vector + directory + samples of the sampled suffix array in one pass over d_sa
int build_markvector(const uint32_t *d_sa, uint64_t n, uint32_t rate, RankBlock *d_blocks, uint64_t *d_super,
                     uint32_t *d_select, uint32_t *d_agg, uint64_t *d_carry, uint64_t *d_ones, uint32_t *d_state,
                     uint32_t *d_samples, cudaStream_t st);
int build_markvector64(const uint64_t *d_sa, uint64_t n, uint32_t rate, RankBlock *d_blocks, uint64_t *d_super,
                       uint32_t *d_select, uint32_t *d_agg, uint64_t *d_carry, uint64_t *d_ones, uint32_t *d_state,
                       uint32_t *d_samples, cudaStream_t st);

constexpr int COUNT_THREADS = 256;

// One lane per pattern.  A warp owns a contiguous chunk of the batch and refills a lane as soon as
// its pattern ends (last symbol consumed or range empty), so ragged lengths (8-64) and early misses do
// not idle the warp: every iteration each busy lane consumes ONE symbol -- both boundaries of the SA
// range walk the wavelet tree together.  find_range (csa/enhanced_fm_index.py:21-32) in half-open form:
//   l = 0, r = n;  per symbol from the end: l = C[c] + occ(c, l), r = C[c] + occ(c, r);
//   l >= r -> (-1, -1).  Result (l, r-1).
template <bool PEERS>
__global__ void __launch_bounds__(COUNT_THREADS)
fm_count_kernel(WtDev wt, const uint8_t *__restrict__ pat, const int64_t *__restrict__ off, uint64_t P,
                int64_t *__restrict__ out_lo, int64_t *__restrict__ out_hi, const uint2 *__restrict__ kmer, uint32_t kk,
                PeerOut po)
{
    __shared__ WtSmem s;
    wt_smem_load(s, wt);
    __syncthreads();
    const uint32_t lane = lane_id();
    const uint32_t n = (uint32_t)wt.n;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint64_t chunk = (P + nwarps - 1) / nwarps;
    uint64_t next = min(P, warp * chunk);
    const uint64_t end = min(P, next + chunk);

    int64_t p = -1;            // pattern owned by this lane, -1 = idle
    int64_t k = 0, b = 0;      // next symbol to consume, first symbol of the pattern
    uint32_t l = 0, r = 0;
    while (true) {
        // ---- refill idle lanes with the next patterns of the chunk, in lane order
        const uint32_t idle = __ballot_sync(0xffffffffu, p < 0);
        if (idle) {
            const uint64_t mine = next + __popc(idle & lanemask_lt());
            if (p < 0 && mine < end) {
                p = (int64_t)mine;
                b = off[mine];
                k = off[mine + 1] - 1;
                l = 0;
                r = n;
                // jump table: the SA range of the pattern's last kk symbols was precomputed (by this same
                // search) for every kk-mer over the alphabet; skip those steps
                if (kmer != nullptr && k - b + 1 >= (int64_t)kk) {
                    uint32_t id = 0;
                    bool known = true;
                    for (uint32_t t = 0; t < kk; ++t) {
                        const uint32_t code = s.code_of_sym[pat[k - kk + 1 + t]];
                        known = known && code != 0xFFFFu;
                        id = id * wt.sigma + (known ? code : 0u);
                    }
                    if (known) {
                        const uint2 e2 = __ldg(kmer + id);
                        l = e2.x;
                        r = e2.y;          // l >= r: the kk-mer does not occur
                        k -= kk;
                        if (l >= r) { l = 1; r = 0; k = b - 1; }   // finishes below as a miss
                    }
                }
            }
            next += __popc(idle);
            if (idle == 0xffffffffu && __ballot_sync(0xffffffffu, p >= 0) == 0) break;
        }
        if (p < 0) continue;
        // ---- one backward-search step (or finish an exhausted / empty pattern)
        bool done = k < b, miss = l >= r;
        if (!done) {
            const uint32_t code = s.code_of_sym[pat[k]];
            if (code == 0xFFFFu) { miss = true; }          // symbol absent: rank 0, C 0 -> empty range
            else {
                if (wt.sigma > 1) wt_rank_code2(s, wt, code, l, r);
                const uint32_t c0 = s.C[code];
                l += c0;
                r += c0;
                miss = l >= r;
            }
            --k;
            done = miss || k < b;
        }
        if (done) {
            put_range<PEERS>(po, out_lo, out_hi, p, miss ? -1 : (int64_t)l, miss ? -1 : (int64_t)r - 1);
            p = -1;
        }
    }
}

// rows[off[p] + k] = lo[p] + k: one warp per pattern
__global__ void expand_ranges_kernel(const int64_t *__restrict__ lo, const int64_t *__restrict__ hi,
                                     const int64_t *__restrict__ out_off, uint64_t P, uint32_t *__restrict__ rows)
{
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = lane_id();
    for (uint64_t p = warp; p < P; p += ((uint64_t)gridDim.x * blockDim.x) >> 5) {
        const int64_t l = lo[p], h = hi[p];
        if (l < 0) continue;
        const int64_t o = out_off[p];
        for (int64_t k = lane; k <= h - l; k += 32) rows[o + k] = (uint32_t)(l + k);
    }
}

// 64-bit rows for indexes beyond 2^32 rows
__global__ void expand_ranges64_kernel(const int64_t *__restrict__ lo, const int64_t *__restrict__ hi,
                                       const int64_t *__restrict__ out_off, uint64_t P, uint64_t *__restrict__ rows)
{
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = lane_id();
    for (uint64_t p = warp; p < P; p += ((uint64_t)gridDim.x * blockDim.x) >> 5) {
        const int64_t l = lo[p], h = hi[p];
        if (l < 0) continue;
        const int64_t o = out_off[p];
        for (int64_t k = lane; k <= h - l; k += 32) rows[o + k] = (uint64_t)(l + k);
    }
}

// ---------------------------------------------------------------- results of a sharded batch to every rank
// (lo, hi) of this rank's slice -> packed {lo: low 32 bits, count: high 32 bits} (a miss is lo = 0xFFFFFFFF, count 0)
// stored at [base + p] of the result array of EVERY rank: 8 bytes per pattern instead of the 16 of (lo, hi) as int64.
// A thread packs two patterns and issues one 16-byte store per destination; with a multicast address (NVSwitch)
// ONE multimem.st reaches all ranks.
struct PushDest {
    uint64_t *out[HKCSA_MAX_PEERS];
    uint64_t *mc;            // multicast address of the same array, or nullptr
    uint32_t n;
};
__device__ __forceinline__ uint64_t pack_range(int64_t l, int64_t h)
{
    return l < 0 ? 0x00000000FFFFFFFFull : ((uint64_t)(uint32_t)l | ((uint64_t)(uint32_t)(h - l + 1) << 32));
}
__device__ __forceinline__ void multimem_st16(uint64_t *p, uint64_t a, uint64_t b)
{
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p),
                 "f"(__uint_as_float((uint32_t)a)), "f"(__uint_as_float((uint32_t)(a >> 32))),
                 "f"(__uint_as_float((uint32_t)b)), "f"(__uint_as_float((uint32_t)(b >> 32))) : "memory");
}
__device__ __forceinline__ void multimem_st8(uint64_t *p, uint64_t a)
{
    asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1, %2};" ::"l"(p),
                 "f"(__uint_as_float((uint32_t)a)), "f"(__uint_as_float((uint32_t)(a >> 32))) : "memory");
}
__global__ void __launch_bounds__(256)
ranges_push_kernel(const int64_t *__restrict__ lo, const int64_t *__restrict__ hi, uint64_t P, uint64_t base, PushDest pd)
{
    const uint64_t first_pair = base >> 1, pairs = ((base + P + 1) >> 1) - first_pair;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < pairs; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t g0 = (first_pair + i) << 1;
        const bool v0 = g0 >= base, v1 = g0 + 1 < base + P;
        const uint64_t a = v0 ? pack_range(lo[g0 - base], hi[g0 - base]) : 0ull;
        const uint64_t b = v1 ? pack_range(lo[g0 + 1 - base], hi[g0 + 1 - base]) : 0ull;
        if (pd.mc) {
            if (v0 && v1) multimem_st16(pd.mc + g0, a, b);
            else if (v0) multimem_st8(pd.mc + g0, a);
            else multimem_st8(pd.mc + g0 + 1, b);
        } else {
            for (uint32_t r = 0; r < pd.n; ++r) {
                if (v0 && v1) *reinterpret_cast<ulonglong2 *>(pd.out[r] + g0) = make_ulonglong2(a, b);
                else if (v0) pd.out[r][g0] = a;
                else pd.out[r][g0 + 1] = b;
            }
        }
    }
}
__global__ void ranges_unpack_kernel(const uint64_t *__restrict__ packed, uint64_t P, int64_t *__restrict__ lo,
                                     int64_t *__restrict__ hi)
{
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t v = packed[p];
        const uint32_t l = (uint32_t)v, c = (uint32_t)(v >> 32);
        const bool miss = c == 0;
        lo[p] = miss ? -1 : (int64_t)l;
        hi[p] = miss ? -1 : (int64_t)l + c - 1;
    }
}

__global__ void gather_u32_kernel(const uint32_t *__restrict__ src, const uint32_t *__restrict__ rows, uint64_t m,
                                  uint32_t *__restrict__ out)
{
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < m) out[q] = src[rows[q]];
}

// position of row j: walk LF until a marked row, pos = sample * rate + steps.
// LF(j) = C[c] + occ(c, j) with c = bwt[j], both from one wavelet descent.
__global__ void __launch_bounds__(256)
locate_rows_kernel(WtDev wt, BitVec marks, const uint32_t *__restrict__ samples, uint32_t rate,
                   const uint32_t *__restrict__ rows, uint64_t m, uint32_t *__restrict__ out)
{
    __shared__ WtSmem s;
    wt_smem_load(s, wt);
    __syncthreads();
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= m) return;
    uint32_t j = rows[q];
    uint32_t steps = 0;
    while (true) {
        const uint64_t g = j / HKCSA_BLOCK_BITS;
        const uint32_t o = j - (uint32_t)g * HKCSA_BLOCK_BITS;
        const RankBlock b = load_block(marks.blocks + g);
        if (block_bit(b, o)) {
            const uint64_t r = marks.super[g / HKCSA_SUPER_BLOCKS] + (uint32_t)(b.w[0] & 0xFFFFFFFFu) + block_rank(b, o);
            out[q] = samples[r] * rate + steps;
            return;
        }
        // a sampled row is at most rate-1 LF steps away when the sentinel is unique; on a text that already holds
        // the sentinel LF is a permutation whose cycles may miss every mark: stop, report HKCSA_NO_POSITION
        if (steps >= rate) { out[q] = HKCSA_NO_POSITION; return; }
        uint32_t occ;
        const uint32_t code = wt_access_rank(s, wt, j, occ);
        j = s.C[code] + occ;
        ++steps;
    }
}

// every kk-mer over the index alphabet, in lexicographic (code) order, as patterns of length kk
__global__ void kmer_patterns_kernel(WtDev wt, uint32_t kk, uint64_t count, uint8_t *__restrict__ pat,
                                     int64_t *__restrict__ off)
{
    const uint64_t id = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id > count) return;
    off[id] = (int64_t)(id * kk);
    if (id == count) return;
    uint64_t v = id;
    for (int t = (int)kk - 1; t >= 0; --t) {
        pat[id * kk + t] = wt.tab->sym_of_code[v % wt.sigma];
        v /= wt.sigma;
    }
}
__global__ void kmer_table_fill_kernel(const int64_t *__restrict__ lo, const int64_t *__restrict__ hi, uint64_t count,
                                       uint2 *__restrict__ table)
{
    const uint64_t id = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= count) return;
    const int64_t l = lo[id], h = hi[id];
    table[id] = (l < 0) ? make_uint2(0u, 0u) : make_uint2((uint32_t)l, (uint32_t)(h + 1));
}

__global__ void symbol_lut_kernel(uint8_t *lut, uint32_t *base, const uint64_t *__restrict__ hist, uint64_t *start)
{
    // single CTA of 256 threads: bucket = byte value, base = exclusive prefix of the histogram
    __shared__ uint64_t s_h[256];
    const uint32_t t = threadIdx.x;
    s_h[t] = hist[t];
    __syncthreads();
    uint64_t pre = 0;
    for (uint32_t c = 0; c < t; ++c) pre += s_h[c];
    lut[t] = (uint8_t)t;
    base[t] = (uint32_t)pre;
    start[t] = pre;
    if (t == 255) start[256] = pre + s_h[255];
}

}  // namespace hkcsa

using namespace hkcsa;

extern "C" int hkcsa_count_batch(const void *d_blob, const hkcsa_wt_plan *h_plan, const uint8_t *d_pat,
                                 const int64_t *d_off, uint64_t P, int64_t *d_lo, int64_t *d_hi, void *stream)
{
    return hkcsa_count_batch_kmer(d_blob, h_plan, nullptr, 0, d_pat, d_off, P, d_lo, d_hi, stream);
}

// number of symbols a jump table covers for an alphabet of sigma symbols: the largest k with sigma^k <= 2^21
extern "C" uint32_t hkcsa_kmer_k(uint32_t sigma)
{
    if (sigma < 2) return 0;
    uint32_t k = 0;
    uint64_t v = 1;
    while (k < 16 && v * sigma <= (1ull << 21)) { v *= sigma; ++k; }
    return k;
}
extern "C" uint64_t hkcsa_kmer_entries(uint32_t sigma, uint32_t k)
{
    uint64_t v = 1;
    for (uint32_t t = 0; t < k; ++t) v *= sigma;
    return v;
}
extern "C" size_t hkcsa_kmer_scratch_bytes(uint32_t sigma, uint32_t k)
{
    const uint64_t cnt = hkcsa_kmer_entries(sigma, k);
    Carver c(nullptr);
    c.take<uint8_t>(cnt * k + 16);
    c.take<int64_t>(cnt + 1);
    c.take<int64_t>(cnt);
    c.take<int64_t>(cnt);
    return c.total();
}
// d_table: uint2[sigma^k] = half-open SA range (l, r) of every k-mer (l >= r: does not occur), computed by the
// count kernel itself, so searches that start from the table are bit-identical to full searches.
extern "C" int hkcsa_kmer_table_build(const void *d_blob, const hkcsa_wt_plan *h_plan, uint32_t k, void *d_table,
                                      void *d_scratch, size_t scratch_bytes, void *stream)
{
    HK_REQUIRE(h_plan && d_blob && d_table && d_scratch, HKCSA_EINVAL, "null pointer");
    HK_REQUIRE(k >= 1 && k <= 16 && h_plan->sigma >= 2, HKCSA_EINVAL, "bad k or alphabet");
    const uint64_t cnt = hkcsa_kmer_entries(h_plan->sigma, k);
    HK_REQUIRE(cnt <= (1ull << 22), HKCSA_EINVAL, "table too large");
    Carver c(d_scratch);
    uint8_t *d_pat = c.take<uint8_t>(cnt * k + 16);
    int64_t *d_off = c.take<int64_t>(cnt + 1);
    int64_t *d_lo = c.take<int64_t>(cnt);
    int64_t *d_hi = c.take<int64_t>(cnt);
    HK_REQUIRE(c.total() <= scratch_bytes, HKCSA_ESCRATCH, "k-mer scratch too small");
    cudaStream_t st = as_stream(stream);
    WtDev wt = make_wt_dev(d_blob, h_plan);
    kmer_patterns_kernel<<<(uint32_t)((cnt + 256) / 256), 256, 0, st>>>(wt, k, cnt, d_pat, d_off);
    HK_LAUNCH_CHECK();
    int rc = hkcsa_count_batch_kmer(d_blob, h_plan, nullptr, 0, d_pat, d_off, cnt, d_lo, d_hi, stream);
    if (rc != HKCSA_OK) return rc;
    kmer_table_fill_kernel<<<(uint32_t)((cnt + 255) / 256), 256, 0, st>>>(d_lo, d_hi, cnt, static_cast<uint2 *>(d_table));
    HK_LAUNCH_CHECK();
    return HKCSA_OK;
}

extern "C" int hkcsa_count_batch_kmer(const void *d_blob, const hkcsa_wt_plan *h_plan, const void *d_kmer_table,
                                      uint32_t k, const uint8_t *d_pat, const int64_t *d_off, uint64_t P,
                                      int64_t *d_lo, int64_t *d_hi, void *stream)
{
    HK_REQUIRE(h_plan && d_blob, HKCSA_EINVAL, "null pointer");
    if (P == 0) return HKCSA_OK;
    HK_REQUIRE(d_off && d_lo && d_hi, HKCSA_EINVAL, "null pointer");
    HK_REQUIRE(h_plan->n >= 1, HKCSA_EINVAL, "empty index");
    PeerOut po;
    memset(&po, 0, sizeof(po));
    return count_wt_launch(make_wt_dev(d_blob, h_plan), static_cast<const uint2 *>(d_kmer_table), k, d_pat, d_off, P,
                           d_lo, d_hi, po, as_stream(stream));
}

int hkcsa::count_wt_launch(const WtDev &wt, const uint2 *kmer, uint32_t k, const uint8_t *d_pat, const int64_t *d_off,
                           uint64_t P, int64_t *d_lo, int64_t *d_hi, const PeerOut &po, cudaStream_t st)
{
    const int blocks = (int)std::min<uint64_t>((P + COUNT_THREADS - 1) / COUNT_THREADS, (uint64_t)num_sms() * 8);
    prof::Scope ps(st, prof::COUNT, 0);
    if (po.n) fm_count_kernel<true><<<blocks, COUNT_THREADS, 0, st>>>(wt, d_pat, d_off, P, d_lo, d_hi, kmer, kmer ? k : 0u, po);
    else fm_count_kernel<false><<<blocks, COUNT_THREADS, 0, st>>>(wt, d_pat, d_off, P, d_lo, d_hi, kmer, kmer ? k : 0u, po);
    HK_LAUNCH_CHECK();
    return HKCSA_OK;
}

extern "C" int hkcsa_ssa_plan_make(uint64_t n, uint32_t rate, hkcsa_ssa_plan *p)
{
    HK_REQUIRE(p && rate >= 1, HKCSA_EINVAL, "bad argument");
    HK_REQUIRE(n <= HKCSA_MAX_N, HKCSA_ERANGE, "n exceeds HKCSA_MAX_N");
    memset(p, 0, sizeof(*p));
    p->n = n;
    p->rate = rate;
    p->n_samples = (n + rate - 1) / rate;
    uint64_t off = 0;
    p->off_blocks = off;
    off = align_up(off + rank_blocks_for(n) * sizeof(RankBlock), 256);
    p->off_super = off;
    off = align_up(off + super_for(n) * sizeof(uint64_t), 256);
    p->off_samples = off;
    off = align_up(off + (p->n_samples + 1) * sizeof(uint32_t), 256);
    p->blob_bytes = off;
    Carver c(nullptr);
    const uint64_t tiles = rank_blocks_for(n) / 64 + 2;
    c.take<uint32_t>(tiles);
    c.take<uint64_t>(tiles);
    c.take<uint64_t>(8);
    c.take<uint32_t>(select_samples_for(n));
    c.take<uint32_t>(tiles + 1);
    p->scratch_bytes = c.total();
    return HKCSA_OK;
}

extern "C" int hkcsa_ssa_plan_make_slice(uint64_t m, uint32_t rate, uint64_t n_marks, hkcsa_ssa_plan *p)
{
    int rc = hkcsa_ssa_plan_make(m, rate, p);
    if (rc != HKCSA_OK) return rc;
    HK_REQUIRE(n_marks <= m, HKCSA_EINVAL, "more marks than rows");
    p->n_samples = n_marks;
    p->off_samples = p->off_super + align_up(super_for(m) * sizeof(uint64_t), 256);
    p->blob_bytes = p->off_samples + align_up((n_marks + 1) * sizeof(uint32_t), 256);
    return HKCSA_OK;
}

template <typename IdT>
static int ssa_build_t(const IdT *d_sa, const hkcsa_ssa_plan *p, void *d_blob, void *d_scratch, size_t scratch_bytes,
                       void *stream)
{
    HK_REQUIRE(p && d_blob && d_scratch && (d_sa || p->n == 0), HKCSA_EINVAL, "null pointer");
    HK_REQUIRE(p->scratch_bytes <= scratch_bytes, HKCSA_ESCRATCH, "sampled-SA scratch too small");
    HK_REQUIRE((reinterpret_cast<uintptr_t>(d_blob) & 31) == 0, HKCSA_EINVAL, "blob must be 32-byte aligned");
    cudaStream_t st = as_stream(stream);
    const uint64_t n = p->n;
    if (n == 0) return HKCSA_OK;
    Carver c(d_scratch);
    const uint64_t tiles = rank_blocks_for(n) / 64 + 2;
    uint32_t *d_agg = c.take<uint32_t>(tiles);
    uint64_t *d_carry = c.take<uint64_t>(tiles);
    uint64_t *d_ones = c.take<uint64_t>(8);
    uint32_t *d_sel = c.take<uint32_t>(select_samples_for(n));
    uint32_t *d_state = c.take<uint32_t>(tiles + 1);
    uint8_t *blob = static_cast<uint8_t *>(d_blob);
    // one pass: ids read once (+ the marked ones again from cache), mark blocks and samples written
    prof::Scope ps(st, prof::SSA_BUILD, n * sizeof(IdT) + n / 7 + (n / p->rate) * (4 + sizeof(IdT)));
    RankBlock *blocks = reinterpret_cast<RankBlock *>(blob + p->off_blocks);
    uint64_t *super = reinterpret_cast<uint64_t *>(blob + p->off_super);
    uint32_t *samples = reinterpret_cast<uint32_t *>(blob + p->off_samples);
    if constexpr (sizeof(IdT) == 8)
        return build_markvector64(d_sa, n, p->rate, blocks, super, d_sel, d_agg, d_carry, d_ones, d_state, samples, st);
    else
        return build_markvector(d_sa, n, p->rate, blocks, super, d_sel, d_agg, d_carry, d_ones, d_state, samples, st);
}

extern "C" int hkcsa_ssa_build(const uint32_t *d_sa, const hkcsa_ssa_plan *p, void *d_blob, void *d_scratch,
                               size_t scratch_bytes, void *stream)
{
    return ssa_build_t<uint32_t>(d_sa, p, d_blob, d_scratch, scratch_bytes, stream);
}

// suffix ids as uint64 (slices of a text beyond 4 GB); samples stay uint32 (id / rate must fit)
extern "C" int hkcsa_ssa_build64(const uint64_t *d_sa, const hkcsa_ssa_plan *p, void *d_blob, void *d_scratch,
                                 size_t scratch_bytes, void *stream)
{
    return ssa_build_t<uint64_t>(d_sa, p, d_blob, d_scratch, scratch_bytes, stream);
}

extern "C" int hkcsa_expand_ranges(const int64_t *d_lo, const int64_t *d_hi, const int64_t *d_out_off, uint64_t P,
                                   uint32_t *d_rows, void *stream)
{
    if (P == 0) return HKCSA_OK;
    HK_REQUIRE(d_lo && d_hi && d_out_off, HKCSA_EINVAL, "null pointer");
    const int blocks = (int)std::min<uint64_t>((P * 32 + 255) / 256, (uint64_t)num_sms() * 16);
    expand_ranges_kernel<<<blocks, 256, 0, as_stream(stream)>>>(d_lo, d_hi, d_out_off, P, d_rows);
    HK_LAUNCH_CHECK();
    return HKCSA_OK;
}

extern "C" int hkcsa_expand_ranges64(const int64_t *d_lo, const int64_t *d_hi, const int64_t *d_out_off, uint64_t P,
                                     uint64_t *d_rows, void *stream)
{
    if (P == 0) return HKCSA_OK;
    HK_REQUIRE(d_lo && d_hi && d_out_off, HKCSA_EINVAL, "null pointer");
    const int blocks = (int)std::min<uint64_t>((P * 32 + 255) / 256, (uint64_t)num_sms() * 16);
    expand_ranges64_kernel<<<blocks, 256, 0, as_stream(stream)>>>(d_lo, d_hi, d_out_off, P, d_rows);
    HK_LAUNCH_CHECK();
    return HKCSA_OK;
}

extern "C" int hkcsa_gather_u32(const uint32_t *d_src, const uint32_t *d_rows, uint64_t m, uint32_t *d_out,
                                void *stream)
{
    if (m == 0) return HKCSA_OK;
    HK_REQUIRE(d_src && d_rows && d_out, HKCSA_EINVAL, "null pointer");
    gather_u32_kernel<<<(uint32_t)((m + 255) / 256), 256, 0, as_stream(stream)>>>(d_src, d_rows, m, d_out);
    HK_LAUNCH_CHECK();
    return HKCSA_OK;
}

extern "C" int hkcsa_locate_rows(const void *d_wt_blob, const hkcsa_wt_plan *h_plan, const void *d_ssa_blob,
                                 const hkcsa_ssa_plan *h_ssa, const uint32_t *d_rows, uint64_t m,
                                 uint32_t *d_out_pos, void *stream)
{
    if (m == 0) return HKCSA_OK;
    HK_REQUIRE(d_wt_blob && h_plan && d_ssa_blob && h_ssa && d_rows && d_out_pos, HKCSA_EINVAL, "null pointer");
    cudaStream_t st = as_stream(stream);
    WtDev wt = make_wt_dev(d_wt_blob, h_plan);
    const uint8_t *sb = static_cast<const uint8_t *>(d_ssa_blob);
    BitVec marks;
    marks.blocks = reinterpret_cast<const RankBlock *>(sb + h_ssa->off_blocks);
    marks.super = reinterpret_cast<const uint64_t *>(sb + h_ssa->off_super);
    marks.len = h_ssa->n;
    prof::Scope ps(st, prof::LOCATE, 0);
    locate_rows_kernel<<<(uint32_t)((m + 255) / 256), 256, 0, st>>>(
        wt, marks, reinterpret_cast<const uint32_t *>(sb + h_ssa->off_samples), h_ssa->rate, d_rows, m, d_out_pos);
    HK_LAUNCH_CHECK();
    return HKCSA_OK;
}

extern "C" int hkcsa_symbol_positions(const uint8_t *d_bwt, uint64_t n, uint32_t *d_pos, uint64_t *d_start,
                                      void *d_scratch, size_t scratch_bytes, void *stream)
{
    HK_REQUIRE(d_start && d_scratch, HKCSA_EINVAL, "null pointer");
    HK_REQUIRE(n <= HKCSA_MAX_N, HKCSA_ERANGE, "n exceeds HKCSA_MAX_N");
    cudaStream_t st = as_stream(stream);
    Carver c(d_scratch);
    uint64_t *d_hist = c.take<uint64_t>(256);
    uint8_t *d_lut = c.take<uint8_t>(256);
    uint32_t *d_base = c.take<uint32_t>(256);
    uint8_t *d_sorted = c.take<uint8_t>(n + 16);
    SortScratch ss = carve_sort_scratch(c, n);
    HK_REQUIRE(c.total() <= scratch_bytes, HKCSA_ESCRATCH, "scratch too small");
    HK_CUDA(hkcsa_byte_hist(d_bwt, n, d_hist, stream) == HKCSA_OK ? cudaSuccess : cudaErrorUnknown);
    symbol_lut_kernel<<<1, 256, 0, st>>>(d_lut, d_base, d_hist, d_start);
    HK_LAUNCH_CHECK();
    if (n == 0) return HKCSA_OK;
    HK_REQUIRE(d_bwt && d_pos, HKCSA_EINVAL, "null pointer");
    HK_CUDA(radix_partition_bytes(d_bwt, d_sorted, d_pos, (uint32_t)n, d_lut, d_base, ss, st));
    return HKCSA_OK;
}

extern "C" size_t hkcsa_symbol_positions_scratch_bytes(uint64_t n)
{
    Carver c(nullptr);
    c.take<uint64_t>(256);
    c.take<uint8_t>(256);
    c.take<uint32_t>(256);
    c.take<uint8_t>(n + 16);
    carve_sort_scratch(c, n);
    return c.total();
}

extern "C" int hkcsa_ranges_push_peers(const int64_t *d_lo, const int64_t *d_hi, uint64_t P, uint64_t out_base,
                                       uint32_t n_peers, const uint64_t *h_peer_out, uint64_t multicast_out, void *stream)
{
    if (P == 0) return HKCSA_OK;
    HK_REQUIRE(d_lo && d_hi && (h_peer_out || multicast_out), HKCSA_EINVAL, "null pointer");
    HK_REQUIRE(n_peers <= HKCSA_MAX_PEERS && (n_peers >= 1 || multicast_out), HKCSA_EINVAL, "n_peers must be in [1, HKCSA_MAX_PEERS]");
    PushDest pd;
    memset(&pd, 0, sizeof(pd));
    pd.n = n_peers;
    pd.mc = reinterpret_cast<uint64_t *>(static_cast<uintptr_t>(multicast_out));
    for (uint32_t r = 0; r < n_peers && h_peer_out; ++r) {
        HK_REQUIRE((h_peer_out[r] & 15) == 0 && h_peer_out[r], HKCSA_EINVAL, "peer arrays must be 16-byte aligned");
        pd.out[r] = reinterpret_cast<uint64_t *>(static_cast<uintptr_t>(h_peer_out[r]));
    }
    HK_REQUIRE((multicast_out & 15) == 0, HKCSA_EINVAL, "multicast array must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    const uint64_t pairs = (P + 2) / 2;
    const int blocks = (int)std::min<uint64_t>((pairs + 255) / 256, (uint64_t)num_sms() * 8);
    // algorithmic bytes: 16 B read, 8 B stored to each of the other ranks over NVLink
    prof::Scope ps(st, prof::OTHER, P * (16 + 8ull * std::max(1u, n_peers)));
    ranges_push_kernel<<<blocks, 256, 0, st>>>(d_lo, d_hi, P, out_base, pd);
    HK_LAUNCH_CHECK();
    return HKCSA_OK;
}

extern "C" int hkcsa_ranges_unpack(const uint64_t *d_packed, uint64_t P, int64_t *d_lo, int64_t *d_hi, void *stream)
{
    if (P == 0) return HKCSA_OK;
    HK_REQUIRE(d_packed && d_lo && d_hi, HKCSA_EINVAL, "null pointer");
    const int blocks = (int)std::min<uint64_t>((P + 255) / 256, (uint64_t)num_sms() * 8);
    ranges_unpack_kernel<<<blocks, 256, 0, as_stream(stream)>>>(d_packed, P, d_lo, d_hi);
    HK_LAUNCH_CHECK();
    return HKCSA_OK;
}
