// multi_slice.cu -- FM backward search over a BWT that was built in slices (distributed build, BASELINE
// config 5): slice s covers global rows [start[s], start[s+1]) and carries its own wavelet tree (built over its
// BWT slice with libhkcsa K3) and its own sampled-SA marks/samples.  occ(c, i) over the whole BWT =
// (occurrences of c in the slices before the one holding i) + a rank walk inside that slice, so the recurrences
// of EnhancedFMIndex.find_range (reference csa/enhanced_fm_index.py:21-32) run unchanged on global row numbers.
// Every GPU holds all slices (they are all-gathered after the build); patterns are sharded as in config 4.
#include "common.cuh"
#include "prof.cuh"
#include "wavelet.cuh"

namespace hkcsa {

struct MultiDesc {
    WtDev slice[HKCSA_MAX_SLICES];
    BitVec marks[HKCSA_MAX_SLICES];
    const uint32_t *samples[HKCSA_MAX_SLICES];
    uint64_t start[HKCSA_MAX_SLICES + 1];
    uint64_t cum[HKCSA_MAX_SLICES + 1][256];   // occurrences of byte c in slices < s  ([S] = total)
    uint64_t C[256];                           // symbols of the whole text smaller than byte c
    uint32_t S;
    uint32_t rate;                             // 0 = no sampled SA
    uint64_t n;
};

__device__ __forceinline__ uint32_t ms_slice_of(const MultiDesc *d, uint64_t i)
{
    uint32_t s = 0;
    for (uint32_t k = 1; k < d->S; ++k) s += (i >= d->start[k]) ? 1u : 0u;
    return s;
}

// rank walk with the node tables read from global memory (they are a few KB per slice and stay in L1/L2)
__device__ __forceinline__ uint32_t ms_rank_local(const WtDev &w, uint32_t code, uint32_t p)
{
    const WtTables *t = w.tab;
    const uint32_t dep = t->depth[code];
    for (uint32_t l = 0; l < dep; ++l) {
        const uint32_t ns = __ldg(&t->node_start[l][code]);
        const uint32_t start = ns & NODE_START_MASK;
        const uint32_t r1 = (uint32_t)bv_rank(w.level[l], (uint64_t)start + p) - __ldg(&t->node_ones[l][code]);
        p = (ns & NODE_BIT_FLAG) ? r1 : (p - r1);
    }
    return p;
}

// occurrences of `byte` in global rows [0, i)
__device__ __forceinline__ uint64_t ms_rank(const MultiDesc *d, uint32_t byte, uint64_t i)
{
    const uint32_t s = ms_slice_of(d, i);
    const WtDev &w = d->slice[s];
    const uint32_t code = __ldg(&w.tab->code_of_sym[byte]);
    uint32_t local = 0;
    if (code != 0xFFFFu) {
        const uint32_t p = (uint32_t)(i - d->start[s]);
        local = (w.sigma == 1) ? p : ms_rank_local(w, code, p);
    }
    return d->cum[s][byte] + local;
}

__global__ void __launch_bounds__(256)
ms_count_kernel(const MultiDesc *__restrict__ d, const uint8_t *__restrict__ pat, const int64_t *__restrict__ off,
                uint64_t P, int64_t *__restrict__ out_lo, int64_t *__restrict__ out_hi)
{
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const int64_t b = off[p], e = off[p + 1];
    uint64_t l = 0, r = d->n;
    bool miss = false;
    for (int64_t k = e - 1; k >= b; --k) {
        const uint32_t c = pat[k];
        if (d->cum[d->S][c] == 0) { miss = true; break; }       // symbol absent from the text
        l = d->C[c] + ms_rank(d, c, l);
        r = d->C[c] + ms_rank(d, c, r);
        if (l >= r) { miss = true; break; }
    }
    out_lo[p] = miss ? -1 : (int64_t)l;
    out_hi[p] = miss ? -1 : (int64_t)r - 1;
}

// access + rank inside one slice, tables from global memory
__device__ __forceinline__ uint32_t ms_access_rank(const WtDev &w, uint32_t i, uint32_t &occ)
{
    const WtTables *t = w.tab;
    uint32_t lo = 0, hi = w.sigma, p = i;
    for (uint32_t l = 0; hi - lo > 1; ++l) {
        const uint32_t mid = lo + (hi - lo) / 2;
        const uint32_t start = __ldg(&t->node_start[l][lo]) & NODE_START_MASK;
        const BitVec &v = w.level[l];
        const uint64_t ia = (uint64_t)start + p;
        const uint64_t ba = ia / HKCSA_BLOCK_BITS;
        const uint32_t o = (uint32_t)(ia - ba * HKCSA_BLOCK_BITS);
        const RankBlock q = load_block(v.blocks + ba);
        const uint32_t r1 = (uint32_t)(v.super[ba / HKCSA_SUPER_BLOCKS] + (uint32_t)(q.w[0] & 0xFFFFFFFFu) +
                                       block_rank(q, o)) - __ldg(&t->node_ones[l][lo]);
        if (block_bit(q, o)) { p = r1; lo = mid; }
        else { p -= r1; hi = mid; }
    }
    occ = p;
    return lo;
}

// text position of global row j: LF walk across slices until a marked row
__global__ void __launch_bounds__(256)
ms_locate_kernel(const MultiDesc *__restrict__ d, const uint64_t *__restrict__ rows, uint64_t m,
                 uint64_t *__restrict__ out)
{
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= m) return;
    uint64_t j = rows[q];
    uint32_t steps = 0;
    while (true) {
        const uint32_t s = ms_slice_of(d, j);
        const uint32_t lj = (uint32_t)(j - d->start[s]);
        const BitVec &mk = d->marks[s];
        const uint64_t g = lj / HKCSA_BLOCK_BITS;
        const uint32_t o = lj - (uint32_t)g * HKCSA_BLOCK_BITS;
        const RankBlock b = load_block(mk.blocks + g);
        if (block_bit(b, o)) {
            const uint64_t r = mk.super[g / HKCSA_SUPER_BLOCKS] + (uint32_t)(b.w[0] & 0xFFFFFFFFu) + block_rank(b, o);
            out[q] = (uint64_t)d->samples[s][r] * d->rate + steps;
            return;
        }
        if (steps >= d->rate) { out[q] = ~0ull; return; }       // sentinel not unique: see locate_rows_kernel
        const WtDev &w = d->slice[s];
        uint32_t occ;
        const uint32_t code = ms_access_rank(w, lj, occ);
        const uint32_t byte = w.tab->sym_of_code[code];
        j = d->C[byte] + d->cum[s][byte] + occ;
        ++steps;
    }
}

}  // namespace hkcsa

using namespace hkcsa;

extern "C" size_t hkcsa_multi_desc_bytes(void) { return sizeof(MultiDesc); }

// h_starts: uint64[S+1] global first row of every slice (h_starts[S] = n).  d_ssa_blobs / h_ssa_plans may be
// NULL (no locate).  Builds the descriptor on the host and copies it into d_desc (hkcsa_multi_desc_bytes()).  syncs.
extern "C" int hkcsa_multi_desc_build(uint32_t S, const void *const *d_wt_blobs, const hkcsa_wt_plan *const *h_plans,
                                      const uint64_t *h_starts, const void *const *d_ssa_blobs,
                                      const hkcsa_ssa_plan *const *h_ssa_plans, void *d_desc, void *stream)
{
    HK_REQUIRE(S >= 1 && S <= HKCSA_MAX_SLICES, HKCSA_EINVAL, "1..HKCSA_MAX_SLICES slices");
    HK_REQUIRE(d_wt_blobs && h_plans && h_starts && d_desc, HKCSA_EINVAL, "null pointer");
    static thread_local MultiDesc D;
    memset(&D, 0, sizeof(D));
    D.S = S;
    D.n = h_starts[S];
    HK_REQUIRE(D.n <= (1ull << 40), HKCSA_ERANGE, "n exceeds 2^40");
    uint64_t tot[256];
    memset(tot, 0, sizeof(tot));
    for (uint32_t s = 0; s < S; ++s) {
        HK_REQUIRE(d_wt_blobs[s] && h_plans[s], HKCSA_EINVAL, "null slice");
        HK_REQUIRE(h_plans[s]->n == h_starts[s + 1] - h_starts[s], HKCSA_EINVAL, "slice length != plan length");
        D.slice[s] = make_wt_dev(d_wt_blobs[s], h_plans[s]);
        D.start[s] = h_starts[s];
        for (int c = 0; c < 256; ++c) D.cum[s][c] = tot[c];
        for (uint32_t k = 0; k < h_plans[s]->sigma; ++k) tot[h_plans[s]->sym_of_code[k]] += h_plans[s]->cnt[k];
        if (d_ssa_blobs && h_ssa_plans && d_ssa_blobs[s] && h_ssa_plans[s]) {
            const uint8_t *sb = static_cast<const uint8_t *>(d_ssa_blobs[s]);
            D.marks[s].blocks = reinterpret_cast<const RankBlock *>(sb + h_ssa_plans[s]->off_blocks);
            D.marks[s].super = reinterpret_cast<const uint64_t *>(sb + h_ssa_plans[s]->off_super);
            D.marks[s].len = h_ssa_plans[s]->n;
            D.samples[s] = reinterpret_cast<const uint32_t *>(sb + h_ssa_plans[s]->off_samples);
            D.rate = h_ssa_plans[s]->rate;
        }
    }
    D.start[S] = h_starts[S];
    uint64_t run = 0;
    for (int c = 0; c < 256; ++c) {
        D.cum[S][c] = tot[c];
        D.C[c] = run;
        run += tot[c];
    }
    HK_REQUIRE(run == D.n, HKCSA_EINVAL, "slice symbol counts do not add up to n");
    cudaStream_t st = as_stream(stream);
    HK_CUDA(cudaMemcpyAsync(d_desc, &D, sizeof(D), cudaMemcpyHostToDevice, st));
    HK_CUDA(cudaStreamSynchronize(st));
    return HKCSA_OK;
}

extern "C" int hkcsa_multi_count_batch(const void *d_desc, const uint8_t *d_pat, const int64_t *d_off, uint64_t P,
                                       int64_t *d_lo, int64_t *d_hi, void *stream)
{
    if (P == 0) return HKCSA_OK;
    HK_REQUIRE(d_desc && d_off && d_lo && d_hi, HKCSA_EINVAL, "null pointer");
    cudaStream_t st = as_stream(stream);
    prof::Scope ps(st, prof::COUNT, 0);
    ms_count_kernel<<<(uint32_t)((P + 255) / 256), 256, 0, st>>>(static_cast<const MultiDesc *>(d_desc), d_pat, d_off, P,
                                                                d_lo, d_hi);
    HK_LAUNCH_CHECK();
    return HKCSA_OK;
}

extern "C" int hkcsa_multi_locate_rows(const void *d_desc, const uint64_t *d_rows, uint64_t m, uint64_t *d_out_pos,
                                       void *stream)
{
    if (m == 0) return HKCSA_OK;
    HK_REQUIRE(d_desc && d_rows && d_out_pos, HKCSA_EINVAL, "null pointer");
    cudaStream_t st = as_stream(stream);
    prof::Scope ps(st, prof::LOCATE, 0);
    ms_locate_kernel<<<(uint32_t)((m + 255) / 256), 256, 0, st>>>(static_cast<const MultiDesc *>(d_desc), d_rows, m,
                                                                 d_out_pos);
    HK_LAUNCH_CHECK();
    return HKCSA_OK;
}
