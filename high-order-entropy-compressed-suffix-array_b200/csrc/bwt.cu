// bwt.cu -- K2: BWT gather.  Replaces bwt_transform (reference csa/bwt.py:3-13):
//   pos = SA[i] - 1; if pos < 0: pos = n - 1; bwt[i] = text[pos].
#include "common.cuh"
#include "prof.cuh"
#include "suffix_array.cuh"
#include <stdlib.h>

namespace hkcsa {

// The gather is random over the whole text, and a text beyond roughly half of the L2 (the two L2 partitions
// each keep their own copy of lines touched from both dies) turns every symbol into a 32-byte DRAM sector
// read.  So the gather runs in PHASES: phase p streams the suffix array again but gathers only the symbols
// whose text position lies in the p-th range of the text (a range that stays L2-resident: evict-last, the
// streams evict-first), and merges them into the output words written by the earlier phases.  Streaming the
// suffix array P times (4 B per entry, sequential) is cheaper than one DRAM sector per symbol.
//
// Each thread handles eight entries: two 16-byte suffix-array loads, one 8-byte BWT word.
template <bool FIRST>
__global__ void __launch_bounds__(256)
bwt_gather_kernel(const uint8_t *__restrict__ text, const uint32_t *__restrict__ sa, uint64_t n,
                  uint8_t *__restrict__ bwt, uint64_t lo, uint64_t span)
{
    const uint64_t pol_keep = l2_policy_evict_last();
    const uint64_t pol_stream = l2_policy_evict_first();
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t i0 = q * 8;
    if (i0 >= n) return;
    if (i0 + 8 <= n) {
        const uint4 a = ld_u32x4_hint(sa + i0, pol_stream), b = ld_u32x4_hint(sa + i0 + 4, pol_stream);
        const uint32_t s[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t w[2] = {0, 0}, m[2] = {0, 0};
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const uint64_t pos = s[t] ? (uint64_t)s[t] - 1 : n - 1;
            if (pos - lo < span) {
                w[t >> 2] |= ld_u8_hint(text + pos, pol_keep) << (8 * (t & 3));
                m[t >> 2] |= 0xFFu << (8 * (t & 3));
            }
        }
        uint2 *out = reinterpret_cast<uint2 *>(bwt + i0);
        if (FIRST) {
            *out = make_uint2(w[0], w[1]);
        } else if (m[0] | m[1]) {
            const uint2 old = *out;
            *out = make_uint2((old.x & ~m[0]) | w[0], (old.y & ~m[1]) | w[1]);
        }
    } else {
        for (uint64_t i = i0; i < n; ++i) {
            const uint32_t v = sa[i];
            const uint64_t pos = v ? (uint64_t)v - 1 : n - 1;
            if (pos - lo < span) bwt[i] = text[pos];
        }
    }
}

__global__ void __launch_bounds__(256)
bwt_gather_scalar_kernel(const uint8_t *__restrict__ text, const uint32_t *__restrict__ sa, uint64_t n,
                         uint8_t *__restrict__ bwt)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t v = sa[i];
    bwt[i] = text[v ? (uint64_t)v - 1 : n - 1];
}

int bwt_gather(const uint8_t *d_text, const uint32_t *d_sa, uint64_t n, uint8_t *d_bwt, cudaStream_t st)
{
    if (n == 0) return HKCSA_OK;
    if (((reinterpret_cast<uintptr_t>(d_sa) & 15) | (reinterpret_cast<uintptr_t>(d_bwt) & 7)) != 0) {
        // views at odd offsets: plain element-wise gather
        prof::Scope ps(st, prof::BWT_GATHER, n * 6);
        bwt_gather_scalar_kernel<<<(uint32_t)((n + 255) / 256), 256, 0, st>>>(d_text, d_sa, n, d_bwt);
        HK_LAUNCH_CHECK();
        return HKCSA_OK;
    }
    // text range per phase: what stays L2-resident next to the streams (tunable: HKCSA_BWT_RANGE_MB)
    static double range_mb = -1;
    if (range_mb < 0) {
        const char *e = getenv("HKCSA_BWT_RANGE_MB");
        range_mb = e ? atof(e) : 70.0;
        if (range_mb <= 0) range_mb = 1e12;
    }
    uint64_t phases = (uint64_t)((double)n / (range_mb * 1e6)) + 1;
    phases = std::min<uint64_t>(phases, 8);
    const uint64_t span = (n + phases - 1) / phases;
    prof::Scope ps(st, prof::BWT_GATHER, n * (4 * phases + 2 * phases));
    const uint64_t threads = (n + 7) / 8;
    const uint32_t grid = (uint32_t)((threads + 255) / 256);
    for (uint64_t p = 0; p < phases; ++p) {
        if (p == 0) bwt_gather_kernel<true><<<grid, 256, 0, st>>>(d_text, d_sa, n, d_bwt, 0, span);
        else bwt_gather_kernel<false><<<grid, 256, 0, st>>>(d_text, d_sa, n, d_bwt, p * span, span);
        count_launch();
    }
    HK_CUDA(cudaGetLastError());
    return HKCSA_OK;
}

}  // namespace hkcsa

using namespace hkcsa;

extern "C" int hkcsa_bwt(const uint8_t *d_text, const uint32_t *d_sa, uint64_t n, uint8_t *d_bwt, void *stream)
{
    if (n == 0) return HKCSA_OK;
    HK_REQUIRE(d_text && d_sa && d_bwt, HKCSA_EINVAL, "null pointer");
    HK_REQUIRE(n <= HKCSA_MAX_N, HKCSA_ERANGE, "n exceeds HKCSA_MAX_N");
    return bwt_gather(d_text, d_sa, n, d_bwt, as_stream(stream));
}
