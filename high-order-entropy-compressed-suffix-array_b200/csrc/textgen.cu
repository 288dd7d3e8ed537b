// textgen.cu -- seeded synthetic workloads on the device (no reference
// counterpart: the Pizza&Chili corpora of tests/dataset_benchmark.py:10-16 are
// unreachable offline).  Byte-for-byte identical to oracle/hkcsa_oracle.c
// hko_gen_text / hko_pattern_*: integer-only, counter-based, 64 KiB chunks with
// a context reset so chunks are independent (one thread per chunk).
#include "common.cuh"

namespace hkcsa {

constexpr uint64_t GEN_CHUNK = 65536;

__constant__ uint8_t c_perm4[24][4] = {
    {0,1,2,3},{0,1,3,2},{0,2,1,3},{0,2,3,1},{0,3,1,2},{0,3,2,1},
    {1,0,2,3},{1,0,3,2},{1,2,0,3},{1,2,3,0},{1,3,0,2},{1,3,2,0},
    {2,0,1,3},{2,0,3,1},{2,1,0,3},{2,1,3,0},{2,3,0,1},{2,3,1,0},
    {3,0,1,2},{3,0,2,1},{3,1,0,2},{3,1,2,0},{3,2,0,1},{3,2,1,0}};

__device__ __forceinline__ uint8_t eng96_symbol(uint32_t id)
{
    // 0x20..0x7E without '$' (94 symbols), then '\n', '\t'
    if (id >= 94) return id == 94 ? 0x0A : 0x09;
    const uint32_t c = 0x20 + id;
    return (uint8_t)(c >= 0x24 ? c + 1 : c);
}

__global__ void gen_text_kernel(int kind, uint64_t seed, uint64_t n, uint8_t *__restrict__ out)
{
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t lo = q * GEN_CHUNK;
    if (lo >= n) return;
    const uint64_t hi = min(lo + GEN_CHUNK, n);
    const uint64_t s1 = splitmix64(seed);
    const uint64_t s2 = splitmix64(seed ^ 0xA5A5A5A5A5A5A5A5ULL);
    if (kind == 0) {
        uint32_t c1 = 0, c2 = 0, c3 = 0;
        for (uint64_t i = lo; i < hi; ++i) {
            const uint64_t r = splitmix64(s1 + i);
            const uint64_t ctx = (uint64_t)c1 + 96u * c2 + 9216u * c3;
            const uint64_t h = splitmix64(s2 ^ ctx);
            const uint32_t u = (uint32_t)(r % 147u);
            const int k = (u < 60) ? 0 : (u < 90) ? 1 : (u < 110) ? 2 : (u < 125) ? 3 : (u < 137) ? 4 : 5;
            const uint32_t id = (uint32_t)((h >> (10 * k)) & 1023u) % 96u;
            out[i] = eng96_symbol(id);
            c3 = c2; c2 = c1; c1 = id;
        }
    } else {
        uint32_t ctx = 0;
        for (uint64_t i = lo; i < hi; ++i) {
            const uint64_t r = splitmix64(s1 + i);
            const uint32_t pidx = (uint32_t)(splitmix64(s2 ^ (uint64_t)ctx) % 24u);
            const uint32_t r4 = (uint32_t)(r & 15u);
            const int slot = (r4 < 8) ? 0 : (r4 < 12) ? 1 : (r4 < 14) ? 2 : 3;
            const uint32_t id = c_perm4[pidx][slot];
            out[i] = (uint8_t)("ACGT"[id]);
            ctx = ((ctx << 2) | id) & 1023u;
        }
    }
}

__global__ void gen_pattern_len_kernel(uint64_t seed, uint64_t P, uint32_t min_len, uint32_t max_len, uint64_t n,
                                       uint32_t *__restrict__ len_out)
{
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const uint64_t s1 = splitmix64(seed);
    const uint64_t r1 = splitmix64(s1 + 3 * p);
    uint32_t len = min_len + (uint32_t)(r1 % (uint64_t)(max_len - min_len + 1));
    if (len > n) len = (uint32_t)n;
    len_out[p] = len;
}

__global__ void gen_pattern_fill_kernel(uint64_t seed, uint64_t P, const uint8_t *__restrict__ text, uint64_t n,
                                        const uint8_t *__restrict__ alpha, uint32_t sigma,
                                        const int64_t *__restrict__ off, uint8_t *__restrict__ out)
{
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const uint64_t s1 = splitmix64(seed);
    const uint64_t r2 = splitmix64(s1 + 3 * p + 1);
    const uint64_t r3 = splitmix64(s1 + 3 * p + 2);
    const uint32_t len = (uint32_t)(off[p + 1] - off[p]);
    const uint64_t start = r2 % (n - len + 1);
    uint8_t *dst = out + off[p];
    for (uint32_t k = 0; k < len; ++k) dst[k] = text[start + k];
    if ((r3 & 1u) && len > 0 && sigma > 1) {
        const uint32_t at = (uint32_t)((r3 >> 1) % len);
        uint32_t pick = (uint32_t)((r3 >> 32) % sigma);
        if (alpha[pick] == dst[at]) pick = (pick + 1) % sigma;
        dst[at] = alpha[pick];
    }
}

}  // namespace hkcsa

using namespace hkcsa;

extern "C" int hkcsa_gen_text(int kind, uint64_t seed, uint64_t n, uint8_t *d_text, void *stream)
{
    HK_REQUIRE(kind == 0 || kind == 1, HKCSA_EINVAL, "kind must be 0 (ENG96) or 1 (DNA4)");
    if (n == 0) return HKCSA_OK;
    HK_REQUIRE(d_text != nullptr, HKCSA_EINVAL, "null pointer");
    const uint64_t chunks = (n + GEN_CHUNK - 1) / GEN_CHUNK;
    gen_text_kernel<<<(uint32_t)((chunks + 63) / 64), 64, 0, as_stream(stream)>>>(kind, seed, n, d_text);
    HK_LAUNCH_CHECK();
    return HKCSA_OK;
}

extern "C" int hkcsa_gen_pattern_lengths(uint64_t seed, uint64_t P, uint32_t min_len, uint32_t max_len, uint64_t n,
                                         uint32_t *d_len, void *stream)
{
    HK_REQUIRE(min_len <= max_len, HKCSA_EINVAL, "min_len > max_len");
    if (P == 0) return HKCSA_OK;
    HK_REQUIRE(d_len != nullptr, HKCSA_EINVAL, "null pointer");
    gen_pattern_len_kernel<<<(uint32_t)((P + 255) / 256), 256, 0, as_stream(stream)>>>(seed, P, min_len, max_len, n,
                                                                                       d_len);
    HK_LAUNCH_CHECK();
    return HKCSA_OK;
}

extern "C" int hkcsa_gen_pattern_bytes(uint64_t seed, uint64_t P, const uint8_t *d_text, uint64_t n,
                                       const uint8_t *d_alphabet, uint32_t sigma, const int64_t *d_offsets,
                                       uint8_t *d_out, void *stream)
{
    if (P == 0) return HKCSA_OK;
    HK_REQUIRE(d_text && d_alphabet && d_offsets && d_out, HKCSA_EINVAL, "null pointer");
    gen_pattern_fill_kernel<<<(uint32_t)((P + 255) / 256), 256, 0, as_stream(stream)>>>(seed, P, d_text, n, d_alphabet,
                                                                                        sigma, d_offsets, d_out);
    HK_LAUNCH_CHECK();
    return HKCSA_OK;
}
