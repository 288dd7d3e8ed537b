// bwt.cu -- K2: BWT gather.  Replaces bwt_transform (reference csa/bwt.py:3-13):
//   pos = SA[i] - 1; if pos < 0: pos = n - 1; bwt[i] = text[pos].
#include "common.cuh"
#include "prof.cuh"

namespace hkcsa {

// Each thread gathers four symbols and writes them as one 32-bit word.
__global__ void __launch_bounds__(256)
bwt_gather_kernel(const uint8_t *__restrict__ text, const uint32_t *__restrict__ sa, uint64_t n,
                  uint8_t *__restrict__ bwt)
{
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t i0 = q * 4;
    if (i0 >= n) return;
    if (i0 + 4 <= n && (reinterpret_cast<uintptr_t>(sa) & 15) == 0 && (reinterpret_cast<uintptr_t>(bwt) & 3) == 0) {
        const uint4 s4 = *reinterpret_cast<const uint4 *>(sa + i0);
        const uint32_t s[4] = {s4.x, s4.y, s4.z, s4.w};
        uint32_t w = 0;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const uint64_t pos = s[t] ? (uint64_t)s[t] - 1 : n - 1;
            w |= (uint32_t)__ldg(text + pos) << (8 * t);
        }
        *reinterpret_cast<uint32_t *>(bwt + i0) = w;
    } else {
        for (uint64_t i = i0; i < n && i < i0 + 4; ++i) {
            const uint32_t v = sa[i];
            bwt[i] = text[v ? (uint64_t)v - 1 : n - 1];
        }
    }
}

}  // namespace hkcsa

using namespace hkcsa;

extern "C" int hkcsa_bwt(const uint8_t *d_text, const uint32_t *d_sa, uint64_t n, uint8_t *d_bwt, void *stream)
{
    if (n == 0) return HKCSA_OK;
    HK_REQUIRE(d_text && d_sa && d_bwt, HKCSA_EINVAL, "null pointer");
    HK_REQUIRE(n <= HKCSA_MAX_N, HKCSA_ERANGE, "n exceeds HKCSA_MAX_N");
    cudaStream_t st = as_stream(stream);
    prof::Scope ps(st, prof::BWT_GATHER, n * 6);
    const uint64_t threads = (n + 3) / 4;
    bwt_gather_kernel<<<(uint32_t)((threads + 255) / 256), 256, 0, st>>>(d_text, d_sa, n, d_bwt);
    HK_LAUNCH_CHECK();
    return HKCSA_OK;
}
