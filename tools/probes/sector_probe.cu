// Random 32-byte sector reads from a buffer far larger than L2: which load flavour moves the fewest bytes / runs fastest.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o sector_probe sector_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x) { x += 0x9E3779B97F4A7C15ULL; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL; x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL; return x ^ (x >> 31); }

template <int MODE>
__device__ __forceinline__ uint64_t load32(const uint8_t *p)
{
    uint64_t a, b, c, d;
    if (MODE == 0) asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    if (MODE == 1) asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    if (MODE == 2) asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    if (MODE == 3) asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    if (MODE == 4) asm volatile("ld.global.cs.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    if (MODE == 5) {   // two 128-bit nc loads (the old load_block)
        uint32_t x0, x1, x2, x3, y0, y1, y2, y3;
        asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3) : "l"(p));
        asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(y0), "=r"(y1), "=r"(y2), "=r"(y3) : "l"(p + 16));
        return (uint64_t)(x0 ^ x1 ^ x2 ^ x3 ^ y0 ^ y1 ^ y2 ^ y3);
    }
    if (MODE == 6) asm volatile("ld.global.nc.L2::64B.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    if (MODE == 7) { uint32_t x; asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(x) : "l"(p)); return x; }
    if (MODE == 8) { uint32_t x; asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(x) : "l"(p)); return x; }
    return a ^ b ^ c ^ d;
}

template <int MODE>
__global__ void probe(const uint8_t *buf, uint64_t sectors, int iters, uint64_t *out)
{
    uint64_t s = mix(blockIdx.x * 1024ull + threadIdx.x), acc = 0;
    for (int i = 0; i < iters; ++i) {
        const uint64_t idx = (s + acc) % sectors;     // dependent chain: one load in flight per thread
        acc += load32<MODE>(buf + idx * 32);
        s = mix(s);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char *name, const uint8_t *buf, uint64_t sectors, uint64_t *out, int blocks, int threads)
{
    const int iters = 64;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    probe<MODE><<<blocks, threads>>>(buf, sectors, 4, out);
    cudaEventRecord(a);
    probe<MODE><<<blocks, threads>>>(buf, sectors, iters, out);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double loads = (double)blocks * threads * iters;
    printf("%-34s blocks %5d x %4d: %7.3f ms  %7.2f G sector-loads/s  (%.2f TB/s of 32 B)\n", name, blocks, threads, ms, loads / ms / 1e6, loads * 32 / ms / 1e9);
}

int main(int argc, char **argv)
{
    const uint64_t bytes = (argc > 1 ? atoll(argv[1]) : 4096ll) << 20;
    uint8_t *buf; uint64_t *out;
    cudaMalloc(&buf, bytes); cudaMemset(buf, 1, bytes);
    cudaMalloc(&out, 1 << 26);
    const uint64_t sectors = bytes / 32;
    for (int blocks : {148 * 8, 148 * 16}) {
        run<0>("ld.global.nc.v4.u64", buf, sectors, out, blocks, 256);
        run<1>("ld.global.v4.u64", buf, sectors, out, blocks, 256);
        run<2>("ld.global.cg.v4.u64", buf, sectors, out, blocks, 256);
        run<3>("ld.global.nc.L1::no_allocate.v4.u64", buf, sectors, out, blocks, 256);
        run<4>("ld.global.cs.v4.u64", buf, sectors, out, blocks, 256);
        run<5>("2 x ld.global.nc.v4.u32", buf, sectors, out, blocks, 256);
        run<6>("ld.global.nc.L2::64B.v4.u64", buf, sectors, out, blocks, 256);
        run<7>("ld.global.nc.u32", buf, sectors, out, blocks, 256);
        run<8>("ld.global.cg.u32", buf, sectors, out, blocks, 256);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
