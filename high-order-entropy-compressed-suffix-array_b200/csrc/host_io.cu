// host_io.cu -- pageable host memory <-> device through a pinned ring filled by several host threads.
//
// The reference hands a Python str to EnhancedFMIndex (csa/enhanced_fm_index.py:8-9); the bytes of that str live in
// pageable memory.  A plain cudaMemcpy from pageable memory is staged by the driver through one small pinned
// buffer on the calling thread (one core's memcpy rate, ~10 GB/s, and the DMA waits for it).  Here T host threads
// copy 2 MB chunks into their own pinned slots and enqueue the DMA of each chunk on the caller's stream as soon as
// it is filled: the host copy runs at T cores' rate and overlaps the PCIe transfer.  The call returns when the last
// chunk has been STAGED (the source may be released); the DMAs complete in stream order.
#include "common.cuh"

#include <stdlib.h>
#include <mutex>
#include <thread>
#include <vector>

namespace hkcsa {
namespace {

constexpr size_t STAGE_CHUNK = 2u << 20;
constexpr int STAGE_MAX_THREADS = 16;
constexpr int STAGE_SLOTS = 2;
constexpr int STAGE_MAX_DEVICES = 16;

struct Stage {
    std::mutex mu;                                           // one staged copy at a time per device
    uint8_t *pin = nullptr;                                  // [threads][slots][chunk]
    cudaEvent_t ev[STAGE_MAX_THREADS][STAGE_SLOTS];
    bool busy[STAGE_MAX_THREADS][STAGE_SLOTS];
    int events = 0;                                          // events created so far (a failed init resumes here)
    bool ready = false;
};
Stage g_stage[STAGE_MAX_DEVICES];

int stage_threads(size_t nbytes, int asked)
{
    int t = asked;
    if (t <= 0) {
        const char *e = getenv("HKCSA_STAGE_THREADS");
        t = e ? atoi(e) : 0;
    }
    // measured on the 16-core box, 100 MB: 2 threads 9.1 ms, 4: 7.4, 6: 7.8, 8: 8.1, 12: 8.7, 16: 9.2 (whole constructor)
    if (t <= 0) t = (int)std::max(1u, std::min(4u, std::thread::hardware_concurrency()));
    t = std::min(t, STAGE_MAX_THREADS);
    const size_t chunks = (nbytes + STAGE_CHUNK - 1) / STAGE_CHUNK;
    return (int)std::max<size_t>(1, std::min<size_t>((size_t)t, chunks));
}

cudaError_t stage_init(Stage &s)
{
    if (s.ready) return cudaSuccess;
    cudaError_t e = cudaSuccess;
    if (!s.pin) {
        e = cudaHostAlloc((void **)&s.pin, (size_t)STAGE_MAX_THREADS * STAGE_SLOTS * STAGE_CHUNK, cudaHostAllocDefault);
        if (e != cudaSuccess) { s.pin = nullptr; return e; }
    }
    for (; s.events < STAGE_MAX_THREADS * STAGE_SLOTS; ++s.events) {
        const int t = s.events / STAGE_SLOTS, k = s.events % STAGE_SLOTS;
        e = cudaEventCreateWithFlags(&s.ev[t][k], cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
        s.busy[t][k] = false;
    }
    s.ready = true;
    return cudaSuccess;
}

// worker t moves chunks t, t + T, t + 2T, ... ; to_device: host -> pinned -> device, else device -> pinned -> host
cudaError_t stage_worker(Stage &s, int dev, int t, int T, uint8_t *d, uint8_t *h, size_t nbytes, bool to_device,
                         cudaStream_t st)
{
    cudaError_t e = cudaSetDevice(dev);
    if (e != cudaSuccess) return e;
    const size_t chunks = (nbytes + STAGE_CHUNK - 1) / STAGE_CHUNK;
    if (to_device) {
        size_t it = 0;
        for (size_t c = t; c < chunks; c += T, ++it) {
            const int k = (int)(it % STAGE_SLOTS);
            uint8_t *slot = s.pin + ((size_t)t * STAGE_SLOTS + k) * STAGE_CHUNK;
            const size_t off = c * STAGE_CHUNK, len = std::min(STAGE_CHUNK, nbytes - off);
            if (s.busy[t][k] && (e = cudaEventSynchronize(s.ev[t][k])) != cudaSuccess) return e;
            memcpy(slot, h + off, len);
            if ((e = cudaMemcpyAsync(d + off, slot, len, cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
            if ((e = cudaEventRecord(s.ev[t][k], st)) != cudaSuccess) return e;
            s.busy[t][k] = true;
        }
        return cudaSuccess;
    }
    // device -> host: keep both slots' DMAs in flight, drain the older one while the newer runs
    size_t pend_c[STAGE_SLOTS];
    bool pend[STAGE_SLOTS] = {false, false};
    auto drain = [&](int k) -> cudaError_t {
        if (!pend[k]) return cudaSuccess;
        cudaError_t e2 = cudaEventSynchronize(s.ev[t][k]);
        if (e2 != cudaSuccess) return e2;
        const size_t off = pend_c[k] * STAGE_CHUNK, len = std::min(STAGE_CHUNK, nbytes - off);
        memcpy(h + off, s.pin + ((size_t)t * STAGE_SLOTS + k) * STAGE_CHUNK, len);
        pend[k] = false;
        s.busy[t][k] = false;
        return cudaSuccess;
    };
    size_t it = 0;
    for (size_t c = t; c < chunks; c += T, ++it) {
        const int k = (int)(it % STAGE_SLOTS);
        if ((e = drain(k)) != cudaSuccess) return e;
        if (s.busy[t][k] && (e = cudaEventSynchronize(s.ev[t][k])) != cudaSuccess) return e;
        uint8_t *slot = s.pin + ((size_t)t * STAGE_SLOTS + k) * STAGE_CHUNK;
        const size_t off = c * STAGE_CHUNK, len = std::min(STAGE_CHUNK, nbytes - off);
        if ((e = cudaMemcpyAsync(slot, d + off, len, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
        if ((e = cudaEventRecord(s.ev[t][k], st)) != cudaSuccess) return e;
        pend[k] = true;
        pend_c[k] = c;
    }
    for (size_t j = 0; j < STAGE_SLOTS; ++j)
        if ((e = drain((int)((it + j) % STAGE_SLOTS))) != cudaSuccess) return e;
    return cudaSuccess;
}

int staged_copy(uint8_t *d, uint8_t *h, size_t nbytes, bool to_device, int threads, cudaStream_t st)
{
    if (nbytes == 0) return HKCSA_OK;
    HK_REQUIRE(d && h, HKCSA_EINVAL, "null pointer");
    int dev = 0;
    HK_CUDA(cudaGetDevice(&dev));
    HK_REQUIRE(dev >= 0 && dev < STAGE_MAX_DEVICES, HKCSA_EINVAL, "device ordinal beyond the staging table");
    Stage &s = g_stage[dev];
    std::lock_guard<std::mutex> lock(s.mu);
    HK_CUDA(stage_init(s));
    const int T = stage_threads(nbytes, threads);
    if (T == 1) {
        HK_CUDA(stage_worker(s, dev, 0, 1, d, h, nbytes, to_device, st));
        return HKCSA_OK;
    }
    std::vector<std::thread> pool;
    std::vector<cudaError_t> err((size_t)T, cudaSuccess);
    for (int t = 1; t < T; ++t)
        pool.emplace_back([&, t] { err[(size_t)t] = stage_worker(s, dev, t, T, d, h, nbytes, to_device, st); });
    err[0] = stage_worker(s, dev, 0, T, d, h, nbytes, to_device, st);
    for (auto &th : pool) th.join();
    for (int t = 0; t < T; ++t) HK_CUDA(err[(size_t)t]);
    return HKCSA_OK;
}

}  // namespace
}  // namespace hkcsa

using namespace hkcsa;

extern "C" int hkcsa_h2d_staged(void *d_dst, const void *h_src, size_t nbytes, int threads, void *stream)
{
    return staged_copy(static_cast<uint8_t *>(d_dst), const_cast<uint8_t *>(static_cast<const uint8_t *>(h_src)),
                       nbytes, true, threads, as_stream(stream));
}

extern "C" int hkcsa_d2h_staged(void *h_dst, const void *d_src, size_t nbytes, int threads, void *stream)
{
    return staged_copy(const_cast<uint8_t *>(static_cast<const uint8_t *>(d_src)), static_cast<uint8_t *>(h_dst),
                       nbytes, false, threads, as_stream(stream));
}
