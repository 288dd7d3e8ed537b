// prof.cuh -- optional CUDA-event timing of the library's own kernel launches,
// on the stream they are launched on.  Off by default (zero overhead beyond a
// branch); bench.py switches it on to report per-kernel durations and the
// roofline of the dominant kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hkcsa {
namespace prof {

enum Class {
    BYTE_HIST = 0, SA_PACK0, SA_KEYBUILD, RADIX_SCAN, ONESWEEP_U64, SEG_REDUCE, SEG_SCAN, SEG_APPLY,
    BWT_GATHER, WT_LEVELS, WT_PACK, WT_DIR, COUNT, LOCATE, SSA_BUILD, OTHER, NUM_CLASSES
};
constexpr int MAX_RECORDS = 16384;

bool enabled();

// Records an event before and after whatever is enqueued during its lifetime.
struct Scope {
    Scope(cudaStream_t st, int cls, uint64_t alg_bytes);
    ~Scope();
    cudaStream_t st_;
    int slot_;
};

}  // namespace prof
}  // namespace hkcsa
