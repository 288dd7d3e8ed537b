// common.cu -- error string, pinned scalar page, optional per-kernel event profiler.
#include "common.cuh"
#include "prof.cuh"

#include <stdarg.h>
#include <mutex>

namespace hkcsa {

static thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// one page per calling thread: two host threads (or two devices driven from two threads) building at the same time
// must not share the histogram / survivor-count read-back slots
void *pinned_page()
{
    static thread_local void *page = nullptr;
    if (!page) {
        cudaError_t e = cudaHostAlloc(&page, 4096, cudaHostAllocDefault);
        if (e != cudaSuccess) {
            set_error("cudaHostAlloc(4096) -> %s", cudaGetErrorString(e));
            page = nullptr;
        }
    }
    return page;
}

// ---------------------------------------------------------------- profiler
namespace prof {
static bool g_on = false;
static uint32_t g_mask = 0xFFFFFFFFu;     // classes that are timed while g_on
struct Rec {
    cudaEvent_t a, b;
    int cls;
    uint64_t bytes;
};
static Rec g_rec[MAX_RECORDS];
static int g_created = 0;
static int g_used = 0;
static const char *g_names[NUM_CLASSES] = {
    "byte_hist", "sa_pack0", "sa_keybuild", "radix_scan", "onesweep_u64", "seg_reduce", "seg_scan",
    "seg_apply", "bwt_gather", "wt_levels", "wt_pack", "wt_dir", "count", "locate", "ssa_build", "other"};

bool enabled() { return g_on; }

static std::mutex g_mu;     // the record array is shared by every thread that launches kernels

Scope::Scope(cudaStream_t st, int cls, uint64_t bytes) : st_(st), slot_(-1)
{
    if (!g_on || !((g_mask >> cls) & 1u)) return;
    {
        std::lock_guard<std::mutex> lock(g_mu);
        if (g_used >= MAX_RECORDS) return;
        if (g_used >= g_created) {
            if (cudaEventCreate(&g_rec[g_created].a) != cudaSuccess) return;
            if (cudaEventCreate(&g_rec[g_created].b) != cudaSuccess) return;
            ++g_created;
        }
        slot_ = g_used++;
    }
    g_rec[slot_].cls = cls;
    g_rec[slot_].bytes = bytes;
    cudaEventRecord(g_rec[slot_].a, st_);
}
Scope::~Scope()
{
    if (slot_ >= 0) cudaEventRecord(g_rec[slot_].b, st_);
}
}  // namespace prof
}  // namespace hkcsa

using namespace hkcsa;

extern "C" int hkcsa_abi_version(void) { return HKCSA_ABI_VERSION; }
extern "C" const char *hkcsa_last_error(void) { return g_err; }

extern "C" unsigned long long hkcsa_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" size_t hkcsa_struct_size(int which)
{
    switch (which) {
        case 0: return sizeof(hkcsa_sa_stats);
        case 1: return sizeof(hkcsa_wt_plan);
        case 2: return sizeof(hkcsa_ssa_plan);
        case 3: return sizeof(hkcsa_prof_entry);
        case 4: return sizeof(hkcsa_occ_plan);
        case 5: return sizeof(hkcsa_dsa_plan);
        case 6: return sizeof(hkcsa_rrr_plan);
        default: return 0;
    }
}

extern "C" int hkcsa_prof_enable(int on)
{
    prof::g_on = (on != 0);
    prof::g_mask = 0xFFFFFFFFu;
    return HKCSA_OK;
}
extern "C" int hkcsa_prof_enable_classes(uint32_t class_mask)
{
    prof::g_on = (class_mask != 0);
    prof::g_mask = class_mask;
    return HKCSA_OK;
}
extern "C" int hkcsa_prof_class_index(const char *name)
{
    for (int c = 0; c < prof::NUM_CLASSES; ++c)
        if (name && strcmp(name, prof::g_names[c]) == 0) return c;
    return -1;
}
extern "C" int hkcsa_prof_reset(void)
{
    prof::g_used = 0;
    return HKCSA_OK;
}
// every bracketed launch in recording order: start / end in ms relative to the first record's start, and its class
extern "C" int hkcsa_prof_timeline(float *h_start_ms, float *h_end_ms, int *h_class, int max_entries, int *h_n)
{
    HK_REQUIRE(h_start_ms && h_end_ms && h_class && h_n, HKCSA_EINVAL, "null pointer");
    const int n = prof::g_used < max_entries ? prof::g_used : max_entries;
    for (int i = 0; i < n; ++i) {
        HK_CUDA(cudaEventSynchronize(prof::g_rec[i].b));
        HK_CUDA(cudaEventElapsedTime(&h_start_ms[i], prof::g_rec[0].a, prof::g_rec[i].a));
        HK_CUDA(cudaEventElapsedTime(&h_end_ms[i], prof::g_rec[0].a, prof::g_rec[i].b));
        h_class[i] = prof::g_rec[i].cls;
    }
    *h_n = n;
    return HKCSA_OK;
}

extern "C" int hkcsa_prof_read(hkcsa_prof_entry *h_out, int max_entries, int *h_n)
{
    HK_REQUIRE(h_out && h_n && max_entries >= prof::NUM_CLASSES, HKCSA_EINVAL, "need room for every class");
    for (int c = 0; c < prof::NUM_CLASSES; ++c) {
        memset(&h_out[c], 0, sizeof(hkcsa_prof_entry));
        strncpy(h_out[c].name, prof::g_names[c], sizeof(h_out[c].name) - 1);
    }
    for (int i = 0; i < prof::g_used; ++i) {
        HK_CUDA(cudaEventSynchronize(prof::g_rec[i].b));
        float ms = 0.f;
        HK_CUDA(cudaEventElapsedTime(&ms, prof::g_rec[i].a, prof::g_rec[i].b));
        hkcsa_prof_entry &e = h_out[prof::g_rec[i].cls];
        e.launches += 1;
        e.ms += ms;
        e.alg_bytes += prof::g_rec[i].bytes;
    }
    *h_n = prof::NUM_CLASSES;
    return HKCSA_OK;
}
