// index_build.cu -- the whole constructor of the reference's EnhancedFMIndex (csa/enhanced_fm_index.py:8-13:
// text -> build_suffix_array -> bwt_transform -> build_occ / build_count) in ONE call: suffix array + BWT
// (hkcsa_sa_bwt_build), then the wavelet tree over the BWT and the sampled suffix array side by side on two streams.
// The host comes back between the kernels only where a result decides what is launched next (the byte histogram, the
// survivors of a refinement round); nothing is allocated and no host-language code runs between the suffix-array
// rounds and the tree kernels, which is what a caller stitching the four entry points together cannot avoid.
#include "common.cuh"
#include "wavelet.cuh"

using namespace hkcsa;

namespace {
struct BuildEvents {
    cudaEvent_t sa_done = nullptr, tree_done = nullptr, ssa_done = nullptr;
    int device = -1;
};
// one set of events per calling thread and device
int build_events(BuildEvents **out)
{
    static thread_local BuildEvents ev;
    int dev = 0;
    HK_CUDA(cudaGetDevice(&dev));
    if (ev.device != dev) {
        if (ev.sa_done) { cudaEventDestroy(ev.sa_done); cudaEventDestroy(ev.tree_done); cudaEventDestroy(ev.ssa_done); }
        HK_CUDA(cudaEventCreateWithFlags(&ev.sa_done, cudaEventDisableTiming));
        HK_CUDA(cudaEventCreateWithFlags(&ev.tree_done, cudaEventDisableTiming));
        HK_CUDA(cudaEventCreateWithFlags(&ev.ssa_done, cudaEventDisableTiming));
        ev.device = dev;
    }
    *out = &ev;
    return HKCSA_OK;
}
}  // namespace

extern "C" int hkcsa_index_build(const uint8_t *d_text, uint64_t n, uint32_t *d_sa, uint8_t *d_bwt, void *d_sa_scratch,
                                 size_t sa_scratch_bytes, hkcsa_wt_plan *h_wt_plan, void *d_wt_blob, size_t wt_blob_cap,
                                 void *d_wt_scratch, size_t wt_scratch_cap, const hkcsa_ssa_plan *h_ssa_plan,
                                 void *d_ssa_blob, void *d_ssa_scratch, size_t ssa_scratch_bytes, void *stream,
                                 void *stream_tree, void *stream_ssa, hkcsa_sa_stats *h_stats)
{
    HK_REQUIRE(h_wt_plan && h_stats, HKCSA_EINVAL, "null pointer");
    HK_REQUIRE(!h_ssa_plan || h_ssa_plan->n == n, HKCSA_EINVAL, "sampled-SA plan is for another length");
    cudaStream_t main = as_stream(stream);
    cudaStream_t tree = stream_tree ? as_stream(stream_tree) : main;
    cudaStream_t ssa = stream_ssa ? as_stream(stream_ssa) : main;
    int rc = hkcsa_sa_bwt_build(d_text, n, d_sa, d_bwt, d_sa_scratch, sa_scratch_bytes, stream, h_stats);
    if (rc != HKCSA_OK) return rc;
    // the BWT is a permutation of the text: the byte histogram of the suffix-array build serves the tree
    rc = hkcsa_wt_plan_from_hist(h_stats->byte_hist, h_wt_plan);
    if (rc != HKCSA_OK) return rc;
    HK_REQUIRE(h_wt_plan->blob_bytes <= wt_blob_cap, HKCSA_ESCRATCH, "wavelet blob too small (hkcsa_wt_blob_bound)");
    HK_REQUIRE(h_wt_plan->scratch_bytes <= wt_scratch_cap, HKCSA_ESCRATCH, "wavelet scratch too small (hkcsa_wt_scratch_bound)");
    BuildEvents *ev = nullptr;
    rc = build_events(&ev);
    if (rc != HKCSA_OK) return rc;
    if (tree != main || ssa != main) {
        HK_CUDA(cudaEventRecord(ev->sa_done, main));
        if (tree != main) HK_CUDA(cudaStreamWaitEvent(tree, ev->sa_done, 0));
        if (ssa != main && h_ssa_plan) HK_CUDA(cudaStreamWaitEvent(ssa, ev->sa_done, 0));
    }
    // the sampled SA goes out first: its launches must be queued before hkcsa_wt_build blocks the host at its end
    if (h_ssa_plan) {
        rc = hkcsa_ssa_build(d_sa, h_ssa_plan, d_ssa_blob, d_ssa_scratch, ssa_scratch_bytes, ssa);
        if (rc != HKCSA_OK) return rc;
        if (ssa != main) HK_CUDA(cudaEventRecord(ev->ssa_done, ssa));
    }
    rc = hkcsa_wt_build(d_bwt, h_wt_plan, d_wt_blob, d_wt_scratch, wt_scratch_cap, tree);
    if (rc != HKCSA_OK) return rc;
    if (tree != main) {
        HK_CUDA(cudaEventRecord(ev->tree_done, tree));
        HK_CUDA(cudaStreamWaitEvent(main, ev->tree_done, 0));
    }
    if (h_ssa_plan && ssa != main) HK_CUDA(cudaStreamWaitEvent(main, ev->ssa_done, 0));
    return HKCSA_OK;
}
